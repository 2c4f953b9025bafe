"""
Batched full-covariance Gaussian mixture EM -- the *definition* shared by this
oracle and the CUDA kernel.  PARITY UNPINNED against the reference.

The reference fits its mixtures with the third-party package ``gmm-gpu==0.2.8``
(``from gmm_gpu.gmm import GMM``, /root/reference/nanocompore/comparisons.py:9;
call sites comparisons.py:341-361 and plotting.py:463-489; pinned in
pyproject.toml:29 / uv.lock:464-474).  That package is not vendored in
/root/reference, is not installed and cannot be downloaded, so its arithmetic
cannot be read.  This module exposes exactly the API surface the reference uses
(ctor kwargs ``n_components, device, random_seed, dtype``; ``fit(X[B,N,D])`` with NaN
rows = missing reads; ``bic(X) -> (B,)``; ``predict_proba(X, force_cpu_result=False)
-> (B,N,K)`` with NaN for missing reads; ``means[k] -> (B,D)``; ``covs[k] -> (B,D,D)``)
and defines the fit as scikit-learn ``GaussianMixture(covariance_type='full')`` does
(sklearn/mixture/_base.py:244-283, _gaussian_mixture.py:168-198, 282-319, 945-978):

* valid reads of a position = reads with every variable non-NaN; N = their count
* E-step  log p(x|k) = -1/2 [D ln 2pi + ln|S_k| + (x-m_k)' S_k^-1 (x-m_k)]
          log r_k = ln w_k + log p(x|k) - logsumexp_k ;  L = mean_valid(logsumexp_k)
* M-step  N_k = sum r_k + 10*eps64 ; w_k = N_k / sum_k N_k ; m_k = sum r_k x / N_k
          S_k = sum r_k (x-m_k)(x-m_k)' / N_k + reg_covar*I          (reg_covar = 1e-6)
* loop    for it in 1..max_iter(100): E, M, stop when |L_it - L_(it-1)| < tol (1e-3)
          (L_0 = -inf); then one final E-step with the final parameters gives the
          posteriors and the log-likelihood used by ``bic``
* BIC     -2 N L + n_params ln N,  n_params = K D + K D (D+1)/2 + K - 1
* K = 1   closed form (sample mean, population covariance + reg_covar*I)
* K = 2   initialisation is deterministic (no RNG, so ``random_seed`` is accepted and
          ignored) and k-means based, like scikit-learn's default ``init_params='kmeans'``:
          centre 0 = the valid read farthest from the mean of the valid reads, centre 1 = the
          valid read farthest from centre 0 (ties -> lowest read index); then Lloyd iterations
          (assign every read to the nearer centre, tie -> cluster 0; centres = cluster means; an
          emptied cluster keeps its centre) until no assignment changes, at most 20; one M-step
          on the final 0/1 responsibilities gives the starting parameters.
          (A first version of this definition seeded the second component on the single
          farthest read; that isolates an extreme read in a component of its own and failed the
          reference's own tests tests/test_comparisons.py:464-504.)
* fewer than 2 valid reads: the fit is undefined -> BIC and posteriors are NaN.

Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

EPS10 = 10 * np.finfo(np.float64).eps
LOG_2PI = float(np.log(2 * np.pi))


class GMM:
    def __init__(self, n_components, device='cpu', random_seed=None, dtype=None,
                 max_iter=100, tol=1e-3, reg_covar=1e-6, work_dtype=np.float64,
                 kmeans_max_iter=20):
        if n_components not in (1, 2):
            raise NotImplementedError("only K=1 and K=2 are used by nanocompore")
        self.K = n_components
        self.max_iter = max_iter
        self.tol = tol
        self.reg = reg_covar
        self.wd = work_dtype
        self.kmeans_max_iter = kmeans_max_iter
        self.n_iter = None
        self._torch_out = False

    # ------------------------------------------------------------------ utils
    @staticmethod
    def _dist2(Xz, c):
        """squared distance of every read to a centre per position: dx*dx + dy*dy (+ ...),
        each product and sum rounded once (no fused multiply-add)."""
        d = Xz - c[:, None, :]
        sq = d * d
        out = sq[:, :, 0]
        for k in range(1, sq.shape[2]):
            out = out + sq[:, :, k]
        return out

    def _prep(self, X):
        try:
            import torch
            if isinstance(X, torch.Tensor):
                self._torch_out = True
                X = X.detach().cpu().numpy()
        except ImportError:
            pass
        X = np.asarray(X)
        valid = ~np.isnan(X).any(2)                       # (B,N)
        Xz = np.where(valid[:, :, None], X, 0).astype(self.wd)
        return Xz, valid

    def _moments(self, Xz, R):
        """M-step for responsibilities R (B,N,K) (zero on invalid reads)."""
        wd = self.wd
        Nk = R.sum(1) + wd(EPS10)                                  # (B,K)
        mu = np.einsum('bnk,bnd->bkd', R, Xz) / Nk[:, :, None]     # (B,K,D)
        d = Xz[:, :, None, :] - mu[:, None, :, :]                  # (B,N,K,D)
        cov = np.einsum('bnk,bnkd,bnke->bkde', R, d, d) / Nk[:, :, None, None]
        D = Xz.shape[2]
        cov = cov + wd(self.reg) * np.eye(D, dtype=wd)
        w = Nk / Nk.sum(1, keepdims=True)
        return w, mu, cov

    def _log_prob(self, Xz, w, mu, cov):
        """weighted log prob (B,N,K)."""
        D = Xz.shape[2]
        inv = np.linalg.inv(cov)                                   # (B,K,D,D)
        sign, logdet = np.linalg.slogdet(cov)                      # (B,K)
        d = Xz[:, :, None, :] - mu[:, None, :, :]
        maha = np.einsum('bnkd,bkde,bnke->bnk', d, inv, d)
        return (np.log(w)[:, None, :]
                - 0.5 * (D * LOG_2PI + logdet[:, None, :] + maha))

    def _e_step(self, Xz, valid, w, mu, cov):
        lp = self._log_prob(Xz, w, mu, cov)
        m = lp.max(2, keepdims=True)
        lse = (m + np.log(np.exp(lp - m).sum(2, keepdims=True)))[:, :, 0]
        R = np.exp(lp - lse[:, :, None]) * valid[:, :, None]
        n = valid.sum(1)
        with np.errstate(invalid='ignore', divide='ignore'):
            L = (lse * valid).sum(1) / n
        return L, R

    # -------------------------------------------------------------------- fit
    def fit(self, X):
        Xz, valid = self._prep(X)
        B, N, D = Xz.shape
        K = self.K
        n = valid.sum(1)
        self._ok = n >= 2
        vf = valid.astype(self.wd)
        with np.errstate(invalid='ignore', divide='ignore'):
            if K == 1:
                R = vf[:, :, None]
                w, mu, cov = self._moments(Xz, R)
                self.n_iter = np.zeros(B, dtype=np.int32)
            else:
                # deterministic 2-means initialisation (see the module docstring)
                ar = np.arange(B)
                mean = (Xz * vf[:, :, None]).sum(1) / np.maximum(n, 1)[:, None]
                dm = self._dist2(Xz, mean)
                c0 = np.where(valid, dm, -np.inf).argmax(1)         # farthest from the mean
                p0 = Xz[ar, c0]
                c1 = np.where(valid, self._dist2(Xz, p0), -np.inf).argmax(1)   # farthest from it
                p1 = Xz[ar, c1]
                assign = None
                self.n_kmeans = np.zeros(B, dtype=np.int32)
                live = self._ok.copy()
                for it in range(1, self.kmeans_max_iter + 1):
                    a0 = self._dist2(Xz, p0) <= self._dist2(Xz, p1)          # tie -> cluster 0
                    if assign is not None:
                        changed = ((a0 != assign) & valid).any(1)
                        live &= changed
                    assign = np.where(live[:, None], a0, assign) if it > 1 else a0
                    if not live.any():
                        break
                    self.n_kmeans[live] = it
                    w0 = (assign & valid).astype(self.wd)
                    w1 = (~assign & valid).astype(self.wd)
                    n0 = w0.sum(1)
                    n1 = w1.sum(1)
                    with np.errstate(invalid='ignore', divide='ignore'):
                        q0 = (Xz * w0[:, :, None]).sum(1) / n0[:, None]
                        q1 = (Xz * w1[:, :, None]).sum(1) / n1[:, None]
                    # an emptied cluster keeps its centre
                    q0 = np.where((n0 > 0)[:, None], q0, p0)
                    q1 = np.where((n1 > 0)[:, None], q1, p1)
                    p0 = np.where(live[:, None], q0, p0)
                    p1 = np.where(live[:, None], q1, p1)
                R = np.stack([assign, ~assign], axis=2).astype(self.wd) * vf[:, :, None]
                w, mu, cov = self._moments(Xz, R)
                # the parameters EM starts from (tests hand them to scikit-learn as *_init)
                self.init_ = (w.copy(), mu.copy(), cov.copy())

                active = self._ok.copy()
                prev = np.full(B, -np.inf)
                self.n_iter = np.zeros(B, dtype=np.int32)
                for it in range(1, self.max_iter + 1):
                    idx = np.nonzero(active)[0]
                    if idx.size == 0:
                        break
                    L, R = self._e_step(Xz[idx], valid[idx], w[idx], mu[idx], cov[idx])
                    w[idx], mu[idx], cov[idx] = self._moments(Xz[idx], R)
                    self.n_iter[idx] = it
                    done = np.abs(L - prev[idx]) < self.tol
                    prev[idx] = L
                    active[idx[done]] = False
        self._w, self._mu, self._cov = w, mu, cov
        return self

    # ------------------------------------------------------------- accessors
    @property
    def means(self):
        return [self._mu[:, k, :] for k in range(self.K)]

    @property
    def covs(self):
        return [self._cov[:, k, :, :] for k in range(self.K)]

    @property
    def weights(self):
        return [self._w[:, k] for k in range(self.K)]

    def loglik(self, X):
        Xz, valid = self._prep(X)
        with np.errstate(invalid='ignore', divide='ignore'):
            L, _ = self._e_step(Xz, valid, self._w, self._mu, self._cov)
        L = np.where(self._ok, L, np.nan)
        return L, valid.sum(1)

    def bic(self, X):
        L, n = self.loglik(X)
        D = self._mu.shape[2]
        K = self.K
        n_params = K * D + K * D * (D + 1) // 2 + K - 1
        with np.errstate(invalid='ignore', divide='ignore'):
            bic = -2.0 * n * L + n_params * np.log(n)
        return self._wrap(bic)

    def predict_proba(self, X, force_cpu_result=False):
        Xz, valid = self._prep(X)
        with np.errstate(invalid='ignore', divide='ignore'):
            _, R = self._e_step(Xz, valid, self._w, self._mu, self._cov)
        R = np.where(valid[:, :, None], R, np.nan)
        R[~self._ok] = np.nan
        return self._wrap(R)

    def _wrap(self, a):
        if self._torch_out:
            import torch
            return torch.from_numpy(np.ascontiguousarray(a))
        return a
