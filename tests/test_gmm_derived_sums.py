"""
CPU model of a design decision of the mixture kernel (position_kernel.cu, NCB_GMM_DERIVE): per read
only ONE component's sufficient statistics are accumulated (the one with the smaller N_k in the pass
before; 2-means cluster 0 in the Lloyd iterations), the other's are the totals of the position minus
the first's.  The model below runs the kernel's formulation of the fit in numpy fp64 -- raw-moment
M-step, responsibilities as 1 / (1 + e^-|d|) and its complement, log-likelihood as max + log(1 + e),
derived sums, optionally the folded products of the A/B macro NCB_ACC_FOLD -- and must agree with
the oracle's definition (oracle/gmm.py: centred moments, softmax responsibilities, every sum taken directly) on the number of Lloyd and EM iterations, the hard
assignment of every read and the BIC, on ordinary, tiny-coverage, quantised and lopsided positions.
It pins the numerics of the formulation (cancellation in the derived sums), not the kernel's bits:
those are the GPU parity tests' and sweeps' job.
"""
import numpy as np
import pytest

from oracle.gmm import GMM, EPS10, LOG_2PI


def kernel_form_fit(X, derive=True, fold=False, reg=1e-6, tol=1e-3, max_iter=100, kmeans_max_iter=20):
    """X: (B, N, 2) with NaN rows.  Returns dict(n_kmeans, n_iter, pred (B,N) int8 with -1 on
    missing reads, loglik (B,))."""
    X = np.asarray(X, dtype=np.float64)
    valid = ~np.isnan(X).any(2)
    Z = np.where(valid[:, :, None], X, 0.0)
    vf = valid.astype(np.float64)
    x, y = Z[:, :, 0], Z[:, :, 1]
    feats = np.stack([vf, x * vf, y * vf, x * x * vf, x * y * vf, y * y * vf], 2)      # (B,N,6)
    T = feats.sum(1)                                                               # totals (B,6)
    n = T[:, 0]
    B = X.shape[0]

    def dist2(cx, cy):
        dx, dy = x - cx[:, None], y - cy[:, None]
        return dx * dx + dy * dy

    # farthest pair
    with np.errstate(invalid='ignore', divide='ignore'):
        c0 = np.where(valid, dist2(T[:, 1] / n, T[:, 2] / n), -np.inf).argmax(1)
    ar = np.arange(B)
    p0x, p0y = x[ar, c0], y[ar, c0]
    c1 = np.where(valid, dist2(p0x, p0y), -np.inf).argmax(1)
    p1x, p1y = x[ar, c1], y[ar, c1]
    # Lloyd
    assign = np.zeros_like(valid)
    live = n >= 2
    n_kmeans = np.zeros(B, dtype=np.int32)
    for it in range(1, kmeans_max_iter + 1):
        a0 = dist2(p0x, p0y) <= dist2(p1x, p1y)
        if it > 1:
            live &= ((a0 != assign) & valid).any(1)
            if not live.any():
                break
        assign = np.where(live[:, None], a0, assign) if it > 1 else a0
        n_kmeans[live] = it
        w0 = (assign & valid).astype(np.float64)
        n0, sx0, sy0 = w0.sum(1), (w0 * x).sum(1), (w0 * y).sum(1)
        if derive:
            n1, sx1, sy1 = n - n0, T[:, 1] - sx0, T[:, 2] - sy0
        else:
            w1 = (~assign & valid).astype(np.float64)
            n1, sx1, sy1 = w1.sum(1), (w1 * x).sum(1), (w1 * y).sum(1)
        with np.errstate(invalid='ignore', divide='ignore'):
            q0x, q0y = np.where(n0 > 0, sx0 / n0, p0x), np.where(n0 > 0, sy0 / n0, p0y)
            q1x, q1y = np.where(n1 > 0, sx1 / n1, p1x), np.where(n1 > 0, sy1 / n1, p1y)
        p0x, p0y = np.where(live, q0x, p0x), np.where(live, q0y, p0y)
        p1x, p1y = np.where(live, q1x, p1x), np.where(live, q1y, p1y)

    def m_step(s):           # s (B,6) -> parameters of the E-step (position_kernel.cu: m_step, comp_e)
        nk = s[:, 0] + EPS10
        mx, my = s[:, 1] / nk, s[:, 2] / nk
        axx = s[:, 3] - 2.0 * mx * s[:, 1] + s[:, 0] * mx * mx
        ayy = s[:, 5] - 2.0 * my * s[:, 2] + s[:, 0] * my * my
        axy = s[:, 4] - mx * s[:, 2] - my * s[:, 1] + s[:, 0] * mx * my
        return nk, mx, my, axx / nk + reg, axy / nk, ayy / nk + reg

    def log_prob(par, w):
        nk, mx, my, cxx, cxy, cyy = par
        det = cxx * cyy - cxy * cxy
        dx, dy = x - mx[:, None], y - my[:, None]
        cst = 0.5 * np.log(w * w / det) - LOG_2PI
        return (cst[:, None] - 0.5 * (cyy / det)[:, None] * dx * dx + (cxy / det)[:, None] * dx * dy
                - 0.5 * (cxx / det)[:, None] * dy * dy)

    def e_step(p0, p1):
        tot = p0[0] + p1[0]
        l0, l1 = log_prob(p0, p0[0] / tot), log_prob(p1, p1[0] / tot)
        d = l0 - l1
        ex = np.exp(np.maximum(-np.abs(d), -700.0))
        rbig = 1.0 / (1.0 + ex)
        r0 = np.where(d >= 0, rbig, ex * rbig)
        r1 = np.where(d >= 0, ex * rbig, rbig)
        ll = ((np.maximum(l0, l1) + np.log1p(ex)) * vf).sum(1)
        return r0 * vf, r1 * vf, ll, d

    def weighted(r):
        if not fold:
            return (feats * r[:, :, None]).sum(1)
        rx, ry = r * x, r * y          # NCB_ACC_FOLD: second moments as (r x) x
        return np.stack([r, rx, ry, rx * x, rx * y, ry * y], 2).sum(1)

    def stats(r0, r1, second):
        """sufficient statistics of both components; `second` (B,) bool: the component summed."""
        if not derive:
            return weighted(r0), weighted(r1)
        rs = np.where(second[:, None], r1, r0)
        small = weighted(rs)
        rest = T - small
        rest[:, 0] = np.maximum(rest[:, 0], 0.0)
        sec = second[:, None]
        return np.where(sec, rest, small), np.where(sec, small, rest)

    with np.errstate(invalid='ignore', divide='ignore'):
        a_f = (assign & valid).astype(np.float64)
        s0, s1 = stats(a_f, vf - a_f, np.zeros(B, dtype=bool))
        p0, p1 = m_step(s0), m_step(s1)
        active = n >= 2
        prev = np.full(B, -np.inf)
        n_iter = np.zeros(B, dtype=np.int32)
        for it in range(1, max_iter + 1):
            if not active.any():
                break
            second = s1[:, 0] < s0[:, 0]
            r0, r1, ll, _ = e_step(p0, p1)
            t0, t1 = stats(r0, r1, second)
            q0, q1 = m_step(t0), m_step(t1)
            upd = active
            s0, s1 = np.where(upd[:, None], t0, s0), np.where(upd[:, None], t1, s1)
            p0 = tuple(np.where(upd, a, b) for a, b in zip(q0, p0))
            p1 = tuple(np.where(upd, a, b) for a, b in zip(q1, p1))
            n_iter[upd] = it
            Lm = ll / n
            done = np.abs(Lm - prev) < tol
            prev = np.where(upd, Lm, prev)
            active = active & ~done
        _, _, ll, d = e_step(p0, p1)
    pred = np.where(valid, (d < 0).astype(np.int8), -1)
    return dict(n_kmeans=n_kmeans, n_iter=n_iter, pred=pred, loglik=ll / n)


def positions(kind, rng, B, N):
    X = np.full((B, N, 2), np.nan)
    for b in range(B):
        if kind == 'ordinary':
            m = rng.integers(max(4, N // 3), N + 1)
            pts = rng.standard_normal((m, 2))
            k = rng.integers(0, m)
            pts[:k] += rng.normal(0, 3, 2)
        elif kind == 'tiny':
            m = rng.integers(3, 13)
            pts = rng.standard_normal((m, 2))
        elif kind == 'quantised':
            m = rng.integers(20, N + 1)
            pts = np.round(rng.standard_normal((m, 2)) * 4) / 4
            pts[: m // 3, 0] += 2.0
        else:   # lopsided: a tight group of 1-3 reads far from everything else
            m = rng.integers(30, N + 1)
            pts = rng.standard_normal((m, 2))
            k = rng.integers(1, 4)
            pts[:k] = rng.normal(0, 0.01, (k, 2)) + rng.choice([-1, 1], 2) * rng.uniform(6, 40)
        pts = (pts - pts.mean(0)) / pts.std(0)            # standardised, like the kernel's input
        X[b, rng.permutation(N)[:m]] = pts
    return X


@pytest.mark.parametrize('kind,B,N', [('ordinary', 1500, 200), ('tiny', 1500, 12), ('quantised', 800, 120),
                                      ('lopsided', 800, 150)])
def test_derived_sums_agree_with_the_definition(kind, B, N):
    rng = np.random.default_rng({'ordinary': 1, 'tiny': 2, 'quantised': 3, 'lopsided': 4}[kind])
    X = positions(kind, rng, B, N)
    ref = GMM(2).fit(X)
    want_pred = np.where(~np.isnan(X).any(2), ref.predict_proba(X).argmax(2), -1)
    want_ll, _ = ref.loglik(X)
    for derive, fold in ((False, False), (True, False), (True, True)):
        got = kernel_form_fit(X, derive=derive, fold=fold)
        np.testing.assert_array_equal(got['n_kmeans'], ref.n_kmeans, err_msg=f'{kind} Lloyd iterations, derive={derive}')
        np.testing.assert_array_equal(got['n_iter'], ref.n_iter, err_msg=f'{kind} EM iterations, derive={derive}')
        np.testing.assert_array_equal(got['pred'], want_pred, err_msg=f'{kind} hard assignment, derive={derive}')
        np.testing.assert_allclose(got['loglik'], want_ll, rtol=1e-9, atol=1e-9, err_msg=f'{kind} log-likelihood')
