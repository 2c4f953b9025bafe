"""
GMM test on cluster assignment vs condition -- numpy restatement of
/root/reference/nanocompore/comparisons.py:
  * ``_gmm_test_split``                338-392
  * ``get_contingency_matrices``       714-729
  * ``get_soft_contingency_matrices``  732-744
  * ``_get_mod_cluster``               564-605
  * ``_get_cluster_counts``            477-517
  * ``_get_soft_cluster_counts``       520-561
  * ``calculate_lors``                 708-711
and of ``scipy.stats.chi2_contingency`` for 2x2 tables
(scipy/stats/contingency.py:317-356: Yates correction, Pearson statistic, dof 1).

The optional logit columns (``GMM_logit_pvalue`` / ``GMM_logit_coef``) have NO
reference code in nanocompore v2 -- parity unpinned; defined
here as the Wald test of the slope of a logistic regression condition ~ cluster on
the +1 table, fitted by IRLS.

Test infrastructure only (see oracle/__init__.py).
"""
import math

import numpy as np
from scipy.stats import chi2_contingency, norm

from oracle.gmm import GMM

F32 = np.float32


def contingency_matrices(conditions, pred):
    """comparisons.py:714-729.  pred (P,R) float with NaN for missing reads."""
    conditions = np.asarray(conditions)
    pred = np.asarray(pred)
    P = pred.shape[0]
    cont = np.zeros((P, 2, 2), dtype=F32)
    for c in (0, 1):
        for k in (0, 1):
            cont[:, c, k] = ((conditions[None, :] == c) & (pred == k)).sum(1)
    return cont


def soft_contingency_matrices(conditions, probs):
    """comparisons.py:732-744."""
    conditions = np.asarray(conditions)
    probs = np.asarray(probs)
    P = probs.shape[0]
    cont = np.zeros((P, 2, 2), dtype=F32)
    for c in (0, 1):
        cont[:, c, :] = np.nansum(probs[:, conditions == c, :].astype(F32), axis=1,
                                  dtype=np.float64).astype(F32)
    return cont


def mod_cluster(cont, depleted_cond_id):
    """comparisons.py:564-605 (argmin of the depleted condition's cluster frequencies)."""
    cond_counts = cont.sum(2).astype(np.uint32)
    with np.errstate(invalid='ignore', divide='ignore'):
        freq = (cont / cond_counts[:, :, None].astype(F32)).astype(F32)
    row = freq[:, depleted_cond_id, :]
    # torch.argmin treats NaN as the minimum (first NaN wins); np.argmin does the same.
    return np.argmin(row, axis=1)


def cluster_counts(cont, samples, pred, sample_ids, depleted_cond_id):
    """comparisons.py:477-517.  sample_ids: dict label -> id."""
    samples = np.asarray(samples)
    mod = mod_cluster(cont, depleted_cond_id)[:, None]
    out = {}
    for label, sid in sample_ids.items():
        sp = pred[:, samples == sid]
        totals = (~np.isnan(sp)).sum(1)
        modc = (sp == mod).sum(1)
        out[f'{label}_mod'] = modc.astype(np.uint32)
        out[f'{label}_unmod'] = (totals - modc).astype(np.uint32)
    return out


def soft_cluster_counts(cont, samples, probs, sample_ids, depleted_cond_id):
    """comparisons.py:520-561."""
    samples = np.asarray(samples)
    mod = mod_cluster(cont, depleted_cond_id)
    B = probs.shape[0]
    probs = probs.astype(F32)
    modp = probs[np.arange(B), :, mod]
    unmodp = probs[np.arange(B), :, 1 - mod]
    out = {}
    for label, sid in sample_ids.items():
        m = samples == sid
        out[f'{label}_mod'] = np.nansum(modp[:, m], axis=1, dtype=np.float64).astype(F32)
        out[f'{label}_unmod'] = np.nansum(unmodp[:, m], axis=1, dtype=np.float64).astype(F32)
    return out


def calculate_lors(cont):
    """comparisons.py:708-711: float32, rounded to 3 decimals (torch.round(decimals=3))."""
    cont = cont.astype(F32)
    with np.errstate(invalid='ignore', divide='ignore'):
        odds1 = cont[:, 0, 0] / cont[:, 0, 1]
        odds2 = cont[:, 1, 0] / cont[:, 1, 1]
        lor = np.log((odds1 / odds2).astype(F32)).astype(F32)
    return round3_f32(lor)


def round3_f32(x):
    """torch.round(x, decimals=3) on float32: rint(x * 1000) / 1000 in float32
    (ATen RoundDecimals: x * ten_pow, nearbyint, / ten_pow)."""
    x = np.asarray(x, dtype=F32)
    with np.errstate(invalid='ignore'):
        return (np.rint((x * F32(1000.0)).astype(F32)) / F32(1000.0)).astype(F32)


def chi2_2x2_p(table):
    """chi2_contingency(table).pvalue for a 2x2 table of positive counts, restated:
    expected = row*col/total; Yates: move observed 0.5 (at most |E-O|) towards expected;
    Pearson statistic; p = chi2.sf(stat, 1)."""
    obs = np.asarray(table, dtype=np.float64)
    total = obs.sum()
    exp = np.outer(obs.sum(1), obs.sum(0)) / total
    diff = exp - obs
    direction = np.sign(diff)
    magnitude = np.minimum(0.5, np.abs(diff))
    obs = obs + magnitude * direction
    stat = ((obs - exp) ** 2 / exp).sum()
    return math.erfc(math.sqrt(stat / 2.0))


def logit_wald(table, max_iter=25, tol=1e-10):
    """
    Logistic regression  condition ~ b0 + b1*cluster  on a 2x2 table of weights
    (rows condition, columns cluster), Newton/IRLS from b = 0.  Returns (coef b1,
    two-sided Wald p).  Closed form for checking: b1 = ln(t11*t00/(t10*t01)),
    se^2 = sum 1/t.
    """
    t = np.asarray(table, dtype=np.float64)
    b0 = b1 = 0.0
    # covariate patterns: cluster x in {0,1}; successes = condition 1
    n = t.sum(0)           # per cluster totals
    y = t[1, :]            # condition-1 counts per cluster
    for _ in range(max_iter):
        eta = np.array([b0, b0 + b1])
        mu = 1.0 / (1.0 + np.exp(-eta))
        wgt = n * mu * (1 - mu)
        g0 = (y - n * mu).sum()
        g1 = (y - n * mu)[1]
        h00 = wgt.sum()
        h01 = wgt[1]
        h11 = wgt[1]
        det = h00 * h11 - h01 * h01
        d0 = (h11 * g0 - h01 * g1) / det
        d1 = (-h01 * g0 + h00 * g1) / det
        b0 += d0
        b1 += d1
        if abs(d0) + abs(d1) < tol:
            break
    eta = np.array([b0, b0 + b1])
    mu = 1.0 / (1.0 + np.exp(-eta))
    wgt = n * mu * (1 - mu)
    det = wgt.sum() * wgt[1] - wgt[1] ** 2
    var_b1 = wgt.sum() / det
    z = b1 / math.sqrt(var_b1)
    return b1, float(2 * norm.sf(abs(z)))


def gmm_test_split(data, samples, conditions, sample_ids, depleted_cond_id,
                   cluster_counts_mode='hard', gmm_cls=GMM, with_logit=False,
                   return_fit=False):
    """comparisons.py:338-392.  data (P,R,D) standardised values (NaN = missing)."""
    data = np.asarray(data, dtype=F32)       # test_data = data.to(self._dtype)
    g1 = gmm_cls(n_components=1).fit(data)
    bic1 = np.asarray(g1.bic(data))
    g2 = gmm_cls(n_components=2).fit(data)
    probs = np.asarray(g2.predict_proba(data))
    nanrow = np.isnan(probs).any(2)
    pred = np.where(nanrow, np.nan, np.argmax(np.nan_to_num(probs, nan=-1.0), axis=2))
    bic2 = np.asarray(g2.bic(data))

    cont = contingency_matrices(conditions, pred)
    if cluster_counts_mode == 'hard':
        counts = cluster_counts(cont, samples, pred, sample_ids, depleted_cond_id)
    elif cluster_counts_mode == 'soft':
        scont = soft_contingency_matrices(conditions, probs)
        counts = soft_cluster_counts(scont, samples, probs, sample_ids, depleted_cond_id)
    else:
        counts = {}
    cont1 = cont + 1
    lors = calculate_lors(cont1)
    pvals = np.array([chi2_contingency(c).pvalue for c in cont1], dtype=np.float64)
    with np.errstate(invalid='ignore'):
        ignored = ~(bic2 < bic1)          # bic1 <= bic2, and NaN BIC -> ignored
    ignored_ref = (bic1 <= bic2)
    # a NaN BIC (fewer than 2 valid reads) cannot occur in the reference's flow
    # (coverage filter); the oracle defines it as "not tested".
    ignored = ignored_ref | np.isnan(bic1) | np.isnan(bic2)
    lors[ignored] = np.nan
    pvals[ignored] = np.nan
    out = {'GMM_chi2_pvalue': pvals, 'GMM_LOR': lors}
    if with_logit:
        coef = np.full(len(cont1), np.nan)
        lp = np.full(len(cont1), np.nan)
        for i, c in enumerate(cont1):
            if not ignored[i]:
                coef[i], lp[i] = logit_wald(c)
        out['GMM_logit_pvalue'] = lp
        out['GMM_logit_coef'] = coef
    out |= counts
    if return_fit:
        return out, dict(bic1=bic1, bic2=bic2, pred=pred, probs=probs, cont=cont,
                         gmm1=g1, gmm2=g2)
    return out
