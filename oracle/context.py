"""
Sequence-context p-value combination (Hou's method) -- restatement of
/root/reference/nanocompore/comparisons.py:
  * ``_combine_context_pvalues``  436-474
  * ``cross_corr_matrix``         627-651
  * ``harmomic_series``           654-656
  * ``combine_pvalues_hou``       659-705

Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np
from scipy.stats import chi2


def harmonic_series(context):
    return [1 / (abs(i) + 1) for i in range(-context, context + 1)]


def cross_corr_matrix(pvalues, context=2):
    pvalues = np.asarray(pvalues, dtype=float)
    if len(pvalues) < (3 * context) + 3:
        raise RuntimeError("Not enough p-values for a context of order %s" % context)
    pvalues = np.nan_to_num(pvalues, nan=1)
    if any(pvalues == 0) or any(np.isinf(pvalues)) or any(pvalues > 1):
        raise RuntimeError("At least one p-value is invalid")
    w = 2 * context + 1
    if all(pvalues == 1):
        return np.ones((w, w))
    m = np.zeros((w, w))
    s = pvalues.size
    for i in range(-context, context + 1):
        for j in range(-context, i + 1):
            a = np.roll(pvalues, i)[context:s - context]
            b = np.roll(pvalues, j)[context:s - context]
            m[i + context, j + context] = np.corrcoef(a, b)[0, 1]
    return m + np.tril(m, -1).T


def combine_pvalues_hou(pvalues, weights, cor_mat):
    if len(pvalues) != len(weights):
        raise ValueError("Can't combine pvalues if pvalues and weights are not the same length.")
    if cor_mat.shape[0] != cor_mat.shape[1] or cor_mat.shape[0] != len(pvalues):
        raise ValueError("The correlation matrix needs to be square")
    if all(p == 1 for p in pvalues):
        return 1
    if any(p == 0 or np.isinf(p) or p > 1 for p in pvalues):
        raise ValueError("At least one p-value is invalid")
    k = len(pvalues)
    cov_sum = 0.0
    sw_sum = 0.0
    w_sum = 0.0
    tau = 0.0
    for i in range(k):
        for j in range(i + 1, k):
            r = cor_mat[i][j]
            cov_sum += weights[i] * weights[j] * ((3.263 * r) + (0.710 * r ** 2) + (0.027 * r ** 3))
        sw_sum += weights[i] ** 2
        w_sum += weights[i]
        tau += weights[i] * (-2 * np.log(pvalues[i]))
    c = (2 * sw_sum + cov_sum) / (2 * w_sum)
    f = (4 * w_sum ** 2) / (2 * sw_sum + cov_sum)
    p = chi2.sf(tau / c, f)
    if p == 0:
        p = np.finfo(np.float64).tiny
    return p


def combine_context_pvalues(pos, pvals, context, weights_mode='uniform'):
    """comparisons.py:436-474.  pos, pvals: per result row."""
    pos = np.asarray(pos).astype(np.int64)
    pvals = np.asarray(pvals, dtype=float)
    weights = harmonic_series(context) if weights_mode == 'harmonic' else [1] * (2 * context + 1)
    min_pos, max_pos = pos.min(), pos.max()
    padded = max_pos + 1 - min_pos + 2 * context
    pv = np.ones(padded, dtype=float)
    idx = pos - min_pos
    pv[idx + context] = pvals
    corr = cross_corr_matrix(pv[context:-context], context)
    win = 2 * context + 1
    out = np.empty(padded - win + 1)
    for s in range(out.size):
        w = pv[s:s + win]
        if np.isnan(w[context]):
            out[s] = np.nan
        else:
            out[s] = combine_pvalues_hou(np.nan_to_num(w, nan=1), weights, corr)
    return out[idx]
