"""
Benjamini-Hochberg FDR with NaN skipping -- restatement of
/root/reference/nanocompore/postprocessing.py:209-285 (``_correct_pvalues``,
``_multipletests_filter_nan``) and of ``statsmodels.stats.multitest.multipletests
(method='fdr_bh')`` (statsmodels is not installed here; its published algorithm is
``fdrcorrection``: sort ascending, q_i = p_i * n / rank_i, running minimum from the
largest p downwards, clip at 1, un-sort).  Pinned by the known answer in the
reference's docstring / tests/test_postprocessing.py:18-24.

Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np


def fdr_bh(p):
    p = np.asarray(p, dtype=np.float64)
    n = p.size
    if n == 0:
        return p.copy()
    order = np.argsort(p, kind='stable')
    ps = p[order]
    q = ps * n / np.arange(1, n + 1)
    q = np.minimum.accumulate(q[::-1])[::-1]
    q = np.minimum(q, 1.0)
    out = np.empty(n)
    out[order] = q
    return out


def fdr_bh_segment(p, above=0, total=None):
    """BH q-values of the numbers in ``p`` (NaNs skipped) when they are one value range of a column of
    ``total`` numbers, ``above`` of them larger than all of ``p``: the running minimum stops at the
    range's own largest value (the caller takes the minimum with the ranges above).  With the
    defaults this is ``multipletests_filter_nan``."""
    p = np.asarray(p, dtype=np.float64)
    nans = np.isnan(p)
    v = p[~nans]
    m = v.size if total is None else int(total)
    out = p.copy()
    if v.size == 0:
        return out
    order = np.argsort(v, kind='stable')
    ranks = m - above - np.arange(v.size - 1, -1, -1)          # ascending ranks of the sorted values
    q = v[order] / (ranks / m)
    q = np.minimum.accumulate(q[::-1])[::-1]
    q = np.minimum(q, 1.0)
    res = np.empty(v.size)
    res[order] = q
    out[~nans] = res
    return out


def multipletests_filter_nan(pvalues):
    pvalues = np.asarray(pvalues, dtype=np.float64)
    nans = np.isnan(pvalues)
    q = pvalues.copy()
    q[~nans] = fdr_bh(pvalues[~nans])
    return q


def auto_qvalues(pvalues, auto_test):
    """postprocessing.py:211-235: BH per test type, scaled by N/n."""
    pvalues = np.asarray(pvalues, dtype=np.float64)
    auto_test = np.asarray(auto_test)
    allq = np.full(pvalues.shape, np.nan)
    for t in np.unique(auto_test):
        m = auto_test == t
        q = multipletests_filter_nan(pvalues[m])
        if len(q) > 0:
            q = q * (len(pvalues) / len(q))
        allq[m] = q
    return allq
