"""
KS / Mann-Whitney / Welch-t per position -- restatement of
/root/reference/nanocompore/comparisons.py:218-252 (``_nonparametric_test``).

The reference loops over positions in Python and hands two 1-D float32 torch
tensors to SciPy.  The oracle does the same loop and calls the same SciPy
functions, but on float64 copies (exact, order preserving): with the SciPy the
reference pins (1.15.2) float32 inputs gave float64 results; SciPy >= 1.16
preserves float32 and would round the p-values (SURVEY.md 8c "float32 trap").

Also restated here, for bit-exact checks of the integer statistics the CUDA
path exposes: ``mwu_stats`` (2*U1, tie term) following
scipy/stats/_mannwhitneyu.py:148-171, 479-537, and ``mwu_p``/``welch_p`` closed
forms used to pin the kernel's special functions.

Test infrastructure only (see oracle/__init__.py).
"""
import math
import warnings

import numpy as np
from scipy import stats as sps
from scipy.stats import mannwhitneyu, ttest_ind
from scipy.stats.mstats import ks_twosamp

INTENSITY = 0
DWELL = 1


def _stat_test(test):
    # comparisons.py:219-226
    if test in ("mann_whitney", "MW"):
        return lambda x, y: mannwhitneyu(x, y, alternative='two-sided')
    if test in ("kolmogorov_smirnov", "KS"):
        return ks_twosamp
    if test in ("t_test", "TT"):
        return lambda x, y: ttest_ind(x, y, equal_var=False)
    raise ValueError("Invalid statistical method name (MW, KS, TT)")


def nonparametric_test(test, test_data, conditions, log=None):
    """comparisons.py:218-252.  test_data (P,R,V) standardised float32."""
    stat_test = _stat_test(test)
    test_data = np.asarray(test_data)
    cond0 = np.asarray(conditions) == 0
    out = {f'{test}_intensity_pvalue': [], f'{test}_dwell_pvalue': []}
    for i in range(test_data.shape[0]):
        for var, name in ((INTENSITY, 'intensity'), (DWELL, 'dwell')):
            a = test_data[i, cond0, var]
            b = test_data[i, ~cond0, var]
            a = a[~np.isnan(a)].astype(np.float64)
            b = b[~np.isnan(b)].astype(np.float64)
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    p = float(stat_test(a, b).pvalue)
            except Exception as e:  # comparisons.py:239-250: swallowed -> NaN
                if log is not None:
                    log("error", f"Error in nonparametric test on the {name}: {e}")
                p = np.nan
            out[f'{test}_{name}_pvalue'].append(p)
    return {k: np.array(v, dtype=np.float64) for k, v in out.items()}


# ---------------------------------------------------------------------------
# integer statistics / closed forms
# ---------------------------------------------------------------------------

def mwu_stats(x0, x1):
    """
    Returns (U1x2, tie_term, n0, n1): twice the U statistic of sample 0 (an exact
    integer: U1 = R1 - n0(n0+1)/2 with average ranks) and sum(t^3 - t) over tie
    groups of the pooled sample.  _mannwhitneyu.py:489-495, 155-156.
    """
    x0 = np.asarray(x0, dtype=np.float64)
    x1 = np.asarray(x1, dtype=np.float64)
    n0, n1 = len(x0), len(x1)
    xy = np.concatenate([x0, x1])
    ranks = sps.rankdata(xy, method='average')
    r1x2 = int(round(2 * ranks[:n0].sum()))
    u1x2 = r1x2 - n0 * (n0 + 1)
    _, t = np.unique(xy, return_counts=True)
    t = t.astype(np.int64)
    tie = int((t ** 3 - t).sum())
    return u1x2, tie, n0, n1


def mwu_asymptotic_p(u1x2, tie, n0, n1):
    """_get_mwu_z + two-sided normal p (_mannwhitneyu.py:148-171, 520-522)."""
    n = n0 + n1
    mu = n0 * n1 / 2
    U1 = u1x2 / 2
    s = math.sqrt(n0 * n1 / 12 * ((n + 1) - tie / (n * (n - 1))))
    num = U1 - mu
    num -= 0.5 * np.sign(num)
    with np.errstate(divide='ignore', invalid='ignore'):
        z = np.float64(num) / np.float64(s)
    p = 2 * sps.norm.sf(abs(z))
    return float(np.clip(p, 0., 1.))


def welch_p(x0, x1):
    """ttest_ind(equal_var=False) two-sided p (scipy/stats/_stats_py.py:6369-6400)."""
    x0 = np.asarray(x0, dtype=np.float64)
    x1 = np.asarray(x1, dtype=np.float64)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return float(ttest_ind(x0, x1, equal_var=False).pvalue)
