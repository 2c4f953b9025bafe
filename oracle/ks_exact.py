"""
Two-sample two-sided KS: integer statistic and exact p-value -- restatement.

Follows SciPy (the reference calls ``scipy.stats.mstats.ks_twosamp`` at
/root/reference/nanocompore/comparisons.py:222, 236-245):
  * ``ks_2samp``                          scipy/stats/_stats_py.py:7879-8133
  * ``_attempt_exact_2kssamp``            scipy/stats/_stats_py.py:7825-7870
  * ``_compute_prob_outside_square``      scipy/stats/_stats_py.py:7713-7748
  * ``_compute_outer_prob_inside_method`` compiled (pythran) in scipy; its published
    algorithm (scipy/stats/_stats_pythran.py in the SciPy source tree) is restated here
    and checked bitwise against the compiled function in tests/test_oracle_ks.py.

Test infrastructure only (see oracle/__init__.py).
"""
import math

import numpy as np


def ks_h(x0, x1):
    """
    Integer KS statistic.  With g = gcd(n0, n1), returns (h, n0, n1) where
        h = max over distinct pooled values v of | c0(v)*(n1/g) - c1(v)*(n0/g) |,
    c_k(v) = #{x_k <= v}.  h / lcm(n0, n1) is the ``d`` SciPy's exact routine
    consumes (``h = int(round(d * lcm))``, _stats_py.py:7834-7836).
    """
    x0 = np.sort(np.asarray(x0, dtype=np.float64))
    x1 = np.sort(np.asarray(x1, dtype=np.float64))
    n0, n1 = len(x0), len(x1)
    if n0 == 0 or n1 == 0:
        return 0, n0, n1
    g = math.gcd(n0, n1)
    allv = np.concatenate([x0, x1])
    c0 = np.searchsorted(x0, allv, side='right').astype(np.int64)
    c1 = np.searchsorted(x1, allv, side='right').astype(np.int64)
    h = int(np.max(np.abs(c0 * (n1 // g) - c1 * (n0 // g))))
    return h, n0, n1


def prob_outside_square(n, h):
    """_stats_py.py:7713-7748 (equal sample sizes, Horner product)."""
    P = 0.0
    k = int(math.floor(n / h))
    while k >= 0:
        p1 = 1.0
        for j in range(h):
            p1 = (n - k * h - j) * p1 / (n + k * h + j + 1)
        P = p1 * (1.0 - P)
        k -= 1
    return 2 * P


def outer_prob_inside_method(m, n, g, h):
    """
    Proportion of lattice paths (0,0)->(m,n) that do NOT stay strictly inside
    |x/m - y/n| < h/lcm.  Sliding-window recurrence
        A(i,j) = (i*A(i-1,j) + j*A(i,j-1)) / (i+j),  A = 1 outside the band.
    """
    if m < n:
        m, n = n, m
    mg = m // g
    ng = n // g
    minj, maxj = 0, min(int(math.ceil(h / mg)), n + 1)
    curlen = maxj - minj
    lenA = min(2 * maxj + 2, n + 1)
    A = np.ones(lenA, dtype=np.float64)
    A[minj:maxj] = 0.0
    for i in range(1, m + 1):
        lastminj, lastlen = minj, curlen
        minj = max(int(math.floor((ng * i - h) / mg)) + 1, 0)
        minj = min(minj, n)
        maxj = min(int(math.ceil((ng * i + h) / mg)), n + 1)
        if maxj <= minj:
            return 1.0
        val = 0.0 if minj == 0 else 1.0
        for jj in range(maxj - minj):
            j = jj + minj
            val = (A[jj + minj - lastminj] * i + val * j) / (i + j)
            A[jj] = val
        curlen = maxj - minj
        if lastlen > curlen:
            A[maxj - minj:maxj - minj + (lastlen - curlen)] = 1
    return float(A[maxj - minj - 1])


def ks_exact_p(n0, n1, h):
    """
    _attempt_exact_2kssamp + the clip of _ks_2samp_prob, two-sided, given the
    integer statistic.  Valid for max(n0, n1) <= 10000 (``MAX_AUTO_N``); beyond that
    SciPy switches to the asymptotic formula, which this path never reaches
    (downsample_high_coverage caps reads at 5000 per condition).
    """
    if n0 == 0 or n1 == 0:
        return float('nan')
    if h == 0:
        return 1.0
    g = math.gcd(n0, n1)
    if n0 == n1:
        prob = prob_outside_square(n0, h)
    else:
        prob = outer_prob_inside_method(n0, n1, g, h)
    if not (0 <= prob <= 1):
        # _attempt_exact_2kssamp reports failure (rounding put the Horner product just above 1)
        # and ks_2samp falls back to the asymptotic branch (_stats_py.py:8156-8171)
        from scipy.stats import distributions
        m, n = sorted([float(n0), float(n1)], reverse=True)
        en = m * n / (m + n)
        d = h * 1.0 / ((n0 // g) * n1)
        prob = float(distributions.kstwo.sf(d, np.round(en)))
    return float(min(max(prob, 0.0), 1.0))
