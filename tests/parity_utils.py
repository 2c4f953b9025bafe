"""Shared helpers of the parity tests: tolerances and oracle-side per-position statistics."""
import numpy as np

from oracle import stats as ostats
from oracle.ks_exact import ks_h, ks_exact_p
from oracle.nonparametric import mwu_stats

# BASELINE.json north_star: "all p-values must match within a stated tolerance
# (1e-5 relative on log10 p)".  A relative bound on log10 p is empty at p ~ 1 (log10 p ~ 0),
# so an absolute bound of 1e-8 on log10 p (2.3e-8 relative on p) applies there.
LOG10P_RTOL = 1e-5
LOG10P_ATOL = 1e-8


def assert_pvalues_close(got, want, what=''):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert (nan_g == nan_w).all(), \
        f"{what}: NaN pattern differs at {np.nonzero(nan_g != nan_w)[0][:10]}"
    ok = ~nan_w
    g, w = got[ok], want[ok]
    # below 1e-300 the tails are denormal or flushed to 0 depending on the erfc at hand
    zero = (w < 1e-300) | (g < 1e-300)
    assert ((g < 1e-300) == (w < 1e-300))[zero].all() if zero.any() else True, \
        f"{what}: underflow pattern differs"
    g, w = g[~zero], w[~zero]
    lg, lw = np.log10(g), np.log10(w)
    bad = ~np.isclose(lg, lw, rtol=LOG10P_RTOL, atol=LOG10P_ATOL)
    if bad.any():
        i = np.nonzero(bad)[0][:10]
        raise AssertionError(f"{what}: {bad.sum()} of {bad.size} p-values off; first: got {g[i]}, want {w[i]}")


def assert_equal_nan(got, want, what=''):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    same = (got == want) | (np.isnan(got.astype(np.float64)) & np.isnan(want.astype(np.float64)))
    if not same.all():
        i = np.nonzero(~same)[0][:10]
        raise AssertionError(f"{what}: {(~same).sum()} of {same.size} differ; first at {i}: "
                             f"got {got[i]}, want {want[i]}")


def oracle_rank_stats(std_data, conditions):
    """Integer statistics per (position, variable) from the standardised data:
    dict of (P, 2) int64 arrays n0, n1, ks_h, mw_u2, mw_tie."""
    std_data = np.asarray(std_data)
    cond0 = np.asarray(conditions) == 0
    P = std_data.shape[0]
    out = {k: np.zeros((P, 2), dtype=np.int64) for k in ('n0', 'n1', 'ks_h', 'mw_u2', 'mw_tie')}
    for i in range(P):
        for v in (0, 1):
            a = std_data[i, cond0, v]
            b = std_data[i, ~cond0, v]
            a = a[~np.isnan(a)]
            b = b[~np.isnan(b)]
            out['n0'][i, v], out['n1'][i, v] = len(a), len(b)
            if len(a) and len(b):
                out['ks_h'][i, v] = ks_h(a, b)[0]
                u2, tie, _, _ = mwu_stats(a, b)
                out['mw_u2'][i, v], out['mw_tie'][i, v] = u2, tie
    return out
