"""
The scalar fp64 routines the kernels call once per position (nanocompore_b200/csrc/ncb_math.cuh),
compiled for the host by __graft_entry__.build() and checked against SciPy on the CPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'build', 'host_check.so')


@pytest.fixture(scope='module')
def hc():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(ROOT, 'tests', 'csrc', 'host_check.cpp')
    hdr = os.path.join(ROOT, 'nanocompore_b200', 'csrc', 'ncb_math.cuh')
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-ffp-contract=off',
                        '-I', os.path.join(ROOT, 'nanocompore_b200', 'csrc'), src, '-o', SO], check=True)
    lib = C.CDLL(SO)
    d, ll = C.c_double, C.c_longlong
    lib.hc_student_t_two_sided.restype = d; lib.hc_student_t_two_sided.argtypes = [d, d]
    lib.hc_welch_p.restype = d; lib.hc_welch_p.argtypes = [d] * 6
    lib.hc_mwu_asymptotic_p.restype = d; lib.hc_mwu_asymptotic_p.argtypes = [ll] * 4
    lib.hc_mwu_exact_p.restype = d; lib.hc_mwu_exact_p.argtypes = [ll] * 3
    lib.hc_chi2_2x2_p.restype = d; lib.hc_chi2_2x2_p.argtypes = [d] * 4
    lib.hc_logit_wald.restype = None
    lib.hc_logit_wald.argtypes = [d] * 4 + [C.POINTER(d), C.POINTER(d)]
    lib.hc_ks_prob_outside_square.restype = d; lib.hc_ks_prob_outside_square.argtypes = [ll, ll]
    lib.hc_ks_equal_p.restype = d; lib.hc_ks_equal_p.argtypes = [ll, ll]
    lib.hc_normal_two_sided.restype = d; lib.hc_normal_two_sided.argtypes = [d]
    lib.hc_igamc.restype = d; lib.hc_igamc.argtypes = [d, d]
    lib.hc_quot_rn.restype = None; lib.hc_quot_rn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, ll]
    return lib


def close_log10(got, want, rtol=1e-5, atol=1e-8):
    if want == 0 or got == 0:
        return want == got or max(want, got) < 1e-300
    return abs(np.log10(got) - np.log10(want)) <= atol + rtol * abs(np.log10(want))


def test_student_t(hc):
    for df in [1, 1.7, 2.5, 5, 17.3, 60, 99.7, 198, 500, 5000, 19998]:
        for t in [0, 1e-3, 0.1, 0.5, 1, 2, 3, 5, 8, 12, 20, 40, 80, 200, 1000]:
            got, want = hc.hc_student_t_two_sided(t, df), 2 * stats.t.sf(t, df)
            assert close_log10(got, want), (df, t, got, want)
            assert hc.hc_student_t_two_sided(-t, df) == got


def test_welch_against_scipy(hc):
    rng = np.random.default_rng(0)
    for _ in range(200):
        n0, n1 = rng.integers(2, 300, size=2)
        a = rng.standard_normal(n0) * rng.uniform(0.5, 2) + rng.uniform(-1, 1)
        b = rng.standard_normal(n1) * rng.uniform(0.5, 2)
        want = stats.ttest_ind(a, b, equal_var=False).pvalue
        got = hc.hc_welch_p(a.mean(), a.var(ddof=1), n0, b.mean(), b.var(ddof=1), n1)
        assert close_log10(got, want), (n0, n1, got, want)


def test_mwu_asymptotic_against_scipy(hc):
    from oracle.nonparametric import mwu_stats
    rng = np.random.default_rng(1)
    for _ in range(200):
        n0, n1 = rng.integers(9, 200, size=2)
        a = np.round(rng.standard_normal(n0) * 4) / 4
        b = np.round(rng.standard_normal(n1) * 4 + rng.uniform(-2, 2)) / 4
        u2, tie, _, _ = mwu_stats(a, b)
        want = stats.mannwhitneyu(a, b, alternative='two-sided').pvalue
        assert close_log10(hc.hc_mwu_asymptotic_p(u2, tie, n0, n1), want)


def test_mwu_exact_against_scipy(hc):
    rng = np.random.default_rng(2)
    for _ in range(300):
        n0 = int(rng.integers(1, 9))
        n1 = int(rng.integers(1, 40))
        if rng.random() < 0.5:
            n0, n1 = n1, n0
        a = rng.standard_normal(n0)
        b = rng.standard_normal(n1) + rng.uniform(-1, 1)
        r = stats.mannwhitneyu(a, b, alternative='two-sided', method='exact')
        u1 = int(r.statistic)
        umin = min(u1, n0 * n1 - u1)
        got = hc.hc_mwu_exact_p(umin, n0, n1)
        assert close_log10(got, r.pvalue), (n0, n1, u1, got, r.pvalue)
        # method='auto' picks the exact distribution here (no ties, a sample of <= 8 values)
        assert stats.mannwhitneyu(a, b, alternative='two-sided').pvalue == r.pvalue


def test_chi2_2x2_against_scipy(hc):
    rng = np.random.default_rng(3)
    for _ in range(300):
        t = rng.integers(1, 300, size=(2, 2)).astype(float)
        want = stats.chi2_contingency(t).pvalue
        got = hc.hc_chi2_2x2_p(t[0, 0], t[0, 1], t[1, 0], t[1, 1])
        assert close_log10(got, want), (t, got, want)


def test_logit_wald_closed_form(hc):
    """On a 2x2 table the logistic MLE is closed form: coef = ln(ad/bc), se^2 = sum 1/t."""
    rng = np.random.default_rng(4)
    for _ in range(200):
        t = rng.integers(1, 400, size=(2, 2)).astype(float)
        coef, p = C.c_double(), C.c_double()
        hc.hc_logit_wald(t[0, 0], t[0, 1], t[1, 0], t[1, 1], C.byref(coef), C.byref(p))
        want_coef = np.log(t[1, 1] * t[0, 0] / (t[1, 0] * t[0, 1]))
        se = np.sqrt((1 / t).sum())
        want_p = 2 * stats.norm.sf(abs(want_coef / se))
        assert abs(coef.value - want_coef) < 1e-8 * max(1, abs(want_coef))
        assert close_log10(p.value, want_p, rtol=1e-5, atol=1e-7)
        from oracle.gmm_test import logit_wald
        oc, op = logit_wald(t)
        assert abs(oc - coef.value) < 1e-12 * max(1, abs(oc)) and close_log10(p.value, op)


def test_ks_equal_sizes_bitwise(hc):
    from scipy.stats._stats_py import _compute_prob_outside_square
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(1, 3000))
        h = int(rng.integers(1, n + 1))
        assert hc.hc_ks_prob_outside_square(n, h) == _compute_prob_outside_square(n, h)


def test_ks_equal_sizes_with_scipy_fallback(hc):
    """Where SciPy's exact attempt 'fails' (Horner product rounds above 1) it reports the
    one-sample Kolmogorov tail instead; ks_equal_p follows it to 4e-13."""
    from oracle.ks_exact import ks_exact_p
    hit = 0
    for n in list(range(1, 420)) + [777, 1000, 2500, 5000]:
        for h in range(1, min(n, 12) + 1):
            want = ks_exact_p(n, n, h)
            got = hc.hc_ks_equal_p(n, h)
            from scipy.stats._stats_py import _compute_prob_outside_square
            raw = _compute_prob_outside_square(n, h)
            if not (0 <= raw <= 1):
                hit += 1
                assert abs(got - want) < 1e-12, (n, h, got, want)
            else:
                assert got == want, (n, h, got, want)
    assert hit > 100
    # the case the fallback is visible in: n0 = n1 = 7, D = 1/7
    a = np.arange(7.0)
    b = np.arange(7.0) + 0.5
    assert abs(hc.hc_ks_equal_p(7, 1) - stats.ks_2samp(a, b).pvalue) < 1e-15


def test_quotient_from_reciprocal_is_the_division(hc):
    """quot_rn(a, b, RN(1/b)) -- the 3-operation quotient the M-step and the exact-KS recurrence
    use -- has the bits of a / b (Markstein), including divisors with an all-ones mantissa."""
    rng = np.random.default_rng(0)
    n = 2_000_000
    a = rng.standard_normal(n) * 10.0 ** rng.uniform(-8, 8, n)
    b = rng.standard_normal(n) * 10.0 ** rng.uniform(-8, 8, n)
    b[b == 0] = 1.0
    # adversarial divisors: mantissa all ones / nearly so, and small integers
    k = n // 10
    b[:k] = np.ldexp(2.0 - np.ldexp(rng.integers(1, 64, k).astype(np.float64), -52), rng.integers(-20, 20, k))
    b[k:2 * k] = rng.integers(1, 400, k).astype(np.float64)
    a[k:2 * k] = rng.integers(1, 10 ** 6, k).astype(np.float64) * 0.1
    q = np.empty(n)
    hc.hc_quot_rn(a.ctypes.data, b.ctypes.data, q.ctypes.data, n)
    assert np.array_equal(q, a / b)


def test_igamc_against_scipy(hc):
    # chi2.sf(t, f) = Q(f/2, t/2) with the fractional degrees of freedom Hou's combination
    # produces (comparisons.py:696-700): f in (0, 2k], k = 2*context + 1 terms
    from scipy.special import gammaincc
    rng = np.random.default_rng(5)
    a = np.concatenate([rng.uniform(0.05, 12.0, 4000), [0.5, 1.0, 2.5, 5.0, 8.5]])
    x = np.concatenate([10.0 ** rng.uniform(-6, 2.95, 4000), [0.0, 1.0, 700.0, 745.0, 2000.0]])
    worst = 0.0
    for ai, xi in zip(a, x):
        got, want = hc.hc_igamc(ai, xi), float(gammaincc(ai, xi))
        assert close_log10(got, want, rtol=1e-9, atol=1e-12), (ai, xi, got, want)
        if want > 1e-300:
            worst = max(worst, abs(got - want) / want)
    assert worst < 1e-9
    # the denormal range is flushed to 0 exactly where SciPy does
    for ai, xi in [(5.056802588890762, 765.6748459794617), (4.8176392460683, 742.9093387455972),
                   (4.8176392460683, 735.0), (2.5, 712.0), (2.5, 716.0), (2.5, 720.0)]:
        assert (hc.hc_igamc(ai, xi) == 0.0) == (float(gammaincc(ai, xi)) == 0.0), (ai, xi)
    assert hc.hc_igamc(3.0, 0.0) == 1.0
    assert hc.hc_igamc(3.0, float('inf')) == 0.0
    assert np.isnan(hc.hc_igamc(3.0, float('nan')))


def test_chi2_matches_the_reference_golden_results_database(hc):
    """The kernel's chi-squared routine (ncb_math.cuh: chi2_2x2_p, called by gmm_tail on the +1 table)
    on the contingency tables behind the 1241 tested rows of the reference's golden results database
    (tests/golden/golden_db.npz: a real run, real SciPy): its GMM_chi2_pvalue to 1e-9."""
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_db.npz'), allow_pickle=False)
    cols = [str(c) for c in g['columns']]
    n = 0
    for tx in ('tx0', 'tx1'):
        d = g[tx]
        col = lambda c: d[:, cols.index(c)]                          # noqa: E731
        kd = (col('KD1_mod') + col('KD2_mod'), col('KD1_unmod') + col('KD2_unmod'))
        wt = (col('WT1_mod') + col('WT2_mod'), col('WT1_unmod') + col('WT2_unmod'))
        p = col('GMM_chi2_pvalue')
        for i in np.nonzero(np.isfinite(p))[0]:
            got = hc.hc_chi2_2x2_p(kd[0][i] + 1, kd[1][i] + 1, wt[0][i] + 1, wt[1][i] + 1)
            assert abs(got - p[i]) <= 1e-9 * p[i], (i, got, p[i])
            n += 1
    assert n == 1241


def test_ks_equal_sizes_match_the_reference_golden_results_database(hc):
    """The kernels' equal-size exact-KS routine (ks_equal_p, the form ks_equal_kernel evaluates) on the
    fixture positions with as many reads in one condition as in the other, against the float64
    KS p-values a real run (real SciPy) stored in the reference's golden results database."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from oracle import prep as oprep, stats as ostats
    from parity_utils import oracle_rank_stats
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'fixture_cfg1.npz'), allow_pickle=False)
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_db.npz'), allow_pickle=False)
    cols = [str(c) for c in g['columns']]
    db_tx = [str(t) for t in g['transcripts']]
    n = ok = 0
    for ti in range(int(z['n_transcripts'])):
        db = g[f"tx{db_tx.index(str(z[f'tx{ti}/name']))}"]
        data, samples, conditions, positions = oprep.prepare_data(
            z[f'tx{ti}/raw'], z[f'tx{ti}/sample_ids'], z[f'tx{ti}/condition_ids'])
        data, positions = oprep.filter_low_cov_positions(data, positions, conditions, 1)
        std, _, bad, _ = ostats.outlier_standardise(data)
        std, positions = std[~bad], positions[~bad]
        rs = oracle_rank_stats(std, conditions)
        row = {int(p): i for i, p in enumerate(db[:, 0])}
        for i, p in enumerate(positions):
            if int(p) not in row:
                continue
            for v, name in enumerate(('KS_intensity_pvalue', 'KS_dwell_pvalue')):
                n0, n1, h = int(rs['n0'][i, v]), int(rs['n1'][i, v]), int(rs['ks_h'][i, v])
                if n0 != n1 or n0 == 0:
                    continue
                want = db[row[int(p)], cols.index(name)]
                got = min(1.0, max(0.0, hc.hc_ks_equal_p(n0, h)))
                n += 1
                ok += abs(got - want) <= 1e-9 * want
    # the few misses are the rows where the database predates a rule of the current reader
    assert n > 1000 and ok >= 0.99 * n, (ok, n)
