"""The oracle's restatements of SciPy / statsmodels arithmetic against the real thing."""
import numpy as np
import pytest
from scipy import stats

from oracle.fdr import auto_qvalues, fdr_bh, multipletests_filter_nan
from oracle.ks_exact import ks_exact_p, ks_h, outer_prob_inside_method, prob_outside_square
from oracle.nonparametric import mwu_asymptotic_p, mwu_stats


def test_ks_exact_matches_scipy_on_samples():
    rng = np.random.default_rng(0)
    exact = 0
    for i in range(400):
        n0, n1 = rng.integers(1, 120, size=2)
        if i % 5 == 0:
            n1 = n0
        a = rng.standard_normal(n0)
        b = rng.standard_normal(n1) + rng.uniform(-1, 1)
        if i % 3 == 0:          # ties
            a, b = np.round(a * 3) / 3, np.round(b * 3) / 3
        h, _, _ = ks_h(a, b)
        r = stats.ks_2samp(a, b)
        g = np.gcd(n0, n1)
        assert h == int(round(r.statistic * (n0 // g) * n1))
        p = ks_exact_p(int(n0), int(n1), h)
        assert abs(p - r.pvalue) <= 1e-12 * max(p, 1e-300) or abs(np.log10(p) - np.log10(r.pvalue)) < 1e-8
        exact += p == r.pvalue
    assert exact > 380


def test_ks_inner_routines_bitwise():
    from scipy.stats._stats_py import _compute_outer_prob_inside_method, _compute_prob_outside_square
    rng = np.random.default_rng(1)
    for _ in range(300):
        m, n = rng.integers(1, 400, size=2)
        g = int(np.gcd(m, n))
        lcm = m // g * n
        h = int(rng.integers(1, lcm + 1))
        if m == n:
            assert prob_outside_square(int(m), min(h, int(m))) == _compute_prob_outside_square(int(m), min(h, int(m)))
        else:
            assert outer_prob_inside_method(int(m), int(n), g, h) == _compute_outer_prob_inside_method(int(m), int(n), g, h)


def test_mwu_statistics_and_p():
    rng = np.random.default_rng(2)
    for i in range(200):
        n0, n1 = rng.integers(9, 150, size=2)
        a = rng.standard_normal(n0)
        b = rng.standard_normal(n1) + 0.3
        if i % 2:
            a, b = np.round(a * 2) / 2, np.round(b * 2) / 2
        u2, tie, _, _ = mwu_stats(a, b)
        r = stats.mannwhitneyu(a, b, alternative='two-sided')
        assert u2 == int(round(2 * r.statistic))
        assert abs(mwu_asymptotic_p(u2, tie, n0, n1) - r.pvalue) <= 1e-12 * r.pvalue + 1e-300


def test_bh_known_answer_of_the_reference():
    # /root/reference/tests/test_postprocessing.py:18-24
    p = np.array([0.1, 0.01, np.nan, 0.01, 0.5, 0.4, 0.01, 0.001, np.nan, np.nan, 0.01, np.nan])
    want = np.array([0.13333333, 0.016, np.nan, 0.016, 0.5, 0.45714286, 0.016, 0.008, np.nan, np.nan, 0.016, np.nan])
    assert np.allclose(multipletests_filter_nan(p), want, equal_nan=True)


def test_bh_properties():
    rng = np.random.default_rng(3)
    p = rng.uniform(size=1000) ** 2
    q = fdr_bh(p)
    order = np.argsort(p)
    assert (np.diff(q[order]) >= 0).all() and (q >= p).all() and (q <= 1).all()
    assert stats.false_discovery_control(p, method='bh') == pytest.approx(q, rel=1e-12)
    a = auto_qvalues(np.array([0.01, 0.5, np.nan, 0.2]), np.array(['KS', 'GMM', 'KS', 'KS']))
    assert np.isnan(a[2]) and a[1] == pytest.approx(0.5 * 4)


def test_gmm_definition_on_the_reference_test_arrays_against_scikit_learn():
    """The arrays of the reference's tests/test_comparisons.py:464-504 (default_rng(42), 1 x 100 x 2;
    second half shifted by +3).  oracle/gmm.py follows scikit-learn's GaussianMixture(covariance_type=
    'full'): same BICs as scikit-learn 1.9 with its k-means initialisation -- 536.34
    vs ~560 on the unimodal array (p must be NaN), 718.48 vs 693.69 on the shifted one -- table + 1 =
    [[51, 1], [1, 51]] and chi-squared p = 7.28e-22, the known answers of the survey's probe."""
    from sklearn.mixture import GaussianMixture
    from oracle.gmm import GMM
    from oracle.gmm_test import gmm_test_split
    rng = np.random.default_rng(42)
    X = rng.standard_normal((1, 100, 2))
    samples, cond = np.repeat([0, 1, 2, 3], 25), np.repeat([0, 1], 50)
    ids = {'kd1': 0, 'kd2': 1, 'wt1': 2, 'wt2': 3}

    def sk_bic(x, k):
        return GaussianMixture(k, covariance_type='full', tol=1e-3, reg_covar=1e-6, max_iter=100,
                               random_state=42).fit(x).bic(x)

    # unimodal: one component wins, no p-value
    b1, b2 = float(GMM(1).fit(X).bic(X)[0]), float(GMM(2).fit(X).bic(X)[0])
    assert abs(b1 - 536.3424) < 1e-3 and abs(b1 - sk_bic(X[0], 1)) < 1e-6 * b1
    assert b2 > b1 and 555 < b2 < 565 and 555 < sk_bic(X[0], 2) < 565
    out = gmm_test_split(X, samples, cond, ids, 0)
    assert np.isnan(out['GMM_chi2_pvalue'][0]) and np.isnan(out['GMM_LOR'][0])
    # two clusters
    Y = X.copy()
    Y[:, 50:, :] += 3
    out, fit = gmm_test_split(Y, samples, cond, ids, 0, return_fit=True)
    y32 = Y[0].astype(np.float32).astype(np.float64)          # the reference hands float32 to the mixture
    assert abs(fit['bic1'][0] - 718.4815) < 1e-3 and abs(fit['bic2'][0] - 693.6874) < 1e-3
    assert abs(fit['bic2'][0] - sk_bic(y32, 2)) < 1e-6 * 693.0 and abs(fit['bic1'][0] - sk_bic(y32, 1)) < 1e-6 * 718.0
    assert sorted((fit['cont'][0] + 1).ravel().tolist()) == [1.0, 1.0, 51.0, 51.0]
    assert abs(out['GMM_chi2_pvalue'][0] / 7.27672954e-22 - 1) < 1e-6
    assert abs(abs(float(out['GMM_LOR'][0])) - 7.864) < 1e-3


@pytest.mark.parametrize('D', [2, 3])
def test_em_equals_scikit_learn_from_the_same_start(D):
    """The EM part of the definition, separated from the initialisation: started from the oracle's own
    initial parameters (weights_init / means_init / precisions_init), scikit-learn's GaussianMixture
    performs the same number of iterations and ends with the same BIC and the same hard assignment
    on every position -- ordinary, tiny (up to 12 reads) and well separated ones, in 2 and in 3
    variables (the motor-dwell fit)."""
    import warnings
    from sklearn.mixture import GaussianMixture
    from oracle.gmm import GMM
    rng = np.random.default_rng(7 + D)
    B, N = 240, 160
    X = np.full((B, N, D), np.nan)
    for b in range(B):
        m = int(rng.integers(2 * D, 13)) if b % 3 == 0 else int(rng.integers(30, N + 1))
        pts = rng.standard_normal((m, D))
        if b % 2:
            pts[: m // 3] += rng.normal(0, 2.5, D)
        pts = (pts - pts.mean(0)) / pts.std(0)
        X[b, rng.permutation(N)[:m]] = pts
    g = GMM(2).fit(X)
    w0, mu0, cov0 = g.init_
    bic = np.asarray(g.bic(X))
    pred = g.predict_proba(X).argmax(2)
    checked = 0
    for b in range(B):
        v = X[b][~np.isnan(X[b]).any(1)]
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')       # scikit-learn warns when max_iter is reached
            sk = GaussianMixture(2, covariance_type='full', tol=1e-3, reg_covar=1e-6, max_iter=100,
                                 weights_init=w0[b], means_init=mu0[b],
                                 precisions_init=np.linalg.inv(cov0[b])).fit(v)
        assert sk.n_iter_ == g.n_iter[b], (b, sk.n_iter_, g.n_iter[b])
        assert abs(sk.bic(v) - bic[b]) <= 1e-7 * abs(bic[b]) + 1e-7, (b, sk.bic(v), bic[b])
        assert (sk.predict(v) == pred[b][~np.isnan(X[b]).any(1)]).all(), b
        checked += 1
    assert checked == B
