"""
GPU parity tests proper: the CUDA path (through the C ABI of libncb.so) against the oracle on
the same seeded inputs.  Integer statistics must be bit-exact, float32 shift statistics
bit-exact under the shared float32 convention (oracle/stats.py), p-values within 1e-5
relative on log10 p (tests/parity_utils.py).
"""
import numpy as np
import pytest
import torch

from nanocompore_b200 import _lib as L
from nanocompore_b200.engine import make_labels, pack_csr
from nanocompore_b200.synthetic import make_transcript
from oracle import prep as oprep
from oracle import stats as ostats
from oracle.gmm_test import gmm_test_split
from oracle.ks_exact import ks_exact_p
from oracle.nonparametric import nonparametric_test
from parity_utils import (assert_equal_nan, assert_pvalues_close, oracle_rank_stats)

pytestmark = pytest.mark.gpu

SAMPLE_IDS = {'kd1': 0, 'kd2': 1, 'wt1': 2, 'wt2': 3}
ALL_TESTS = ('KS', 'MW', 'TT', 'GMM')


def prepared(seed, P, n, **kw):
    rng = np.random.default_rng(seed)
    tx = make_transcript(rng, P, n, **kw)
    data, samples, conditions, positions = oprep.prepare_data(tx.data, tx.samples, tx.conditions)
    return data, samples, conditions, positions


def run_dense(engine, data, samples, conditions):
    dev = torch.from_numpy(data).cuda()
    lab = torch.from_numpy(make_labels(samples, conditions)).cuda()
    return engine.compare_dense(dev, lab, diagnostics=True).numpy()


def check_against_oracle(res, data, samples, conditions, tests=ALL_TESTS, depleted=0,
                         with_logit=True):
    status = res['status']
    # --- shift statistics (comparisons.py:406-433)
    shift = ostats.shift_stats(data, conditions)
    std_data, outliers, bad, (mean1, std1, mean2, std2) = ostats.outlier_standardise(data)
    assert_equal_nan((status & L.NCB_POS_BAD_STD) != 0, bad, 'bad-std pattern')
    ok = ~bad
    for name, row in [(f"c{c+1}_{s}_{d}", L.ST[f"C{c+1}_{s.upper()}_{d.upper()}"])
                      for c in (0, 1) for s in ('mean', 'median', 'sd') for d in ('intensity', 'dwell')]:
        assert_equal_nan(res['stats'][row][ok], shift[name][ok], name)
    di, df = res['diag_i64'], res['diag_f64']
    assert_equal_nan(di[L.DI['N_OUTLIERS']][ok], outliers.sum(1)[ok], 'outlier count')
    assert_equal_nan(df[L.DF['MEAN2_X']][ok].astype(np.float32), mean2[ok, 0], 'mean after filter')
    assert_equal_nan(df[L.DF['STD2_Y']][ok].astype(np.float32), std2[ok, 1], 'std after filter')

    # --- integer rank statistics (bit-exact)
    rs = oracle_rank_stats(std_data[ok], conditions)
    assert_equal_nan(di[L.DI['N0_INTENSITY']][ok], rs['n0'][:, 0], 'n0 intensity')
    assert_equal_nan(di[L.DI['N1_INTENSITY']][ok], rs['n1'][:, 0], 'n1 intensity')
    assert_equal_nan(di[L.DI['N0_DWELL']][ok], rs['n0'][:, 1], 'n0 dwell')
    assert_equal_nan(di[L.DI['N1_DWELL']][ok], rs['n1'][:, 1], 'n1 dwell')
    if 'KS' in tests:
        assert_equal_nan(di[L.DI['KS_H_INTENSITY']][ok], rs['ks_h'][:, 0], 'KS h intensity')
        assert_equal_nan(di[L.DI['KS_H_DWELL']][ok], rs['ks_h'][:, 1], 'KS h dwell')
    if 'MW' in tests:
        assert_equal_nan(di[L.DI['MW_U2_INTENSITY']][ok], rs['mw_u2'][:, 0], 'MWU 2*U1 intensity')
        assert_equal_nan(di[L.DI['MW_U2_DWELL']][ok], rs['mw_u2'][:, 1], 'MWU 2*U1 dwell')
        assert_equal_nan(di[L.DI['MW_TIE_INTENSITY']][ok], rs['mw_tie'][:, 0], 'MWU tie intensity')
        assert_equal_nan(di[L.DI['MW_TIE_DWELL']][ok], rs['mw_tie'][:, 1], 'MWU tie dwell')

    # --- p-values of the nonparametric tests against SciPy (comparisons.py:218-252)
    for t in ('KS', 'MW', 'TT'):
        if t in tests:
            want = nonparametric_test(t, std_data[ok], conditions)
            assert_pvalues_close(res['pvalues'][L.PV[f'{t}_INTENSITY']][ok], want[f'{t}_intensity_pvalue'],
                                 f'{t} intensity p')
            assert_pvalues_close(res['pvalues'][L.PV[f'{t}_DWELL']][ok], want[f'{t}_dwell_pvalue'],
                                 f'{t} dwell p')

    # --- GMM (comparisons.py:338-392) against the shared definition in oracle/gmm.py
    if 'GMM' in tests:
        want, fit = gmm_test_split(std_data[ok], samples, conditions, SAMPLE_IDS, depleted,
                                   'hard', with_logit=with_logit, return_fit=True)
        np.testing.assert_allclose(df[L.DF['BIC1']][ok], fit['bic1'], rtol=1e-9, err_msg='BIC 1 component')
        np.testing.assert_allclose(df[L.DF['BIC2']][ok], fit['bic2'], rtol=1e-7, err_msg='BIC 2 components')
        assert_equal_nan(di[L.DI['EM_ITERS']][ok], fit['gmm2'].n_iter, 'EM iterations')
        cont = fit['cont'].astype(np.int64)
        for (c, k) in ((0, 0), (0, 1), (1, 0), (1, 1)):
            assert_equal_nan(di[L.DI[f'CONT_{c}{k}']][ok], cont[:, c, k], f'contingency[{c}][{k}]')
        np.testing.assert_allclose(df[L.DF['MU0_X']][ok], fit['gmm2'].means[0][:, 0], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(df[L.DF['C1_YY']][ok], fit['gmm2'].covs[1][:, 1, 1], rtol=1e-6, atol=1e-9)
        assert_pvalues_close(res['pvalues'][L.PV['GMM_CHI2']][ok], want['GMM_chi2_pvalue'], 'GMM chi2 p')
        assert_equal_nan(res['stats'][L.ST['GMM_LOR']][ok], want['GMM_LOR'], 'GMM LOR')
        for label, sid in SAMPLE_IDS.items():
            assert_equal_nan(res['counts'][sid, 0][ok].astype(np.uint32), want[f'{label}_mod'], f'{label}_mod')
            assert_equal_nan(res['counts'][sid, 1][ok].astype(np.uint32), want[f'{label}_unmod'], f'{label}_unmod')
        if with_logit:
            assert_pvalues_close(res['pvalues'][L.PV['GMM_LOGIT']][ok], want['GMM_logit_pvalue'], 'GMM logit p')
            np.testing.assert_allclose(res['pvalues'][L.PV['GMM_LOGIT_COEF']][ok], want['GMM_logit_coef'],
                                       rtol=1e-9, atol=1e-12, equal_nan=True)
    return ok


@pytest.mark.parametrize('n,quantised', [(50, False), (50, True), (100, False), (150, True)])
def test_dense_all_tests(engine_factory, n, quantised):
    eng = engine_factory(tests=ALL_TESTS, with_logit=True, min_coverage=0)
    data, samples, conditions, _ = prepared(100 + n + quantised, 160, n, quantised=quantised)
    res = run_dense(eng, data, samples, conditions)
    ok = check_against_oracle(res, data, samples, conditions)
    assert ok.sum() > 100


def test_csr_matches_dense_and_oracle(engine_factory):
    eng = engine_factory(tests=ALL_TESTS, with_logit=True, min_coverage=0)
    txs = []
    for i, n in enumerate((40, 70, 120)):
        data, samples, conditions, _ = prepared(7 + i, 80, n, quantised=bool(i % 2))
        txs.append((data, samples, conditions))
    batch, ptx, pidx = pack_csr(txs, pin=True)
    host = eng.compare_csr(batch, diagnostics=True).numpy()
    dev = eng.compare_csr(batch.to('cuda'), diagnostics=True).numpy()
    for k in ('status', 'pvalues', 'stats', 'counts', 'diag_i64'):
        assert_equal_nan(host[k].ravel(), dev[k].ravel(), f'host vs device staging: {k}')
    for ti, (data, samples, conditions) in enumerate(txs):
        sel = ptx == ti
        sub = {k: (v[..., sel] if v is not None else None) for k, v in host.items()}
        check_against_oracle(sub, data, samples, conditions)


def test_raw_dwell_and_min_coverage(engine_factory):
    """log10(dwell + 1e-10) and the coverage filter inside the kernel (run.py:511-515, 573)."""
    eng = engine_factory(tests=('KS',), min_coverage=30, dwell_is_raw=True)
    rng = np.random.default_rng(5)
    tx = make_transcript(rng, 120, 36, span_start_max=0.6)
    raw = tx.data.copy()
    res = eng.compare_dense(torch.from_numpy(raw).cuda(),
                            torch.from_numpy(make_labels(tx.samples, tx.conditions)).cuda(),
                            diagnostics=True).numpy()
    data, samples, conditions, positions = oprep.prepare_data(tx.data, tx.samples, tx.conditions)
    status = res['status'][positions]          # _prepare_data drops positions without any read
    c0 = (~np.isnan(data[:, conditions == 0, 0])).sum(1)
    c1 = (~np.isnan(data[:, conditions == 1, 0])).sum(1)
    low = (c0 < 30) | (c1 < 30)
    assert low.any() and (~low).any()
    assert_equal_nan((status & L.NCB_POS_LOW_COVERAGE) != 0, low, 'low coverage pattern')
    dropped = np.setdiff1d(np.arange(raw.shape[0]), positions)
    assert ((res['status'][dropped] & L.NCB_POS_LOW_COVERAGE) != 0).all()
    keep = positions[~low]
    shift = ostats.shift_stats(data[~low], conditions)
    # CUDA's and glibc's fp64 log10 may differ in the last bit before the rounding to float32,
    # so the dwell statistics get a one-ulp-level tolerance; intensity stays exact
    assert_equal_nan(res['stats'][L.ST['C1_MEAN_INTENSITY']][keep], shift['c1_mean_intensity'], 'mean')
    np.testing.assert_allclose(res['stats'][L.ST['C2_MEDIAN_DWELL']][keep], shift['c2_median_dwell'], rtol=2e-7)
    std_data, _, bad, _ = ostats.outlier_standardise(data[~low])
    want = nonparametric_test('KS', std_data[~bad], conditions)
    got = res['pvalues'][L.PV['KS_INTENSITY']][keep][~bad]
    assert_pvalues_close(got, want['KS_intensity_pvalue'], 'KS intensity p (raw dwell path)')


@pytest.mark.parametrize('n', [3, 6, 12])
def test_tiny_coverage_fixture_like(engine_factory, n):
    """<= 12 reads per condition as on the reference's fixture: exact MWU, NaN-heavy rows."""
    eng = engine_factory(tests=ALL_TESTS, with_logit=False, min_coverage=0)
    data, samples, conditions, _ = prepared(31 + n, 150, n, gap_rate=0.15)
    res = run_dense(eng, data, samples, conditions)
    check_against_oracle(res, data, samples, conditions, with_logit=False)


@pytest.mark.parametrize('n,cls', [(200, 'warp16'), (300, 'cta128'), (505, 'cta128'), (700, 'cta256'),
                                   (1100, 'cta512'), (2040, 'cta512'), (2600, 'cta1024')])
def test_large_size_classes(engine_factory, n, cls):
    eng = engine_factory(tests=ALL_TESTS, with_logit=False, min_coverage=0)
    P = 12 if n > 1000 else 24
    data, samples, conditions, _ = prepared(900 + n, P, n, quantised=(n in (505, 700)))
    assert {'warp16': 256 < data.shape[1] <= 512, 'cta128': 512 < data.shape[1] <= 1024,
            'cta256': 1024 < data.shape[1] <= 2048, 'cta512': 2048 < data.shape[1] <= 4096,
            'cta1024': data.shape[1] > 4096}[cls]
    res = run_dense(eng, data, samples, conditions)
    check_against_oracle(res, data, samples, conditions, with_logit=False)


def test_ks_exact_pvalue_unit(engine_factory):
    """ncb_ks_exact_pvalues against the oracle restatement of SciPy's exact routines."""
    eng = engine_factory(tests=('KS',))
    rng = np.random.default_rng(11)
    n0 = rng.integers(1, 400, size=600)
    n1 = rng.integers(1, 400, size=600)
    n1[:100] = n0[:100]                              # equal sizes: Horner path
    n0[100:110] = rng.integers(1500, 5000, size=10)  # wavefront kernels, both buffers
    n1[100:105] = rng.integers(1100, 5000, size=5)
    n1[105:110] = rng.integers(2, 900, size=5)
    g = np.gcd(n0, n1)
    lcm = n0 // g * n1
    d = rng.uniform(0.0, 0.6, size=600) ** 2
    h = np.maximum(np.round(d * lcm).astype(np.int64), 0)
    h[590:] = 0
    h[580:590] = lcm[580:590]
    got = eng.ks_exact_pvalues(n0, n1, h)
    want = np.array([ks_exact_p(int(a), int(b), int(c)) for a, b, c in zip(n0, n1, h)])
    assert_pvalues_close(got, want, 'exact KS p')
    # the cell operations are the IEEE operations SciPy performs: the match is bitwise
    assert (got == want).mean() > 0.99


def test_ks_exact_pvalue_cta_class(engine_factory):
    """The CTA-per-problem wavefront (the problems that get a whole CTA): bitwise SciPy's compiled
    lattice recurrence, chunk edges included (m a multiple of 256, one row left over, n = 1 ...)."""
    from scipy.stats._stats_py import _compute_outer_prob_inside_method
    eng = engine_factory(tests=('KS',))
    shapes = [(5000, 4999, 0.05), (5120, 5119, 0.02), (4100, 257, 0.4), (3000, 2048, 0.1), (5121, 5000, 0.03),
              (4097, 4096, 0.06), (2305, 2304, 0.12), (5000, 4861, 0.3), (4608, 3001, 0.08), (10000, 240, 0.5),
              (5120, 1300, 0.2), (3329, 3328, 0.2)]
    n0 = np.array([a for a, _, _ in shapes], dtype=np.int32)
    n1 = np.array([b for _, b, _ in shapes], dtype=np.int32)
    n0[::2], n1[::2] = n1[::2].copy(), n0[::2].copy()        # either sample may be the larger one
    g = np.gcd(n0, n1)
    lcm = n0.astype(np.int64) // g * n1
    h = np.array([max(1, round(d * l)) for (_, _, d), l in zip(shapes, lcm)], dtype=np.int64)
    m, n = np.maximum(n0, n1), np.minimum(n0, n1)
    cells = m * np.minimum(2.0 * h / (m // g) + 1, n + 1)
    assert (cells >= 2 ** 19).all()
    got = eng.ks_exact_pvalues(n0, n1, h)
    want = np.array([min(max(_compute_outer_prob_inside_method(int(a), int(b), int(c), int(d)), 0.0), 1.0)
                     for a, b, c, d in zip(n0, n1, g, h)])
    assert (want > 0).sum() >= 8                            # not only underflowed tails
    assert np.array_equal(got, want)


def test_ks_exact_pvalue_wavefront_edges(engine_factory):
    """The wavefront forms on chunk edges: m a multiple of 32 / 256, one row left over, bands narrower
    than a chunk and wider than the warp form's ring -- bitwise SciPy's compiled recurrence.
    Submitted alone (each problem is then among the dearest of its batch: CTA form) and 400 at once
    (the queue hands several problems to every CTA and warp)."""
    from scipy.stats._stats_py import _compute_outer_prob_inside_method
    eng = engine_factory(tests=('KS',))
    shapes = [(4128, 40, 0.5), (4129, 37, 0.45), (4127, 41, 0.6), (5121, 5119, 0.25), (5100, 4999, 0.28),
              (4096, 31, 0.9), (6000, 23, 0.8), (9000, 1240, 0.3), (2049, 2047, 0.3), (5119, 64, 0.7)]
    for rep in range(2):
        for a, b, d in shapes:
            n0, n1 = (a, b) if rep == 0 else (b, a)
            g = int(np.gcd(n0, n1))
            lcm = n0 // g * n1
            h = max(1, round(d * lcm))
            cells = max(n0, n1) * min(2.0 * h / (max(n0, n1) // g) + 1, min(n0, n1) + 1)
            assert cells >= 131072, (a, b, d, cells)
            got = eng.ks_exact_pvalues([n0], [n1], [h])
            want = min(max(_compute_outer_prob_inside_method(n0, n1, g, h), 0.0), 1.0)
            assert got[0] == want, (n0, n1, h, got[0], want)
    n0 = np.array([a for a, _, _ in shapes] * 40, dtype=np.int32)
    n1 = np.array([b for _, b, _ in shapes] * 40, dtype=np.int32)
    g = np.gcd(n0, n1)
    lcm = n0.astype(np.int64) // g * n1
    h = np.array([max(1, round(d * l)) for (_, _, d), l in zip(shapes * 40, lcm)], dtype=np.int64)
    got = eng.ks_exact_pvalues(n0, n1, h)
    assert np.array_equal(got[:10], got[-10:]) and np.array_equal(got.reshape(40, 10), np.tile(got[:10], (40, 1)))


def test_ks_exact_pvalue_equal_sizes_large(engine_factory):
    """Equal sample sizes up to the 5000-per-condition cap: the Horner product evaluated with table
    reciprocals (ks_prob_outside_square_fast) has the bits of SciPy's divisions, down through the
    denormal range and where the leading factor reaches zero."""
    from oracle.ks_exact import ks_exact_p
    eng = engine_factory(tests=('KS',))
    rng = np.random.default_rng(12)
    n = np.concatenate([rng.integers(2, 5121, size=300), [5120, 5000, 4999, 1, 2, 3, 4097]])
    d = np.concatenate([rng.uniform(0.0, 0.75, size=300) ** 2, [0.4, 0.39, 0.001, 1.0, 0.5, 1.0, 0.3]])
    h = np.clip(np.round(d * n).astype(np.int64), 1, n)
    h[:20] = n[:20]                                  # h = n: every factor of the k = 1 block is <= 0
    h[20:40] = np.maximum(n[20:40] // 2, 1)
    got = eng.ks_exact_pvalues(n, n, h)
    want = np.array([ks_exact_p(int(a), int(a), int(c)) for a, c in zip(n, h)])
    assert ((want > 0) & (want < 1e-280)).sum() >= 3   # the true-division tail is exercised
    assert np.array_equal(got, want)
    # samples beyond the reciprocal table take the plain form
    got = eng.ks_exact_pvalues([6000, 20000], [6000, 20000], [300, 700])
    want = [ks_exact_p(6000, 6000, 300), ks_exact_p(20000, 20000, 700)]
    assert np.array_equal(got, np.array(want))


def test_bh_fdr_unit(engine_factory):
    from oracle.fdr import multipletests_filter_nan
    eng = engine_factory(tests=('KS',))
    rng = np.random.default_rng(3)
    p = rng.uniform(size=100000) ** 3
    p[rng.random(p.size) < 0.1] = np.nan
    p[:50] = p[50:100]                                # ties
    got = eng.bh_fdr(p)
    want = multipletests_filter_nan(p)
    assert_equal_nan(np.isnan(got), np.isnan(want), 'NaN pattern')
    np.testing.assert_allclose(got[~np.isnan(want)], want[~np.isnan(want)], rtol=1e-14)
    # known answer of the reference (tests/test_postprocessing.py:18-24)
    q = eng.bh_fdr(np.array([0.1, 0.01, np.nan, 0.01, 0.5, 0.4, 0.01, 0.001, np.nan, np.nan, 0.01, np.nan]))
    expect = np.array([0.13333333, 0.016, np.nan, 0.016, 0.5, 0.45714286, 0.016, 0.008, np.nan, np.nan, 0.016, np.nan])
    np.testing.assert_allclose(q, expect, rtol=1e-6, equal_nan=True)
    # device-resident variant
    qd = eng.bh_fdr(torch.from_numpy(p).cuda()).cpu().numpy()
    assert_equal_nan(qd, got, 'device vs host staging')


def check_three_variables(res, data, samples, conditions, positions, motor_dwell_offset=7):
    """Every check of the 3-variable path against the oracle; returns (kept positions, fitted in 3 variables)."""
    from oracle.comparator import OracleComparator, OracleConfig
    cmp_ = OracleComparator(OracleConfig(comparison_methods=list(ALL_TESTS), motor_dwell_offset=motor_dwell_offset))
    df = cmp_.compare_transcript(0, data, samples, conditions, positions)
    shift = ostats.shift_stats(data, conditions)
    std_data, outliers, bad, _ = ostats.outlier_standardise(data)
    assert_equal_nan((res['status'] & L.NCB_POS_BAD_STD) != 0, bad, 'bad-std pattern (3 variables)')
    ok = ~bad
    # the last `motor_dwell_offset` positions have no motor dwell
    assert ok.sum() == len(df) and bad.sum() >= motor_dwell_offset
    for c in (1, 2):
        for st in ('mean', 'median', 'sd'):
            for dim, key in (('intensity', 'INTENSITY'), ('dwell', 'DWELL'), ('motor_dwell', 'MOTOR')):
                name = f"c{c}_{st}_{dim}"
                assert_equal_nan(res['stats'][L.ST[f"C{c}_{st.upper()}_{key}"]][ok], shift[name][ok], name)
    assert_equal_nan(res['diag_i64'][L.DI['N_OUTLIERS']][ok], outliers.sum(1)[ok], 'outliers over 3 variables')
    for t in ('KS', 'MW', 'TT'):
        for v in ('intensity', 'dwell'):
            assert_pvalues_close(res['pvalues'][L.PV[f'{t}_{v.upper()}']][ok], df[f'{t}_{v}_pvalue'].to_numpy(),
                                 f'{t} {v} p')
    valid = ~np.isnan(std_data[ok])
    use3 = (valid[:, :, 0] & valid[:, :, 2]).sum(1) >= 0.9 * valid[:, :, 0].sum(1)
    assert_pvalues_close(res['pvalues'][L.PV['GMM_CHI2']][ok], df['GMM_chi2_pvalue'].to_numpy(), 'GMM chi2 p')
    assert_equal_nan(res['stats'][L.ST['GMM_LOR']][ok], df['GMM_LOR'].to_numpy(dtype=np.float32), 'LOR')
    for label, sid in SAMPLE_IDS.items():
        assert_equal_nan(res['counts'][sid, 0][ok].astype(np.uint32), df[f'{label}_mod'].to_numpy(), f'{label}_mod')
        assert_equal_nan(res['counts'][sid, 1][ok].astype(np.uint32), df[f'{label}_unmod'].to_numpy(), f'{label}_unmod')
    # reads in the fit: all three variables where the position is fitted in 3 variables
    n_fit = np.where(use3, valid.all(2).sum(1), (valid[:, :, 0] & valid[:, :, 1]).sum(1))
    cont = sum(res['diag_i64'][L.DI[f'CONT_{c}{k}']] for c in (0, 1) for k in (0, 1))[ok]
    assert_equal_nan(cont, n_fit, 'reads in the fit')
    return ok, use3


@pytest.mark.parametrize('n,gap', [(40, 0.08), (120, 0.03)])
def test_motor_dwell_three_variables(engine_factory, n, gap):
    """motor_dwell_offset > 0: 18 shift statistics, outlier rule over 3 variables, and the split into
    3-variable / 2-variable mixtures by the 90 % rule (comparisons.py:285-335, 395-403)."""
    eng = engine_factory(tests=ALL_TESTS, with_logit=False, min_coverage=0)
    rng = np.random.default_rng(77 + n)
    tx = make_transcript(rng, 90, n, gap_rate=gap)
    data, samples, conditions, positions = oprep.prepare_data(tx.data, tx.samples, tx.conditions, 7)
    assert data.shape[2] == 3
    res = run_dense(eng, data, samples, conditions)
    _, use3 = check_three_variables(res, data, samples, conditions, positions)
    # both kinds of position occur
    assert use3.any() and ((~use3).any() or gap < 0.05)    # the gappy case has both kinds


def test_correct_pvalues_mirror(engine_factory):
    """nanocompore_b200.postprocessing.correct_pvalues = Postprocessor._correct_pvalues
    (postprocessing.py:209-252), auto_qvalue scaling included, against the oracle restatement."""
    import pandas as pd
    from nanocompore_b200.postprocessing import correct_pvalues
    from oracle.fdr import auto_qvalues, multipletests_filter_nan
    eng = engine_factory(tests=('KS',))
    rng = np.random.default_rng(8)
    n = 5000
    df = pd.DataFrame({'pos': np.arange(n), 'KS_intensity_pvalue': rng.random(n) ** 3,
                       'GMM_chi2_pvalue': rng.random(n) ** 2, 'auto_pvalue': rng.random(n) ** 4,
                       'auto_test': np.where(rng.random(n) < 0.3, 'GMM', 'KS'), 'GMM_LOR': rng.random(n)})
    for col in ('KS_intensity_pvalue', 'GMM_chi2_pvalue', 'auto_pvalue'):
        df.loc[rng.random(n) < 0.1, col] = np.nan
    want_auto = auto_qvalues(df['auto_pvalue'].to_numpy(), df['auto_test'].to_numpy())
    want_ks = multipletests_filter_nan(df['KS_intensity_pvalue'].to_numpy())
    out = correct_pvalues(eng, df.copy(), auto_test=df['auto_test'])
    assert list(out.columns) == sorted(out.columns)
    assert {'KS_intensity_qvalue', 'GMM_chi2_qvalue', 'auto_qvalue'} <= set(out.columns) and 'GMM_LOR' in out.columns
    np.testing.assert_allclose(out['KS_intensity_qvalue'].to_numpy(), want_ks, rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(out['auto_qvalue'].to_numpy(), want_auto, rtol=1e-12, equal_nan=True)
