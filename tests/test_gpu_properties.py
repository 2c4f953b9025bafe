"""
GPU tests at BASELINE.json's full shapes, through size-independent properties (the oracle is too
slow there), plus spot checks against the oracle on a random sample of the same batch:
  config 2  100 transcripts x 1000 positions x 50 reads/condition, KS only        (full size)
  config 3  one bench step: 50 transcripts x 2000 positions x 100 reads/condition  (1/20 of it)
  config 4  150 reads/condition, GMM + MW, BH FDR over the p-value columns          (a shard)
  config 5  5000 reads/condition mixed with a 30..200 reads/condition tail          (a shard)
"""
import numpy as np
import pytest
import torch

from nanocompore_b200 import _lib as L
from nanocompore_b200.engine import make_labels, pack_csr
from nanocompore_b200.synthetic import BASE_SEED, make_skew_tail, make_transcript
from oracle import prep as oprep
from oracle import stats as ostats
from oracle.fdr import multipletests_filter_nan
from oracle.gmm_test import gmm_test_split
from oracle.nonparametric import nonparametric_test
from parity_utils import assert_equal_nan, assert_pvalues_close, oracle_rank_stats

pytestmark = pytest.mark.gpu
SAMPLE_IDS = {'kd1': 0, 'kd2': 1, 'wt1': 2, 'wt2': 3}


def transcripts(config_index, n_tx, P, n, **kw):
    rng = np.random.default_rng(BASE_SEED + config_index)
    return [make_transcript(rng, P, n, **kw) for _ in range(n_tx)]


def run(engine, txs, device=True, diagnostics=True):
    batch, ptx, pidx = pack_csr([(t.data, t.samples, t.conditions) for t in txs], pin=not device)
    if device:
        batch = batch.to('cuda')
    return engine.compare_csr(batch, diagnostics=diagnostics).numpy(), ptx, pidx


def check_ranges(res, tests):
    ok = (res['status'] & L.NCB_POS_DROP_MASK) == 0
    assert ok.any()
    for name in tests:
        p = res['pvalues'][L.PV[name]][ok]
        fin = p[~np.isnan(p)]
        assert ((fin >= 0) & (fin <= 1)).all(), name
    return ok


def spot_check(res, txs, ptx, which, tests, min_cov=30):
    """Oracle comparison on a few transcripts of the batch."""
    for ti in which:
        t = txs[ti]
        data, samples, conditions, positions = oprep.prepare_data(t.data, t.samples, t.conditions)
        c0 = (~np.isnan(data[:, conditions == 0, 0])).sum(1)
        c1 = (~np.isnan(data[:, conditions == 1, 0])).sum(1)
        keep = (c0 >= min_cov) & (c1 >= min_cov)
        sel = np.nonzero(ptx == ti)[0][positions[keep]]
        std, _, bad, _ = ostats.outlier_standardise(data[keep])
        assert not bad.any()
        assert ((res['status'][sel] & L.NCB_POS_DROP_MASK) == 0).all()
        rs = oracle_rank_stats(std, conditions)
        if 'KS' in tests:
            assert_equal_nan(res['diag_i64'][L.DI['KS_H_INTENSITY']][sel], rs['ks_h'][:, 0], 'KS h')
            want = nonparametric_test('KS', std, conditions)
            assert_pvalues_close(res['pvalues'][L.PV['KS_DWELL']][sel], want['KS_dwell_pvalue'], 'KS dwell p')
        if 'MW' in tests:
            assert_equal_nan(res['diag_i64'][L.DI['MW_U2_DWELL']][sel], rs['mw_u2'][:, 1], 'MWU 2U')
            want = nonparametric_test('MW', std, conditions)
            assert_pvalues_close(res['pvalues'][L.PV['MW_INTENSITY']][sel], want['MW_intensity_pvalue'], 'MW p')
        if 'GMM' in tests:
            want, fit = gmm_test_split(std, samples, conditions, SAMPLE_IDS, 0, 'hard', return_fit=True)
            assert_equal_nan(res['diag_i64'][L.DI['CONT_01']][sel], fit['cont'][:, 0, 1].astype(np.int64), 'cont')
            assert_equal_nan(res['diag_i64'][L.DI['EM_ITERS']][sel], fit['gmm2'].n_iter, 'EM iterations')
            assert_pvalues_close(res['pvalues'][L.PV['GMM_CHI2']][sel], want['GMM_chi2_pvalue'], 'chi2 p')
            assert_equal_nan(res['counts'][2, 0][sel].astype(np.uint32), want['wt1_mod'], 'wt1_mod')


def test_config2_full_size_ks_only(engine_factory):
    eng = engine_factory(tests=('KS',), min_coverage=30, dwell_is_raw=True)
    txs = transcripts(2, 100, 1000, 50)
    res, ptx, pidx = run(eng, txs)
    assert res['status'].size == 100_000
    ok = check_ranges(res, ['KS_INTENSITY', 'KS_DWELL'])
    # nothing but KS was asked for
    assert np.isnan(res['pvalues'][L.PV['GMM_CHI2']]).all() and np.isnan(res['pvalues'][L.PV['MW_DWELL']]).all()
    # the statistic is bounded by lcm and consistent with the p-value being 1 only at h == 0 ...
    di = res['diag_i64']
    n0, n1, h = di[L.DI['N0_INTENSITY']][ok], di[L.DI['N1_INTENSITY']][ok], di[L.DI['KS_H_INTENSITY']][ok]
    lcm = n0 // np.gcd(n0, n1) * n1
    assert ((h >= 1) & (h <= lcm)).all()
    # ... and p is a non-increasing function of D = h / lcm for fixed (n0, n1)
    p = res['pvalues'][L.PV['KS_INTENSITY']][ok]
    key = n0 * 1000 + n1
    order = np.lexsort((h / lcm, key))
    same = key[order][1:] == key[order][:-1]
    assert (np.diff(p[order])[same] <= 1e-12).all()
    # modified positions are detected far more often than unmodified ones
    mod = np.concatenate([t.modified for t in txs])[ok]
    assert (p[mod] < 1e-3).mean() > 0.5 > 0.01 > (p[~mod] < 1e-3).mean()
    spot_check(res, txs, ptx, [0, 57], ('KS',))
    # determinism and staging independence: pinned-host staging gives the same bits
    res2, _, _ = run(eng, txs, device=False)
    for k in ('status', 'pvalues', 'stats', 'diag_i64'):
        assert_equal_nan(res[k].ravel(), res2[k].ravel(), k)


def test_config3_step_read_order_invariance(engine_factory):
    """Shuffling the reads of every position must not change any integer statistic, count or
    p-value: the kernels' results may not depend on the read order (only the deterministic
    tie-breaks of the GMM initialisation do, and they need exact distance ties)."""
    eng = engine_factory(tests=('GMM', 'KS'), min_coverage=30, dwell_is_raw=True)
    txs = transcripts(3, 10, 2000, 100)
    res, ptx, _ = run(eng, txs)
    rng = np.random.default_rng(0)
    shuffled = []
    for t in txs:
        perm = rng.permutation(t.data.shape[1])
        shuffled.append(type(t)(t.name, t.data[:, perm], t.samples[perm], t.conditions[perm], t.modified))
    res2, _, _ = run(eng, shuffled)
    ok = check_ranges(res, ['KS_INTENSITY', 'KS_DWELL', 'GMM_CHI2'])
    for row in ('KS_H_INTENSITY', 'KS_H_DWELL', 'N0_INTENSITY', 'N_OUTLIERS', 'N_GMM'):
        assert_equal_nan(res['diag_i64'][L.DI[row]], res2['diag_i64'][L.DI[row]], row)
    assert_equal_nan(res['stats'][:12].ravel(), res2['stats'][:12].ravel(), 'shift statistics')
    assert_equal_nan(res['pvalues'][:2].ravel(), res2['pvalues'][:2].ravel(), 'KS p')
    same = np.isclose(res['pvalues'][L.PV['GMM_CHI2']], res2['pvalues'][L.PV['GMM_CHI2']], rtol=1e-9, equal_nan=True)
    assert same[ok].mean() > 0.999
    spot_check(res, txs, ptx, [3], ('KS', 'GMM'))


def test_config4_shard_gmm_mw_and_bh(engine_factory):
    eng = engine_factory(tests=('GMM', 'MW'), min_coverage=30, dwell_is_raw=True)
    txs = transcripts(4, 40, 1000, 150)
    res, ptx, _ = run(eng, txs)
    ok = check_ranges(res, ['MW_INTENSITY', 'MW_DWELL', 'GMM_CHI2'])
    assert np.isnan(res['pvalues'][L.PV['KS_INTENSITY']]).all()
    spot_check(res, txs, ptx, [11], ('MW', 'GMM'))
    # BH FDR over each p-value column of the shard (postprocessing.py:209-252), on the device
    for row in ('MW_INTENSITY', 'MW_DWELL', 'GMM_CHI2'):
        p = res['pvalues'][L.PV[row]].copy()
        p[~ok] = np.nan
        q = eng.bh_fdr(torch.from_numpy(p).cuda()).cpu().numpy()
        want = multipletests_filter_nan(p)
        np.testing.assert_allclose(q, want, rtol=1e-13, equal_nan=True)
        m = ~np.isnan(p)
        assert (q[m] >= p[m]).all() and (q[m] <= 1).all()
        order = np.argsort(p[m])
        assert (np.diff(q[m][order]) >= 0).all()          # q is monotone in p


def test_config5_skew_high_coverage_and_tail(engine_factory):
    eng = engine_factory(tests=('GMM', 'KS'), min_coverage=30, dwell_is_raw=True)
    rng = np.random.default_rng(BASE_SEED + 5)
    high = [make_transcript(rng, 24, 5000, name='high') for _ in range(2)]
    tail = make_skew_tail(rng, 30, n_positions=200)
    txs = [high[0]] + tail[:15] + [high[1]] + tail[15:]
    res, ptx, _ = run(eng, txs)
    ok = check_ranges(res, ['KS_INTENSITY', 'KS_DWELL', 'GMM_CHI2'])
    assert not (res['status'] & L.NCB_POS_TOO_MANY_READS).any()
    n_reads = res['diag_i64'][L.DI['N0_INTENSITY']] + res['diag_i64'][L.DI['N1_INTENSITY']]
    assert n_reads[ok].max() > 9000 and n_reads[ok].min() < 200
    spot_check(res, txs, ptx, [0, 5, 16], ('KS', 'GMM'))


def test_too_many_reads_is_reported(engine_factory):
    eng = engine_factory(tests=('KS',), min_coverage=0)
    rng = np.random.default_rng(1)
    data = rng.standard_normal((2, L.NCB_MAX_READS + 8, 2)).astype(np.float32)
    lab = torch.from_numpy(make_labels(np.zeros(data.shape[1]), np.arange(data.shape[1]) % 2)).cuda()
    from nanocompore_b200.engine import Results
    out = Results(2, 4, torch.device('cuda'))
    with pytest.raises(L.NcbError) as err:          # ncb_wait reports it (include/ncb.h)
        eng.compare_dense(torch.from_numpy(data).cuda(), lab, out)
    assert err.value.code == L.NCB_ERR_TOO_MANY_READS
    res = out.numpy()
    assert ((res['status'] & L.NCB_POS_TOO_MANY_READS) != 0).all()
    assert np.isnan(res['pvalues']).all() and np.isnan(res['stats']).all() and (res['counts'] == 0).all()


def test_soft_cluster_counts(engine_factory):
    from oracle.gmm_test import gmm_test_split
    eng = engine_factory(tests=('GMM',), min_coverage=0, cluster_counts='soft')
    rng = np.random.default_rng(12)
    tx = make_transcript(rng, 120, 60, mod_fraction=0.5)
    data, samples, conditions, _ = oprep.prepare_data(tx.data, tx.samples, tx.conditions)
    res = eng.compare_dense(torch.from_numpy(data).cuda(),
                            torch.from_numpy(make_labels(samples, conditions)).cuda(), diagnostics=True).numpy()
    std, _, bad, _ = ostats.outlier_standardise(data)
    ok = ~bad
    want = gmm_test_split(std[ok], samples, conditions, SAMPLE_IDS, 0, 'soft')
    for label, sid in SAMPLE_IDS.items():
        for j, kind in enumerate(('mod', 'unmod')):
            got = res['counts'][sid, j][ok].view(np.float32)
            np.testing.assert_allclose(got, want[f'{label}_{kind}'], rtol=2e-6, atol=1e-6, err_msg=f'{label}_{kind}')
    # chi-squared and LOR always come from the hard table (comparisons.py:367)
    assert_pvalues_close(res['pvalues'][L.PV['GMM_CHI2']][ok], want['GMM_chi2_pvalue'], 'chi2 p (soft counts)')


def test_depleted_condition_flips_mod_cluster(engine_factory):
    rng = np.random.default_rng(13)
    tx = make_transcript(rng, 100, 80, mod_fraction=1.0)
    data, samples, conditions, _ = oprep.prepare_data(tx.data, tx.samples, tx.conditions)
    dev = torch.from_numpy(data).cuda()
    lab = torch.from_numpy(make_labels(samples, conditions)).cuda()
    a = engine_factory(tests=('GMM',), min_coverage=30, depleted_condition=0).compare_dense(dev, lab).numpy()
    b = engine_factory(tests=('GMM',), min_coverage=30, depleted_condition=1).compare_dense(dev, lab).numpy()
    ok = ((a['status'] & L.NCB_POS_DROP_MASK) == 0) & ~np.isnan(a['pvalues'][L.PV['GMM_CHI2']])
    assert ok.sum() > 30
    # same clustering; mod/unmod swap wherever the two conditions disagree on the rarer cluster
    tot_a = a['counts'][:, 0] + a['counts'][:, 1]
    tot_b = b['counts'][:, 0] + b['counts'][:, 1]
    assert (tot_a == tot_b).all()
    swapped = (a['counts'][:, 0] == b['counts'][:, 1]).all(0) & (a['counts'][:, 1] == b['counts'][:, 0]).all(0)
    same = (a['counts'] == b['counts']).all((0, 1))
    assert (swapped | same)[ok].all() and swapped[ok].mean() > 0.5
    for label, sid in SAMPLE_IDS.items():
        std, _, _, _ = ostats.outlier_standardise(data[ok])
    want = gmm_test_split(ostats.outlier_standardise(data[ok])[0], samples, conditions, SAMPLE_IDS, 1, 'hard')
    assert_equal_nan(b['counts'][0, 0][ok].astype(np.uint32), want['kd1_mod'], 'kd1_mod with depleted = wt')


def test_pipelined_engines_match_a_single_engine(engine_factory):
    """What bench.py and readers.TranscriptPipeline do: batches issued round-robin on several
    engines / streams (each engine forks its exact-KS kernels to an auxiliary stream), host batches
    and device batches mixed.  Every result column must be bitwise what one engine gives when it
    runs the batches one after the other -- the schedule must not leak into the results."""
    from nanocompore_b200.engine import Results
    opts = dict(tests=('GMM', 'KS', 'MW', 'TT'), min_coverage=30, dwell_is_raw=True, with_logit=True)
    rng = np.random.default_rng(BASE_SEED + 77)
    batches = []
    for b in range(8):
        txs = [make_transcript(rng, 400, int(n)) for n in rng.integers(40, 140, size=6)]
        batches.append(pack_csr([(t.data, t.samples, t.conditions) for t in txs], pin=True)[0])
    ref_engine = engine_factory(**opts)
    want = [ref_engine.compare_csr(b.to('cuda')).numpy() for b in batches]
    engines = [engine_factory(**opts) for _ in range(4)]
    streams = [torch.cuda.Stream() for _ in range(4)]
    results = [Results(b.n_positions, 4, None if i % 2 else torch.device('cuda')) for i, b in enumerate(batches)]
    # odd batches stay in pinned host memory, even ones are device resident (and stay alive: the
    # calls below only enqueue work)
    sources = [b if i % 2 else b.to('cuda') for i, b in enumerate(batches)]
    torch.cuda.synchronize()
    for rep in range(2):                      # the second round reuses warm buffers in another order
        order = range(8) if rep == 0 else reversed(range(8))
        for i in order:
            s = i % 4
            engines[s].compare_csr(sources[i], results[i], stream=streams[s], wait=False)
        for s in range(4):
            engines[s].wait(streams[s])
        torch.cuda.synchronize()
        for i in range(8):
            got = results[i].numpy()
            for key in ('status', 'pvalues', 'stats', 'counts'):
                assert_equal_nan(got[key].ravel(), want[i][key].ravel(), f'round {rep} batch {i} {key}')
