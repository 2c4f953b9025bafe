"""Edge cases of the C-ABI path on a GPU: empty and degenerate inputs must give the reference's
answer (NaN / dropped rows), never a fault."""
import ctypes as C

import numpy as np
import pytest
import torch

from nanocompore_b200 import _lib as L
from nanocompore_b200.engine import CsrBatch, make_labels, pack_csr

pytestmark = pytest.mark.gpu


def labels(R):
    return torch.from_numpy(make_labels(np.arange(R) % 4, (np.arange(R) % 4) // 2)).cuda()


def test_empty_batches(engine_factory):
    eng = engine_factory(tests=('GMM', 'KS', 'MW', 'TT'), min_coverage=0)
    res = eng.compare_dense(torch.zeros((0, 8, 2), device='cuda'), labels(8)).numpy()
    assert res['status'].size == 0
    batch, _, _ = pack_csr([])
    res = eng.compare_csr(batch.to('cuda')).numpy()
    assert res['status'].size == 0
    # positions without any read, and a dense tensor without reads at all
    off = torch.zeros(6, dtype=torch.int64, device='cuda')
    res = eng.compare_csr(CsrBatch(off, torch.zeros((0, 2), device='cuda'),
                                   torch.zeros(0, dtype=torch.uint8, device='cuda'))).numpy()
    assert ((res['status'] & L.NCB_POS_BAD_STD) != 0).all() and np.isnan(res['pvalues']).all()
    res = eng.compare_dense(torch.zeros((3, 0, 2), device='cuda'), labels(0)).numpy()
    assert ((res['status'] & L.NCB_POS_DROP_MASK) != 0).all()


def test_degenerate_positions(engine_factory):
    eng = engine_factory(tests=('GMM', 'KS', 'MW', 'TT'), min_coverage=0, with_logit=True)
    rng = np.random.default_rng(0)
    R = 40
    data = rng.standard_normal((8, R, 2)).astype(np.float32)
    cond = (np.arange(R) % 4) // 2
    data[0] = np.nan                          # nothing at all
    data[1, cond == 1] = np.nan               # one condition missing entirely
    data[2, :, 1] = np.nan                    # dwell missing everywhere
    data[3, :, 0] = 7.0                       # constant intensity -> std 0
    data[4, 1:] = np.nan                      # a single read
    data[5, :, :] = data[5, :1, :]            # all reads identical
    data[6, 2:] = np.nan                      # two reads, same condition? (reads 0,1 are condition 0)
    res = eng.compare_dense(torch.from_numpy(data).cuda(), labels(R), diagnostics=True).numpy()
    st = res['status']
    for p in (0, 2, 3, 4, 5):
        assert st[p] & L.NCB_POS_BAD_STD, p
    assert st[1] == 0 and st[7] == 0
    # one condition only: the tests have nothing to compare -> NaN p-values, no fault;
    # the mixture fit itself is defined (all reads of condition 0)
    for row in ('KS_INTENSITY', 'KS_DWELL', 'MW_INTENSITY', 'TT_DWELL'):
        assert np.isnan(res['pvalues'][L.PV[row]][1])
    assert np.isfinite(res['diag_f64'][L.DF['BIC1']][1])
    assert (res['counts'][2:, :, 1] == 0).all()          # samples of condition 1 saw no read
    # position 6: two reads of one condition
    assert st[6] == 0 or st[6] & L.NCB_POS_BAD_STD
    ok = res['pvalues'][:, 7]
    assert np.isfinite(ok[L.PV['KS_INTENSITY']]) and np.isfinite(ok[L.PV['TT_DWELL']])


def test_min_coverage_above_everything(engine_factory):
    eng = engine_factory(tests=('KS',), min_coverage=1000)
    data = np.random.default_rng(1).standard_normal((5, 30, 2)).astype(np.float32)
    res = eng.compare_dense(torch.from_numpy(data).cuda(), labels(30)).numpy()
    assert (res['status'] == L.NCB_POS_LOW_COVERAGE).all() and np.isnan(res['pvalues']).all()


def test_invalid_arguments_return_errors(engine_factory):
    eng = engine_factory(tests=('KS',))
    lib = eng.lib
    b, r = L.NcbBatch(), L.NcbResults()
    b.layout, b.memory, b.n_positions = 7, L.NCB_MEM_DEVICE, 4
    assert lib.ncb_compare(eng._handle, C.byref(b), C.byref(r), None) == L.NCB_ERR_INVALID
    assert b"required" in lib.ncb_last_error(eng._handle) or b"layout" in lib.ncb_last_error(eng._handle)
    o = L.default_opts()
    o.n_samples = 1000
    assert lib.ncb_set_opts(eng._handle, C.byref(o)) == L.NCB_ERR_INVALID
    with pytest.raises(L.NcbError):
        eng.configure(n_samples=1000)
    assert eng.ks_exact_pvalues([], [], []).size == 0
    assert eng.bh_fdr(np.zeros(0)).size == 0
    assert np.isnan(eng.bh_fdr(np.array([np.nan, np.nan]))).all()


def test_many_samples_and_unknown_sample_ids(engine_factory):
    """Sample ids beyond n_samples are ignored in the counts (the reference only reports the
    configured samples); 64 samples are supported."""
    eng = engine_factory(tests=('GMM',), min_coverage=0, n_samples=64)
    rng = np.random.default_rng(3)
    R = 256
    data = rng.standard_normal((4, R, 2)).astype(np.float32)
    data[:, R // 2:, :] += 3
    samples = np.arange(R) % 64
    cond = (np.arange(R) >= R // 2).astype(int)
    lab = torch.from_numpy(make_labels(samples, cond)).cuda()
    res = eng.compare_dense(torch.from_numpy(data).cuda(), lab, diagnostics=True).numpy()
    tot = (res['counts'][:, 0] + res['counts'][:, 1])
    assert (tot.sum(0) == res['diag_i64'][L.DI['N_GMM']]).all()
    eng4 = engine_factory(tests=('GMM',), min_coverage=0, n_samples=4)
    res4 = eng4.compare_dense(torch.from_numpy(data).cuda(), lab).numpy()
    assert (res4['counts'][:4] == res['counts'][:4]).all()


def test_declared_max_entries_skips_classes_but_not_results(engine_factory):
    """ncb_batch.n_reads of a CSR batch = the most entries a position has (the packers know it):
    the size classes above it are not launched.  Same results as with an undeclared bound; a bound
    that is too low is reported per position (NCB_POS_TOO_MANY_READS), never silently ignored."""
    from nanocompore_b200.synthetic import make_transcript
    eng = engine_factory(tests=('GMM', 'KS'), min_coverage=30, dwell_is_raw=True)
    rng = np.random.default_rng(3)
    txs = [make_transcript(rng, 200, 40), make_transcript(rng, 150, 150)]      # classes 0-1 and 2
    batch, _, _ = pack_csr([(t.data, t.samples, t.conditions) for t in txs])
    counts = np.diff(batch.pos_offsets.numpy())
    assert batch.max_entries == counts.max() and 256 < batch.max_entries <= 512
    dev = batch.to('cuda')
    assert dev.max_entries == batch.max_entries
    n0 = eng.kernel_launches()
    want = eng.compare_csr(dev).numpy()
    launched_declared = eng.kernel_launches() - n0
    dev.max_entries = 0                                                        # undeclared: every class is launched
    n0 = eng.kernel_launches()
    got = eng.compare_csr(dev).numpy()
    launched_all = eng.kernel_launches() - n0
    assert launched_all - launched_declared == 2 * 4                           # four CTA classes x two phases
    for key in ('status', 'pvalues', 'stats', 'counts'):
        a, b = got[key].ravel(), want[key].ravel()
        assert ((a == b) | (np.isnan(a.astype(np.float64)) & np.isnan(b.astype(np.float64)))).all(), key
    dev.max_entries = 100                                                      # too low: the wait reports it,
    from nanocompore_b200.engine import Results                                #   the other rows are complete
    out = Results(dev.n_positions, 4, dev.signal.device)
    with pytest.raises(L.NcbError) as err:
        eng.compare_csr(dev, out)
    assert err.value.code == L.NCB_ERR_TOO_MANY_READS
    res = out.numpy()
    over = counts > 100
    assert over.any() and ((res['status'][over] & L.NCB_POS_TOO_MANY_READS) != 0).all()
    assert ((res['status'][~over] & L.NCB_POS_TOO_MANY_READS) == 0).all()
    same = ~over
    assert (res['pvalues'][L.PV['KS_INTENSITY']][same] == want['pvalues'][L.PV['KS_INTENSITY']][same]).all() or \
        np.array_equal(np.isnan(res['pvalues'][L.PV['KS_INTENSITY']][same]), np.isnan(want['pvalues'][L.PV['KS_INTENSITY']][same]))


@pytest.mark.parametrize('n_small,n_large', [(3, 40), (8, 120), (5, 400), (8, 900), (6, 1800), (8, 3500), (7, 6000)])
def test_exact_mann_whitney_small_against_large(engine_factory, n_small, n_large):
    """A sample of at most 8 values without ties: SciPy enumerates the null distribution whatever the
    size of the other sample (_mannwhitneyu.py:212-223).  The counting row of such a position does not
    fit on chip; it lives in the handle's global scratch -- a p-value, not NaN + a status bit."""
    from oracle import stats as ostats
    from oracle.nonparametric import nonparametric_test
    from parity_utils import assert_pvalues_close
    eng = engine_factory(tests=('MW',), min_coverage=0)
    rng = np.random.default_rng(n_small * 10007 + n_large)
    P, R = 12, n_small + n_large
    data = rng.standard_normal((P, R, 2)).astype(np.float32)
    # the small sample is condition 0 on even positions' layout, condition 1 on the odd seed
    cond = np.zeros(R, dtype=np.int64)
    if n_large % 2:
        cond[:n_large] = 1
    else:
        cond[n_small:] = 1
    shift = np.linspace(-1.2, 1.2, P).astype(np.float32)
    small = (cond == 0) if (cond == 0).sum() == n_small else (cond == 1)
    data[:, small, :] += shift[:, None, None]
    samples = cond * 2
    lab = torch.from_numpy(make_labels(samples, cond)).cuda()
    res = eng.compare_dense(torch.from_numpy(data).cuda(), lab, diagnostics=True).numpy()
    std_data, _, bad, _ = ostats.outlier_standardise(data)
    assert not bad.any()
    assert ((res['status'] & L.NCB_POS_MW_EXACT_SKIPPED) == 0).all()
    want = nonparametric_test('MW', std_data, cond)
    assert_pvalues_close(res['pvalues'][L.PV['MW_INTENSITY']], want['MW_intensity_pvalue'], 'exact MW intensity p')
    assert_pvalues_close(res['pvalues'][L.PV['MW_DWELL']], want['MW_dwell_pvalue'], 'exact MW dwell p')
    # (float32 values of a large sample may collide: SciPy and the kernel then both take the asymptotic form)
    assert (res['diag_i64'][L.DI['MW_TIE_INTENSITY']] == 0).sum() >= P // 2
