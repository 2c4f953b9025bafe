"""The host<->device wire of ncb_compare on a GPU: int16 entries (NCB_LAYOUT_CSR_I16) give the bits of
float32 entries; host results receive the rows of the configured tests and nothing else; a position
beyond the declared bound ends in NCB_ERR_TOO_MANY_READS with the defaults in its row; an entry point
leaves the caller's current device alone."""
import numpy as np
import pytest
import torch

from nanocompore_b200 import _lib as L
from nanocompore_b200.engine import CsrBatch, Results, pack_csr
from nanocompore_b200.synthetic import make_transcript

pytestmark = pytest.mark.gpu


def uncalled4_batches(seed, shapes, **kw):
    rng = np.random.default_rng(seed)
    txs = [make_transcript(rng, P, n, uncalled4=True) for P, n in shapes]
    items = [(t.data, t.samples, t.conditions) for t in txs]
    return pack_csr(items, pin=True, wire='f32', **kw)[0], pack_csr(items, pin=True, wire='i16', **kw)[0], txs


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == 'f':
        return np.array_equal(a, b, equal_nan=True)
    return np.array_equal(a, b)


@pytest.mark.parametrize('where', ['host', 'device'])
def test_int16_entries_give_the_bits_of_float32_entries(engine_factory, where):
    eng = engine_factory(tests=('GMM', 'KS', 'MW', 'TT'), min_coverage=30, dwell_is_raw=True, with_logit=True)
    # warp classes (<= 128, 256, 512 entries) and a CTA class (> 512)
    f32, i16, _ = uncalled4_batches(11, [(120, 40), (90, 100), (40, 230), (12, 600)])
    assert i16.nbytes() < 0.6 * f32.nbytes()
    if where == 'device':
        f32, i16 = f32.to('cuda'), i16.to('cuda')
    a = eng.compare_csr(f32, diagnostics=True).numpy()
    b = eng.compare_csr(i16, diagnostics=True).numpy()
    assert ((a['status'] & L.NCB_POS_DROP_MASK) == 0).sum() > 200
    for k in a:
        assert same(a[k], b[k]), k


@pytest.mark.parametrize('where', ['host', 'device'])
def test_label_runs_give_the_results_of_label_bytes(engine_factory, where):
    """NCB_LAYOUT_CSR_I16_RUNS (4 bytes per entry on the link, labels rebuilt on the device from the run ends
    of every position) against NCB_LAYOUT_CSR_I16: every result column to the bit, ragged positions, empty
    positions, a sample without reads on a transcript, warp and CTA classes."""
    eng = engine_factory(tests=('GMM', 'KS', 'MW', 'TT'), min_coverage=30, dwell_is_raw=True, with_logit=True)
    rng = np.random.default_rng(21)
    txs = [make_transcript(rng, P, n, uncalled4=True) for P, n in [(120, 40), (90, 100), (40, 230), (12, 600)]]
    items = [(t.data, t.samples, t.conditions) for t in txs]
    # a transcript on which sample 2 has no read at all, and one position without any read
    d, s, c = items[1]
    items[1] = (d[:, s != 2], s[s != 2], c[s != 2])
    items[0][0][5] = np.nan
    i16 = pack_csr(items, pin=True, wire='i16')[0]
    runs = pack_csr(items, pin=True, wire='i16r')[0]
    assert runs.wire == 'i16r' and runs.run_labels.size == 4 and runs.nbytes() < 0.82 * i16.nbytes()
    if where == 'device':
        i16, runs = i16.to('cuda'), runs.to('cuda')
    a = eng.compare_csr(i16, diagnostics=True).numpy()
    b = eng.compare_csr(runs, diagnostics=True).numpy()
    assert ((a['status'] & L.NCB_POS_DROP_MASK) == 0).sum() > 200
    for k in a:
        assert same(a[k], b[k]), k


def test_int16_entries_with_the_motor_dwell_column(engine_factory):
    eng = engine_factory(tests=('GMM', 'KS'), min_coverage=30, dwell_is_raw=True)
    rng = np.random.default_rng(12)
    tx = make_transcript(rng, 80, 60, uncalled4=True)
    off = 3
    motor = np.full(tx.data.shape[:2], np.nan, dtype=np.float32)
    motor[:-off] = tx.data[off:, :, 1]
    d3 = np.concatenate([tx.data, motor[:, :, None]], axis=2)
    valid = ~np.isnan(d3).all(2)
    counts = valid.sum(1)
    offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64))
    sig = torch.from_numpy(d3[valid])
    from nanocompore_b200.engine import make_labels
    lab = torch.from_numpy(np.broadcast_to(make_labels(tx.samples, tx.conditions)[None, :], valid.shape)[valid].copy())
    codes = torch.where(torch.isnan(sig), torch.tensor(-32768.0), sig).to(torch.int16)
    a = eng.compare_csr(CsrBatch(offsets.cuda(), sig.cuda(), lab.cuda()), diagnostics=True).numpy()
    b = eng.compare_csr(CsrBatch(offsets.cuda(), codes.cuda(), lab.cuda()), diagnostics=True).numpy()
    assert np.isfinite(a['stats'][L.ST['C1_MEAN_MOTOR']]).sum() > 50
    for k in a:
        assert same(a[k], b[k]), k


@pytest.mark.parametrize('tests,logit', [(('KS',), False), (('GMM', 'MW'), False), (('GMM', 'KS'), True),
                                        (('TT',), False), (('AUTO',), False)])
def test_host_results_receive_the_rows_of_the_configured_tests(engine_factory, tests, logit):
    eng = engine_factory(tests=tests, min_coverage=30, dwell_is_raw=True, with_logit=logit)
    f32, _, _ = uncalled4_batches(13, [(150, 60)])
    dev = eng.compare_csr(f32.to('cuda')).numpy()
    P = f32.n_positions
    host = Results(P, 4, None)
    sentinel = -7.0
    host.pvalues.fill_(sentinel)
    host.stats.fill_(sentinel)
    host.counts.fill_(-7)
    eng.compare_csr(f32, host)
    h = host.numpy()
    assert same(h['status'], dev['status'])
    t = {x.upper() for x in tests}
    ks, mw, tt = bool(t & {'KS', 'AUTO'}), 'MW' in t, 'TT' in t
    gmm = bool(t & {'GMM', 'AUTO'})
    want_pv = [ks, ks, mw, mw, tt, tt, gmm, gmm and logit, gmm and logit]
    for r, w in enumerate(want_pv):
        if w:
            assert same(h['pvalues'][r], dev['pvalues'][r]), r
        else:
            assert (h['pvalues'][r] == sentinel).all(), r
    for r in range(L.NCB_ST_ROWS):
        if r < 12 or (r == 12 and gmm):
            assert same(h['stats'][r], dev['stats'][r]), r
        else:
            assert (h['stats'][r] == sentinel).all(), r
    if gmm:
        assert same(h['counts'], dev['counts'])
    else:
        assert (h['counts'] == -7).all()
    copied = 4 + 8 * sum(want_pv) + 4 * (13 if gmm else 12) + (32 if gmm else 0)
    assert eng.results_host_bytes(P) == copied * P
    # a fresh Results object reads "not tested" in the rows that were not asked for
    fresh = eng.compare_csr(f32).numpy()
    for r, w in enumerate(want_pv):
        if not w:
            assert np.isnan(fresh['pvalues'][r]).all()


def test_a_position_beyond_the_declared_bound_fails_the_wait(engine_factory):
    eng = engine_factory(tests=('GMM', 'KS'), min_coverage=0, dwell_is_raw=True)
    f32, _, txs = uncalled4_batches(14, [(40, 20), (10, 90)])
    counts = np.diff(f32.pos_offsets.numpy())
    assert counts.max() > 128
    lying = CsrBatch(f32.pos_offsets, f32.signal, f32.label, max_entries=100)
    big = counts > 100
    for batch in (lying, lying.to('cuda')):
        res = Results(batch.n_positions, 4, batch.signal.device if batch.signal.is_cuda else None, True)
        with pytest.raises(L.NcbError) as e:
            eng.compare_csr(batch, res)
        assert e.value.code == L.NCB_ERR_TOO_MANY_READS
        r = res.numpy()
        assert (r['status'][big] == L.NCB_POS_TOO_MANY_READS).all()
        assert ((r['status'][~big] & L.NCB_POS_TOO_MANY_READS) == 0).all()
        assert np.isnan(r['pvalues'][:, big]).all() and np.isnan(r['stats'][:, big]).all()
        assert (r['counts'][:, :, big] == 0).all() and (r['diag_i64'][:, big] == 0).all()
        tested = ~big & (r['status'] == 0)
        assert tested.sum() > 20 and np.isfinite(r['pvalues'][L.PV['KS_INTENSITY']][tested]).all()
        # the flag is consumed by the failed wait: the next batch is clean
        ok = eng.compare_csr(f32 if not batch.signal.is_cuda else f32.to('cuda')).numpy()
        assert ((ok['status'] & L.NCB_POS_TOO_MANY_READS) == 0).all()


def test_entry_points_leave_the_current_device_alone(engine_factory):
    eng = engine_factory(tests=('KS',), min_coverage=0)
    before = torch.cuda.current_device()
    f32, _, _ = uncalled4_batches(15, [(20, 20)])
    eng.compare_csr(f32)
    eng.bh_fdr(np.array([0.01, 0.5, np.nan]))
    assert torch.cuda.current_device() == before
    if torch.cuda.device_count() > 1:
        from nanocompore_b200.engine import Engine
        other = Engine(1, tests=('KS',), min_coverage=0)
        other.compare_csr(f32)
        assert torch.cuda.current_device() == before
        other.close()
