"""The host-side CSR packer of libncb (ncb_pack_count / ncb_pack_fill) against the numpy packing."""
import numpy as np
import pytest

from nanocompore_b200.engine import pack_csr, pack_csr_numpy
from nanocompore_b200.synthetic import make_transcript


def dense(rng, P, R, V=2, nan_rate=0.3):
    d = rng.standard_normal((P, R, V)).astype(np.float32)
    d[rng.random((P, R)) < nan_rate] = np.nan
    return d


@pytest.mark.parametrize('threads', [1, 3, 8])
def test_matches_numpy_packing(threads):
    rng = np.random.default_rng(0)
    txs = []
    for P, R, V in ((17, 9, 2), (1, 1, 2), (40, 33, 3), (0, 5, 2), (64, 200, 2)):
        d = dense(rng, P, R, V)
        if P > 3:
            d[2] = np.nan                      # a position without any read
            d[3, :, 1] = np.nan                # dwell missing, intensity present: reads are kept
        txs.append((d, rng.integers(0, 4, R), rng.integers(0, 2, R)))
    batch, ptx, pidx = pack_csr(txs, threads=threads)
    off, sig, lab = pack_csr_numpy(txs)
    assert (batch.pos_offsets.numpy() == off).all()
    assert np.array_equal(batch.signal.numpy(), sig, equal_nan=True)
    assert (batch.label.numpy() == lab).all()
    assert ptx.tolist() == sum(([i] * t[0].shape[0] for i, t in enumerate(txs)), [])
    assert batch.n_positions == sum(t[0].shape[0] for t in txs) and batch.n_entries == off[-1]


def test_empty_and_pinned():
    batch, ptx, pidx = pack_csr([])
    assert batch.n_positions == 0 and batch.n_entries == 0 and ptx.size == 0
    rng = np.random.default_rng(1)
    tx = make_transcript(rng, 50, 20)
    a, _, _ = pack_csr([(tx.data, tx.samples, tx.conditions)], pin=False)
    off, sig, lab = pack_csr_numpy([(tx.data, tx.samples, tx.conditions)])
    assert np.array_equal(a.signal.numpy(), sig, equal_nan=True)
    # read order inside a position is the column order of the dense tensor
    p = 30
    cols = np.nonzero(~np.isnan(tx.data[p, :, 0]))[0]
    assert np.array_equal(a.signal.numpy()[off[p]:off[p + 1], 0], tx.data[p, cols, 0])


def test_rejects_bad_arguments():
    from nanocompore_b200 import _lib as L
    lib = L.load()
    c = np.zeros(4, dtype=np.int64)
    assert lib.ncb_pack_count(None, 4, 2, 2, c.ctypes.data, 1) == L.NCB_ERR_INVALID
    d = np.zeros((4, 2, 1), dtype=np.float32)
    assert lib.ncb_pack_count(d.ctypes.data, 4, 2, 1, c.ctypes.data, 1) == L.NCB_ERR_INVALID


def test_pack_csr_rejects_mismatched_ids():
    import pytest
    from nanocompore_b200.engine import pack_csr
    data = np.zeros((3, 5, 2), dtype=np.float32)
    with pytest.raises(ValueError, match='5 reads'):
        pack_csr([(data, np.zeros(4, dtype=np.int64), np.zeros(4, dtype=np.int64))])
    with pytest.raises(ValueError):
        pack_csr([(data, np.zeros(5, dtype=np.int64), np.zeros(4, dtype=np.int64))])


def test_int16_wire_is_lossless_on_uncalled4_like_data():
    """NCB_LAYOUT_CSR_I16: Uncalled4 hands over int16 tag values (uncalled4.py:118-192); packed as
    int16 codes they decode to the float32 entries bit for bit, NaN included."""
    rng = np.random.default_rng(5)
    txs = []
    for P, n in ((60, 25), (31, 40)):
        tx = make_transcript(rng, P, n, uncalled4=True)
        tx.data[3, :, 0] = np.nan              # intensity missing, dwell present: entries are kept
        txs.append((tx.data, tx.samples, tx.conditions))
    f32, _, _ = pack_csr(txs)
    i16, _, _ = pack_csr(txs, wire='i16', threads=3)
    auto, _, _ = pack_csr(txs, wire='auto')
    assert f32.wire == 'f32' and i16.wire == 'i16' and auto.wire == 'i16'
    assert i16.signal.dtype.itemsize == 2 and i16.nbytes() < 0.6 * f32.nbytes()
    assert np.array_equal(i16.pos_offsets.numpy(), f32.pos_offsets.numpy())
    assert np.array_equal(i16.label.numpy(), f32.label.numpy())
    assert np.array_equal(i16.signal_f32().numpy(), f32.signal.numpy(), equal_nan=True)
    assert np.array_equal(auto.signal.numpy(), i16.signal.numpy())
    assert (i16.signal.numpy()[np.isnan(f32.signal.numpy())] == -32768).all()


@pytest.mark.parametrize('bad', [0.5, 40000.0, -32768.0, np.inf])
def test_int16_wire_refuses_values_that_are_not_codes(bad):
    rng = np.random.default_rng(6)
    tx = make_transcript(rng, 20, 10, uncalled4=True)
    tx.data[7, 4, 1] = bad
    item = [(tx.data, tx.samples, tx.conditions)]
    with pytest.raises(ValueError, match='int16'):
        pack_csr(item, wire='i16')
    auto, _, _ = pack_csr(item, wire='auto')           # falls back to float32 entries
    f32, _, _ = pack_csr(item)
    assert auto.wire == 'f32' and np.array_equal(auto.signal.numpy(), f32.signal.numpy(), equal_nan=True)


def test_label_runs_of_a_packed_batch():
    """ncb_pack_label_runs: run ends per position from the label bytes (NCB_LAYOUT_CSR_I16_RUNS); labels that
    do not come as runs in one common order are refused (such a batch keeps its label bytes)."""
    import torch
    from nanocompore_b200 import _lib as L
    from nanocompore_b200.engine import CsrBatch
    rng = np.random.default_rng(4)
    txs = [make_transcript(rng, P, n, uncalled4=True) for P, n in [(60, 30), (25, 200), (7, 3)]]
    items = [(t.data, t.samples, t.conditions) for t in txs]
    d, s, c = items[1]
    items[1] = (d[:, s != 1], s[s != 1], c[s != 1])       # a sample without reads on one transcript
    items[0][0][3] = np.nan                               # a position without reads
    plain = pack_csr(items, wire='i16')[0]
    runs = pack_csr(items, wire='i16r')[0]
    assert runs.wire == 'i16r' and plain.wire == 'i16'
    off, lab = plain.pos_offsets.numpy(), plain.label.numpy()
    ends = runs.run_ends.numpy().astype(np.int64)
    assert ends.shape == (plain.n_positions, runs.run_labels.size)
    assert runs.nbytes() == plain.nbytes() - lab.size + 2 * ends.size
    for p in range(plain.n_positions):
        seg = lab[off[p]:off[p + 1]]
        rebuilt = np.repeat(runs.run_labels, np.diff(np.concatenate([[0], ends[p]])))
        np.testing.assert_array_equal(rebuilt, seg)
    # out of order inside a position / unknown label / too many runs
    bad = CsrBatch(plain.pos_offsets, plain.signal, plain.label.clone(), plain.max_entries)
    n0 = int(off[1])
    bad.label[:n0] = torch.flip(bad.label[:n0], [0])
    assert bad.with_label_runs() is None
    lib = L.load()
    ends16 = np.zeros((plain.n_positions, 2), np.uint16)
    two = np.array([lab[0], lab[0] ^ 1], np.uint8)
    assert lib.ncb_pack_label_runs(off.ctypes.data, lab.ctypes.data, plain.n_positions, two.ctypes.data, 2,
                                   ends16.ctypes.data, 2) == L.NCB_ERR_INVALID
    assert lib.ncb_pack_label_runs(off.ctypes.data, lab.ctypes.data, plain.n_positions, two.ctypes.data, 17,
                                   ends16.ctypes.data, 2) == L.NCB_ERR_INVALID
    dup = np.array([lab[0], lab[0]], np.uint8)
    assert lib.ncb_pack_label_runs(off.ctypes.data, lab.ctypes.data, plain.n_positions, dup.ctypes.data, 2,
                                   ends16.ctypes.data, 2) == L.NCB_ERR_INVALID
