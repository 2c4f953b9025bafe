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
