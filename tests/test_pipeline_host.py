"""
Host half of readers.TranscriptPipeline on the CPU (read -> select -> pack into one CSR batch), on
the inputs of the GPU pipeline tests (tests/test_pipeline_gpu.py): Uncalled4 BAMs and SQLite stores,
a transcript without reads, downsampling and the motor-dwell column.  The packed batch must equal the
dense packing of the rows the reader returns.
"""
import numpy as np
import pytest

import test_pipeline_gpu as T
from nanocompore_b200 import readers as R
from nanocompore_b200.engine import pack_csr_numpy


@pytest.fixture
def unpinned(monkeypatch):
    pack_rows = R.pack_rows       # pinned memory needs a CUDA context
    monkeypatch.setattr(R, 'pack_rows',
                        lambda rows, offset=0, pin=False, threads=None, wire='f32': pack_rows(rows, offset, wire=wire))


class Comparator:
    def __init__(self):
        self._worker = T.Worker()


@pytest.mark.parametrize('kind', ['bam', 'db'])
def test_prepare_packs_what_the_reader_returns(tmp_path, unpinned, kind):
    rng = np.random.default_rng(31 if kind == 'bam' else 32)
    seq = ''.join(rng.choice(list('ACGT'), 260))
    transcripts = [T.Transcript(1, 'txA|g|1', seq[:240]), T.Transcript(2, 'txB|g|2', seq[:150]),
                   T.Transcript(3, 'txEmpty|g|3', seq[:90])]
    mk, cfg_dict = T.write_inputs(tmp_path, rng, transcripts[:2], [34, 29, 31, 40], kind)
    cfg = mk(dict(cfg_dict, comparison_methods=['GMM', 'KS'], min_coverage=30))
    reader = R.Uncalled4Reader(cfg) if kind == 'bam' else R.DbReader(cfg)
    cmp_ = Comparator()
    pipe = R.TranscriptPipeline(cfg, reader, cmp_, 'cuda:0')
    _, live, row_sets, masks, (batch, ptx, pidx) = pipe.prepare(transcripts)
    assert live == [0, 1] and pipe.skipped == [] and cmp_._worker.messages == [] and masks == [None, None]
    assert batch.n_positions == 240 + 150 and ptx.tolist() == [0] * 240 + [1] * 150
    off, sig, lab = pack_csr_numpy([(r.dense(), r.sample_ids, r.condition_ids) for r in row_sets])
    assert np.array_equal(batch.pos_offsets.numpy(), off)
    assert np.array_equal(batch.label.numpy(), lab)
    # Uncalled4 rows are int16 tag values and travel as int16 entries; SQLite blobs as float32
    assert batch.wire == ('i16' if kind == 'bam' else 'f32')
    assert np.array_equal(batch.signal_f32().numpy(), sig, equal_nan=True)
    reader.close()


def test_prepare_downsample_and_motor(tmp_path, unpinned):
    rng = np.random.default_rng(33)
    seq = ''.join(rng.choice(list('ACGT'), 200))
    transcripts = [T.Transcript(1, 'txA|g|1', seq[:180])]
    mk, cfg_dict = T.write_inputs(tmp_path, rng, transcripts, [45, 44, 46, 47], 'bam')
    cfg = mk(dict(cfg_dict, downsample_high_coverage=60, min_coverage=50))
    pipe = R.TranscriptPipeline(cfg, R.Uncalled4Reader(cfg), Comparator(), 'cuda:0')
    _, live, row_sets, masks, (batch, _, _) = pipe.prepare(transcripts)
    assert live == [0] and row_sets[0].n_reads == 120 and masks[0] is not None and masks[0].shape == (180,)
    assert batch.signal.shape[1] == 2
    cfg = mk(dict(cfg_dict, motor_dwell_offset=4, min_coverage=30))
    pipe = R.TranscriptPipeline(cfg, R.Uncalled4Reader(cfg), Comparator(), 'cuda:0')
    _, live, row_sets, masks, (batch, _, _) = pipe.prepare(transcripts)
    assert live == [0] and masks == [None] and batch.signal.shape[1] == 3
    # third column = the dwell 4 positions downstream of the same read (run.py:577-584)
    r = row_sets[0]
    off = batch.pos_offsets.numpy()
    sig = batch.signal_f32().numpy()
    p = 10
    # a read is kept at a position when any of its three variables is a number (include/ncb.h)
    rows_at_p = np.nonzero(~(np.isnan(r.intensity[:, p]) & np.isnan(r.dwell[:, p]) & np.isnan(r.dwell[:, p + 4])))[0]
    assert off[p + 1] - off[p] == rows_at_p.size
    np.testing.assert_array_equal(sig[off[p]:off[p + 1], 2], r.dwell[rows_at_p, p + 4])
    np.testing.assert_array_equal(sig[off[p]:off[p + 1], 0], r.intensity[rows_at_p, p])
