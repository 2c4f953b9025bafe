"""Host logic of the drop-in comparator: the result frame built from libncb's result columns
(comparisons.py:83-84, 108-148, 204-215 of the reference).  No GPU: the columns are made up."""
import numpy as np
import pandas as pd
import pytest
import torch

from nanocompore_b200 import _lib as L
from nanocompore_b200.comparisons import SHIFT_COLUMNS, TranscriptComparator
from nanocompore_b200.config import PathConfig

DATA = {'kd': {'kd1': {}, 'kd2': {}}, 'wt': {'wt1': {}, 'wt2': {}}}


class Worker:
    def __init__(self):
        self.messages = []

    def log(self, level, msg):
        self.messages.append((level, msg))


class Transcript:
    id, name = 7, 'tx'


def columns(P, rng, bad=(), auto_gmm=()):
    status = np.zeros(P, np.int32)
    status[list(bad)] |= L.NCB_POS_BAD_STD
    status[list(auto_gmm)] |= L.NCB_POS_AUTO_GMM
    pv = rng.random((L.NCB_PV_ROWS, P))
    st = rng.random((L.NCB_ST_ROWS, P)).astype(np.float32)
    counts = rng.integers(0, 50, (4, 2, P)).astype(np.int32)
    return status, pv, st, counts


def test_frame_layout_and_dropped_rows():
    rng = np.random.default_rng(1)
    P = 40
    status, pv, st, counts = columns(P, rng, bad=(3, 17))
    worker = Worker()
    cfg = PathConfig(dict(data=DATA, depleted_condition='kd', comparison_methods=['GMM', 'KS', 'MW']))
    df = TranscriptComparator(cfg, worker)._assemble(Transcript, torch.arange(100, 100 + P, dtype=torch.int16),
                                                     status, pv, st, counts, False)
    keep = np.ones(P, bool)
    keep[[3, 17]] = False
    assert list(df.columns[:2]) == ['transcript_id', 'pos'] and list(df.columns[2:14]) == [c for c, _ in SHIFT_COLUMNS]
    assert list(df.columns[14:]) == ['GMM_chi2_pvalue', 'GMM_LOR', 'kd1_mod', 'kd1_unmod', 'kd2_mod', 'kd2_unmod',
                                     'wt1_mod', 'wt1_unmod', 'wt2_mod', 'wt2_unmod', 'KS_intensity_pvalue',
                                     'KS_dwell_pvalue', 'MW_intensity_pvalue', 'MW_dwell_pvalue']
    # rows with an undefined standard deviation are dropped with a warning; labels keep their gaps
    assert list(df.index) == list(np.arange(P)[keep]) and (df['transcript_id'] == 7).all()
    np.testing.assert_array_equal(df['pos'].to_numpy(), np.arange(100, 100 + P)[keep])
    assert any(level == 'warning' and '[103, 117]' in msg for level, msg in worker.messages)
    np.testing.assert_array_equal(df['KS_dwell_pvalue'].to_numpy(), pv[L.PV['KS_DWELL']][keep])
    np.testing.assert_array_equal(df['GMM_LOR'].to_numpy(), st[L.ST['GMM_LOR']][keep])
    np.testing.assert_array_equal(df['wt1_unmod'].to_numpy(), counts[2, 1][keep].astype(np.uint32))
    assert df['kd1_mod'].dtype == np.uint32 and df['GMM_LOR'].dtype == np.float32 and df['pos'].dtype == np.int16


def test_auto_routing_columns():
    """comparisons.py:150-215: with 'auto' every row gets the p-value of its automatic test."""
    rng = np.random.default_rng(2)
    P = 30
    gmm_rows = (2, 3, 11, 29)
    status, pv, st, counts = columns(P, rng, auto_gmm=gmm_rows)
    cfg = PathConfig(dict(data=DATA, depleted_condition='kd', comparison_methods=['auto']))
    df = TranscriptComparator(cfg, Worker())._assemble(Transcript, torch.arange(P, dtype=torch.int16),
                                                       status, pv, st, counts, False)
    is_gmm = np.zeros(P, bool)
    is_gmm[list(gmm_rows)] = True
    assert list(df['auto_test']) == ['GMM' if g else 'KS' for g in is_gmm]
    want = np.where(is_gmm, pv[L.PV['GMM_CHI2']], pv[L.PV['KS_INTENSITY']])
    np.testing.assert_array_equal(df['auto_pvalue'].to_numpy(), want)
    assert df['KS_intensity_pvalue'][is_gmm].isna().all() and df['GMM_chi2_pvalue'][~is_gmm].isna().all()


def test_context_columns_need_the_gpu():
    """The context combination is a libncb kernel (ncb_context_combine): called on its own it makes an
    engine on cuda:0, and without a GPU that fails -- there is no host path to fall back to."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rng = np.random.default_rng(3)
    P = 60
    status, pv, st, counts = columns(P, rng)
    cfg = PathConfig(dict(data=DATA, depleted_condition='kd', comparison_methods=['KS'], sequence_context=2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TranscriptComparator(cfg, Worker())._assemble(Transcript, torch.arange(P, dtype=torch.int16),
                                                      status, pv, st, counts, False)


def test_too_many_reads_raises():
    from nanocompore_b200.comparisons import NanocomporeError
    rng = np.random.default_rng(4)
    status, pv, st, counts = columns(5, rng)
    status[2] |= L.NCB_POS_TOO_MANY_READS
    cfg = PathConfig(dict(data=DATA, depleted_condition='kd'))
    with pytest.raises(NanocomporeError):
        TranscriptComparator(cfg, Worker())._assemble(Transcript, torch.arange(5, dtype=torch.int16),
                                                      status, pv, st, counts, False)


def test_correction_method_is_checked_when_the_configuration_is_read():
    """config.py:175 admits 'fdr_bh' only: anything else is refused up front, with the key's name."""
    import pytest
    from nanocompore_b200.config import PathConfig
    from nanocompore_b200.postprocessing import check_correction_method
    base = dict(data={'kd': {'kd1': {'bam': 'a'}}, 'wt': {'wt1': {'bam': 'b'}}}, depleted_condition='kd')
    assert PathConfig(base).get_correction_method() == 'fdr_bh'
    assert PathConfig(dict(base, correction_method='fdr_bh')).get_correction_method() == 'fdr_bh'
    with pytest.raises(ValueError, match='correction_method'):
        PathConfig(dict(base, correction_method='bonferroni'))
    with pytest.raises(ValueError, match='holm'):
        check_correction_method('holm')


@pytest.mark.parametrize('methods', [['GMM', 'KS', 'MW'], ['auto'], ['KS']])
def test_batch_frames_equal_the_per_transcript_frames(methods):
    """``assemble_batch`` (one frame over the rows of a batch, cut per transcript) gives the frames, row
    labels, dtypes, log messages and the kmer column of ``_assemble`` transcript by transcript."""
    from nanocompore_b200.readers import encode_kmers
    rng = np.random.default_rng(3)

    class T:
        def __init__(self, i, n):
            self.id, self.name, self.length = 10 + i, f'tx{i}', n
            self.seq = ''.join(rng.choice(list('ACGT'), n))

    lens = [50, 1, 37, 80, 12]
    txs = [T(i, n) for i, n in enumerate(lens)]
    starts = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    P = int(starts[-1])
    bad = [starts[0] + 4, starts[3] + 7, starts[3] + 8] + list(range(starts[4], starts[5]))   # tx4: every row
    status, pv, st, counts = columns(P, rng, bad=bad, auto_gmm=range(0, P, 3))
    status[starts[2] + 5] |= L.NCB_POS_MW_EXACT_SKIPPED
    keeps = [rng.random(n) < 0.7 for n in lens]
    keeps[1][:] = False                                   # a transcript with no tested position
    keeps[4][:] = True
    keeps[0][4] = keeps[3][7] = keeps[3][8] = keeps[2][5] = True
    status[np.concatenate([~k for k in keeps])] |= L.NCB_POS_LOW_COVERAGE
    res = {'status': status, 'pvalues': pv, 'stats': st, 'counts': counts}
    cfg = PathConfig(dict(data=DATA, depleted_condition='kd', comparison_methods=methods))
    w1, w2 = Worker(), Worker()
    got = TranscriptComparator(cfg, w1).assemble_batch(txs, starts, keeps, res, kmer_len=5)
    one = TranscriptComparator(cfg, w2)
    for k, t in enumerate(txs):
        a, b = starts[k], starts[k + 1]
        positions = np.nonzero(keeps[k])[0]
        want = one.assemble(t, positions.astype(np.int16), status[a:b][keeps[k]], pv[:, a:b][:, keeps[k]],
                            st[:, a:b][:, keeps[k]], counts[:, :, a:b][:, :, keeps[k]])
        if want is None:
            assert got[k] is None
            continue
        want['kmer'] = encode_kmers(t.seq, want.pos.to_numpy(), 5)
        pd.testing.assert_frame_equal(got[k], want, check_exact=True)
    assert got[1] is None and len(got[4]) == 0 and len(got[0]) > 0
    # the same messages (the batch logs all warnings of a kind together)
    assert sorted(w1.messages) == sorted(w2.messages) and len(w1.messages) >= (3 if 'MW' in methods else 1)
