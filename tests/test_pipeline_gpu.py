"""
Signal files -> result frames on the GPU (SURVEY 8f N1 joined to the hot path): synthetic
Uncalled4 BAMs / SQLite stores go through readers.TranscriptPipeline (native reader, rows -> CSR
packer, one ncb_compare with raw dwell and the in-kernel coverage filter) and must give the frames
of the reference's sequence of steps -- _read_data, _prepare_data, _filter_low_cov_positions,
_downsample, compare_transcript -- evaluated by the oracle on the dense tensor.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import signal_files as SF  # noqa: E402
from nanocompore_b200 import readers as R  # noqa: E402
from nanocompore_b200.comparisons import TranscriptComparator  # noqa: E402
from nanocompore_b200.config import PathConfig  # noqa: E402
from oracle import prep as oprep  # noqa: E402
from oracle.comparator import OracleComparator, OracleConfig  # noqa: E402
from parity_utils import assert_equal_nan, assert_pvalues_close  # noqa: E402

pytestmark = pytest.mark.gpu

SAMPLES = ('kd1', 'kd2', 'wt1', 'wt2')
SAMPLE_IDS = {s: i for i, s in enumerate(SAMPLES)}


class Transcript:
    def __init__(self, id, name, seq):
        self.id, self.name, self.seq, self.length = id, name, seq, len(seq)


class Worker:
    def __init__(self):
        self.messages = []

    def log(self, level, msg):
        self.messages.append((level, msg))


def encode_read(inten, dwell, klen, name, ref):
    """One-segment Uncalled4 tags of a read covering [a, b): -32768 for missing kmers inside."""
    covered = np.nonzero(~np.isnan(inten))[0]
    a, b = int(covered[0]), int(covered[-1]) + 1
    uc = np.where(np.isnan(inten[a:b]), -32768, inten[a:b]).astype(np.int64)
    ul = dwell[a:b].astype(np.int64)
    return dict(ref=ref, name=name, flag=0, ur=[a, b + klen - 1], uc=uc[::-1].tolist(), ul=ul[::-1].tolist())


def synthetic_transcript(rng, P, per_sample, modified_frac=0.15):
    """Rows per sample: integer currents (int16 range) and integer dwell samples, a contiguous
    3'-anchored span per read and a few missing kmers inside it."""
    mu = rng.uniform(-6000, 6000, P)
    sd = rng.uniform(300, 900, P)
    modified = rng.random(P) < modified_frac
    rows = {}
    for si, s in enumerate(SAMPLES):
        cond = si // 2
        n = per_sample[si]
        inten = np.rint(rng.normal(mu, sd, (n, P)))
        dwell = np.maximum(1, np.rint(rng.lognormal(3.7, 0.5, (n, P))))
        if cond == 1:
            hit = modified[None, :] & (rng.random((n, P)) < 0.7)
            inten = np.where(hit, inten + np.rint(3 * sd), inten)
            dwell = np.where(hit, np.rint(dwell * 1.6), dwell)
        inten = np.clip(inten, -30000, 30000)
        for k in range(n):
            start = int(rng.integers(0, P // 3))
            inten[k, :start] = np.nan
            gaps = (rng.random(P) < 0.02) & (np.arange(P) > start + 1) & (np.arange(P) < P - 1)
            inten[k, gaps] = np.nan
        rows[s] = (inten, dwell)
    return rows


def write_inputs(tmp_path, rng, transcripts, per_sample, kind):
    klen = 5
    refs = [(t.name, t.length) for t in transcripts]
    per_tx = [synthetic_transcript(rng, t.length, per_sample) for t in transcripts]
    paths = {}
    for s in SAMPLES:
        if kind == 'bam':
            reads = []
            for ti, rows in enumerate(per_tx):
                inten, dwell = rows[s]
                for k in range(inten.shape[0]):
                    reads.append(encode_read(inten[k], dwell[k], klen, f'{rng.integers(1 << 40):012x}', ti))
            paths[s] = str(tmp_path / f'{s}.bam')
            SF.write_bam(paths[s], refs, reads, block_bytes=30000)
        else:
            signal, reads, rid = [], [], 0
            for ti, rows in enumerate(per_tx):
                inten, dwell = rows[s]
                for k in range(inten.shape[0]):
                    d = np.where(np.isnan(inten[k]), np.nan, dwell[k] / 4000.0)
                    signal.append((ti + 1, rid, inten[k], d))
                    reads.append((f'r{rid}', rid, 0.0))
                    rid += 1
            paths[s] = str(tmp_path / f'{s}.db')
            SF.write_db(paths[s], [(t.name, i + 1) for i, t in enumerate(transcripts)], reads, signal)
    key = 'bam' if kind == 'bam' else 'db'
    return PathConfig, dict(data={'kd': {'kd1': {key: paths['kd1']}, 'kd2': {key: paths['kd2']}},
                                  'wt': {'wt1': {key: paths['wt1']}, 'wt2': {key: paths['wt2']}}},
                            depleted_condition='kd', kit='RNA002')


def oracle_frame(cfg, rows, transcript, methods, motor=0):
    """The reference's steps after _read_data (run.py:437-451) on the dense tensor."""
    data, s, c, pos = oprep.prepare_data(rows.dense(), rows.sample_ids, rows.condition_ids, motor)
    data, pos = oprep.filter_low_cov_positions(data, pos, c, cfg.get_min_coverage())
    data, s, c = oprep.downsample(data, s, c, cfg.get_downsample_high_coverage())
    if len(pos) == 0:
        return None
    cmp_ = OracleComparator(OracleConfig(comparison_methods=methods, sample_ids=SAMPLE_IDS,
                                         motor_dwell_offset=motor))
    return cmp_.compare_transcript(transcript.name, data, s, c, pos)


def check_frame(df, want, methods):
    assert (df is None) == (want is None)
    if want is None:
        return
    assert len(df) == len(want)
    np.testing.assert_array_equal(df['pos'].to_numpy(), want['pos'].to_numpy())
    for col in want.columns:
        if col in ('transcript_id', 'pos'):
            continue
        got, ref = df[col].to_numpy(), want[col].to_numpy()
        if 'pvalue' in col:
            assert_pvalues_close(got, ref, col)
        else:
            assert_equal_nan(got.astype(np.float64), ref.astype(np.float64), col)


@pytest.mark.parametrize('kind', ['bam', 'db'])
def test_files_to_frames(tmp_path, kind):
    rng = np.random.default_rng(31 if kind == 'bam' else 32)
    seq = ''.join(rng.choice(list('ACGT'), 260))
    transcripts = [Transcript(1, 'txA|g|1', seq[:240]), Transcript(2, 'txB|g|2', seq[:150]),
                   Transcript(3, 'txEmpty|g|3', seq[:90])]
    mk, cfg_dict = write_inputs(tmp_path, rng, transcripts[:2], [34, 29, 31, 40], kind)
    methods = ['GMM', 'KS', 'MW']
    cfg = mk(dict(cfg_dict, comparison_methods=methods, min_coverage=30))
    reader = R.Uncalled4Reader(cfg) if kind == 'bam' else R.DbReader(cfg)
    comparator = TranscriptComparator(cfg, Worker())
    pipe = R.TranscriptPipeline(cfg, reader, comparator, 'cuda:0')
    out = pipe.run(transcripts)
    assert out[2][1] is None            # no reads: run.py:428-435
    tested = 0
    for (t, df) in out[:2]:
        rows, _ = pipe.read(t)
        want = oracle_frame(cfg, rows, t, methods)
        check_frame(df, want, methods)
        assert list(df['kmer']) == [R.encode_kmer(t.seq[p:p + 5]) for p in df['pos']]
        tested += len(df)
    assert tested > 200
    # the two-stage pipeline (read/pack of the next batch under the kernels of this one) gives
    # the same frames, in order
    streamed = list(pipe.run_stream(transcripts, batch_size=1))
    assert [t.name for t, _ in streamed] == [t.name for t in transcripts]
    for (_, a), (_, b) in zip(streamed, out):
        assert (a is None) == (b is None)
        if a is not None:
            assert a.equals(b)
    reader.close()


def test_files_to_frames_downsample_and_motor(tmp_path):
    """downsample_high_coverage below the read count (run.py:518-529: the positions are chosen
    before the reads are dropped) and motor_dwell_offset > 0 (3-variable CSR batch)."""
    rng = np.random.default_rng(33)
    seq = ''.join(rng.choice(list('ACGT'), 200))
    transcripts = [Transcript(1, 'txA|g|1', seq[:180])]
    mk, cfg_dict = write_inputs(tmp_path, rng, transcripts, [45, 44, 46, 47], 'bam')
    methods = ['GMM', 'KS']
    for extra in (dict(downsample_high_coverage=60, min_coverage=50), dict(motor_dwell_offset=4, min_coverage=30)):
        cfg = mk(dict(cfg_dict, comparison_methods=methods, **extra))
        reader = R.Uncalled4Reader(cfg)
        pipe = R.TranscriptPipeline(cfg, reader, TranscriptComparator(cfg, Worker()), 'cuda:0')
        (t, df), = pipe.run(transcripts)
        rows = reader.read(t.name, t.length)
        want = oracle_frame(cfg, rows, t, methods, extra.get('motor_dwell_offset', 0))
        if 'downsample_high_coverage' in extra:
            assert pipe.read(t)[0].n_reads == 120 < rows.n_reads
        check_frame(df, want, methods)
        assert len(df) > 50
        reader.close()
