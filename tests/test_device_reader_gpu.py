"""
The device-side Uncalled4 decoder on the GPU (csrc/decode.cu behind ncb_decoder_*): the CSR batch it leaves
in device memory must be the batch of the host path (``Uncalled4Reader.read`` + ``pack_rows(wire='i16')``)
bit for bit, and ``DeviceTranscriptPipeline`` must return the frames of ``TranscriptPipeline`` -- including
for the transcripts it hands back to the host reader (downsampled, flagged, unreadable).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import signal_files as SF  # noqa: E402
from nanocompore_b200 import readers as R  # noqa: E402
from nanocompore_b200.comparisons import TranscriptComparator  # noqa: E402
from nanocompore_b200.device_reader import DeviceDecoder, DeviceTranscriptPipeline  # noqa: E402
from test_pipeline_gpu import Transcript, Worker, write_inputs  # noqa: E402
from test_readers import bam_config, synthetic_bams  # noqa: E402

pytestmark = pytest.mark.gpu


def host_batch(cfg, transcripts, motor=0):
    reader = R.Uncalled4Reader(cfg, threads=2)
    rows = [reader.read(name, length) for name, length in transcripts]
    batch, _, _ = R.pack_rows(rows, motor, wire='i16')
    reader.close()
    return batch.pos_offsets.numpy(), batch.signal.numpy(), batch.label.numpy(), [r.n_reads for r in rows]


@pytest.mark.parametrize('kit,unsorted,motor', [('RNA002', False, 0), ('RNA004', True, 0), ('RNA002', False, 7)])
def test_device_batch_equals_the_host_batch(tmp_path, kit, unsorted, motor):
    refs, paths = synthetic_bams(tmp_path, kit, 41 + unsorted + motor, n_reads=70, unsorted=unsorted)
    cfg = bam_config(paths, kit=kit, max_invalid_kmers_freq=0.08)
    reader = R.Uncalled4Reader(cfg, threads=2)
    dec = DeviceDecoder(cfg, reader, 0)
    for slot in (0, 1, 0):                      # slots are reused
        dec.stage(slot, refs)
        dec.launch(slot, motor_offset=motor)
        flags, reads, entries = dec.collect(slot)
        off, sig, lab = dec.download(slot, 3 if motor else 2)
        want_off, want_sig, want_lab, want_reads = host_batch(cfg, refs, motor)
        assert (flags == 0).all() and list(reads) == want_reads and entries == want_off[-1]
        assert np.array_equal(off, want_off) and np.array_equal(sig, want_sig) and np.array_equal(lab, want_lab)
    info = dec.staged_bytes(0)
    assert 0 < info['compressed'] < info['inflated'] and info['members'] > 4 and dec.kernel_launches(0) >= 7
    dec.close()
    reader.close()


def test_larger_members_and_many_reads(tmp_path):
    """64 KB members, a few hundred reads per transcript: every lane of the copy loops and long literal runs."""
    rng = np.random.default_rng(77)
    seq = ''.join(rng.choice(list('ACGT'), 900))
    transcripts = [Transcript(i + 1, f'tx{i}|g', seq[:900 - 130 * i]) for i in range(4)]
    mk, cfg_dict = write_inputs(tmp_path, rng, transcripts, [70, 64, 66, 75], 'bam')
    cfg = mk(dict(cfg_dict))
    refs = [(t.name, t.length) for t in transcripts]
    reader = R.Uncalled4Reader(cfg)
    dec = DeviceDecoder(cfg, reader, 0)
    dec.stage(0, refs)
    dec.launch(0)
    flags, reads, _ = dec.collect(0)
    off, sig, lab = dec.download(0)
    want_off, want_sig, want_lab, want_reads = host_batch(cfg, refs)
    assert (flags == 0).all() and list(reads) == want_reads
    assert np.array_equal(off, want_off) and np.array_equal(sig, want_sig) and np.array_equal(lab, want_lab)
    dec.close()
    reader.close()


def frames_equal(a, b):
    assert (a is None) == (b is None)
    if a is not None:
        assert list(a.columns) == list(b.columns)
        assert a.equals(b)


def test_pipeline_frames_equal_the_host_pipeline(tmp_path):
    rng = np.random.default_rng(51)
    seq = ''.join(rng.choice(list('ACGT'), 400))
    transcripts = [Transcript(1, 'txA|g|1', seq[:340]), Transcript(2, 'txB|g|2', seq[:150]),
                   Transcript(3, 'txC|g|3', seq[:260]), Transcript(4, 'txEmpty|g|4', seq[:90])]
    mk, cfg_dict = write_inputs(tmp_path, rng, transcripts[:3], [34, 29, 31, 40], 'bam')
    for extra in (dict(min_coverage=30), dict(min_coverage=30, motor_dwell_offset=4),
                  dict(min_coverage=40, downsample_high_coverage=100), dict(min_coverage=0)):
        cfg = mk(dict(cfg_dict, comparison_methods=['GMM', 'KS', 'MW'], **extra))
        reader = R.Uncalled4Reader(cfg)
        worker = Worker()
        comparator = TranscriptComparator(cfg, worker)
        want = R.TranscriptPipeline(cfg, reader, comparator, 'cuda:0').run(transcripts)
        pipe = DeviceTranscriptPipeline(cfg, reader, comparator, 'cuda:0')
        got = pipe.run(transcripts)
        assert [t.name for t, _ in got] == [t.name for t in transcripts]
        for (_, a), (_, b) in zip(got, want):
            frames_equal(a, b)
        assert got[3][1] is None and sum(len(df) for _, df in got[:3]) > 300
        # three batches in flight: two batches, then one transcript per batch (the slots wrap around)
        for batch_size in (2, 1):
            streamed = list(pipe.run_stream(transcripts, batch_size=batch_size))
            assert [t.name for t, _ in streamed] == [t.name for t in transcripts]
            for (_, a), (_, b) in zip(streamed, want):
                frames_equal(a, b)
        pipe.close()
        reader.close()


def test_flagged_transcripts_go_through_the_host_reader(tmp_path):
    """A dwell that is no int16 code (host path: float32 wire) and a malformed tag (host path: logged,
    skipped) next to an ordinary transcript."""
    rng = np.random.default_rng(5)
    seq = ''.join(rng.choice(list('ACGT'), 150))
    transcripts = [Transcript(1, 'ok', seq), Transcript(2, 'wide', seq), Transcript(3, 'broken', seq)]
    refs = [(t.name, t.length) for t in transcripts]
    paths = []
    for sample in ('kd1', 'kd2', 'wt1', 'wt2'):
        reads = []
        for ri in range(3):
            for k in range(12):
                r = SF.make_uncalled4_read(rng, 150, 'RNA002', f'{sample}-{ri}-{k}', ref=ri, n_segments=1)
                if ri == 1 and k == 3:
                    r['ul_type'] = 'i'
                    r['ul'][5] = 70000
                if ri == 2 and k == 2:
                    r['ur'] = r['ur'][:1]
                reads.append(r)
        p = str(tmp_path / f'{sample}.bam')
        SF.write_bam(p, refs, reads, block_bytes=3000)
        paths.append(p)
    cfg = bam_config(paths, max_invalid_kmers_freq=0.5, comparison_methods=['KS'], min_coverage=5)
    reader = R.Uncalled4Reader(cfg)
    dec = DeviceDecoder(cfg, reader, 0)
    dec.stage(0, refs)
    dec.launch(0)
    flags, reads, _ = dec.collect(0)
    assert list(flags) == [0, 2, 1] and reads[1] == 0 and reads[2] == 0
    dec.close()
    worker = Worker()
    comparator = TranscriptComparator(cfg, worker)
    want = R.TranscriptPipeline(cfg, reader, comparator, 'cuda:0').run(transcripts)
    n_logged = len(worker.messages)
    pipe = DeviceTranscriptPipeline(cfg, reader, comparator, 'cuda:0')
    got = pipe.run(transcripts)
    for (_, a), (_, b) in zip(got, want):
        frames_equal(a, b)
    assert got[0][1] is not None and got[1][1] is not None and got[2][1] is None
    assert pipe.skipped == ['broken'] and len(worker.messages) == 2 * n_logged > 0
    pipe.close()
    reader.close()
