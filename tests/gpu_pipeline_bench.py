"""
Files to result frames on one GPU: the whole of `Worker.run` (run.py:376-509) for a set of
transcripts -- native readers (Uncalled4 BAM or SQLite blobs), read selection, rows -> pinned CSR,
ncb_compare, frame assembly -- through readers.TranscriptPipeline.run_stream (read / pack of the next
batch under the kernels of this one).  Supplementary to bench.py, which starts from packed batches.

    python tests/gpu_pipeline_bench.py [--reads 400] [--positions 2000] [--transcripts 24] [--batch 8]

Prints one JSON line per input format: tested positions/s from files, and where the time goes.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import bench_readers as BR  # noqa: E402
from nanocompore_b200 import readers as R  # noqa: E402
from nanocompore_b200.comparisons import TranscriptComparator  # noqa: E402
from nanocompore_b200.config import PathConfig  # noqa: E402


class Transcript:
    def __init__(self, ref_id, name, seq):
        self.id, self.name, self.seq, self.length = ref_id, name, seq, len(seq)


class Worker:
    def log(self, level, msg):
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reads', type=int, default=400, help='reads per transcript (all samples)')
    ap.add_argument('--positions', type=int, default=2000)
    ap.add_argument('--transcripts', type=int, default=24)
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--repeat', type=int, default=16, help='device path: the task list is the transcripts this many times')
    ap.add_argument('--device-batches', type=int, nargs='+', default=[32, 64, 128])
    ap.add_argument('--device-only', action='store_true', help='only the device-side decoding pipeline (profiling runs)')
    args = ap.parse_args()
    rng = np.random.default_rng(5)
    with tempfile.TemporaryDirectory() as tmp:
        refs, bams, dbs = BR.make_files(tmp, args.transcripts, args.positions, args.reads // 4)
        seq = ''.join(rng.choice(list('ACGT'), args.positions))
        transcripts = [Transcript(i + 1, name, seq) for i, (name, _) in enumerate(refs)]
        for kind, paths in (('uncalled4_bam', bams), ('sqlite_db', dbs)):
            if args.device_only and kind != 'uncalled4_bam':
                continue
            key = 'bam' if kind == 'uncalled4_bam' else 'db'
            cfg = PathConfig(dict(data={'kd': {'kd1': {key: paths['kd1']}, 'kd2': {key: paths['kd2']}},
                                        'wt': {'wt1': {key: paths['wt1']}, 'wt2': {key: paths['wt2']}}},
                                  depleted_condition='kd', kit='RNA002', comparison_methods=['GMM', 'KS'],
                                  min_coverage=30))
            reader = R.Uncalled4Reader(cfg) if kind == 'uncalled4_bam' else R.DbReader(cfg)
            if args.device_only:
                device_pipeline(cfg, reader, transcripts, args)
                reader.close()
                continue
            pipe = R.TranscriptPipeline(cfg, reader, TranscriptComparator(cfg, Worker()), 'cuda:0')
            list(pipe.run_stream(transcripts[:2], batch_size=2))          # warm-up: context, buffers
            t0 = time.perf_counter()
            rows = 0
            for _, df in pipe.run_stream(transcripts, batch_size=args.batch):
                rows += 0 if df is None else len(df)
            total = time.perf_counter() - t0
            # the stages alone, for the breakdown
            t0 = time.perf_counter()
            prepared = pipe.prepare(transcripts[:args.batch])
            t_prepare = (time.perf_counter() - t0) / args.batch
            t0 = time.perf_counter()
            pipe.finish(prepared)
            t_finish = (time.perf_counter() - t0) / args.batch
            # finer: the host stages of one batch, one after the other
            tb = transcripts[:args.batch]
            t0 = time.perf_counter()
            raw = [reader.read(t.name, t.length) for t in tb]
            t_read = (time.perf_counter() - t0) / args.batch
            t0 = time.perf_counter()
            sel = [R.select_reads(r, cfg.get_downsample_high_coverage(), cfg.get_min_coverage(), 0) for r in raw]
            t_select = (time.perf_counter() - t0) / args.batch
            t0 = time.perf_counter()
            packed = R.pack_rows([r for r, _ in sel], 0, pin=True)
            t_pack = (time.perf_counter() - t0) / args.batch
            t0 = time.perf_counter()
            res = pipe._comparator.compare_packed(packed[0], 'cuda:0', min_coverage=30)
            t_compare = (time.perf_counter() - t0) / args.batch
            stages = {"read": 1e3 * t_read, "select": 1e3 * t_select, "pack_pinned": 1e3 * t_pack,
                      "compare_packed": 1e3 * t_compare, "frames": 1e3 * (t_finish - t_compare)}
            print(json.dumps({
                "format": kind, "transcripts": len(transcripts), "positions_per_transcript": args.positions,
                "reads_per_transcript": args.reads, "host_cores": os.cpu_count(), "batch": args.batch,
                "tested_positions": rows, "seconds": total, "tested_positions_per_s": rows / total,
                "ms_per_transcript": {"read_select_pack": 1e3 * t_prepare, "compare_and_frames": 1e3 * t_finish,
                                      "stages": stages}}),
                flush=True)
            if kind == 'uncalled4_bam':
                device_pipeline(cfg, reader, transcripts, args)
            reader.close()


def device_pipeline(cfg, reader, transcripts, args):
    """The same files through DeviceTranscriptPipeline: compressed members up, decoding on the GPU, three
    batches in flight (stage / GPU / frames).  The task list is the transcripts ``--repeat`` times over
    (writing more distinct synthetic files costs 0.8 s each; nothing is cached between tasks: every task
    stages, uploads, inflates and decodes its members again)."""
    import torch
    from nanocompore_b200.device_reader import DeviceTranscriptPipeline
    tasks = transcripts * args.repeat
    for batch in args.device_batches:
        pipe = DeviceTranscriptPipeline(cfg, reader, TranscriptComparator(cfg, Worker()), 'cuda:0')
        list(pipe.run_stream(tasks[:3 * batch], batch_size=batch))             # warm-up: buffers of every slot
        t0 = time.perf_counter()
        rows = 0
        for _, df in pipe.run_stream(tasks, batch_size=batch):
            rows += 0 if df is None else len(df)
        total = time.perf_counter() - t0
        tb = tasks[:batch]
        t0 = time.perf_counter()
        prepared = pipe.prepare(tb)
        t_stage = (time.perf_counter() - t0) / len(tb)
        info = pipe._decoder.staged_bytes(prepared[1])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        launched = pipe.launch(prepared)
        t_launch = (time.perf_counter() - t0) / len(tb)
        launched[2].synchronize()
        t_gpu = (time.perf_counter() - t0) / len(tb)
        t0 = time.perf_counter()
        pipe.collect(launched)
        t_frames = (time.perf_counter() - t0) / len(tb)
        # the decode kernels alone: upload + kernels, CUDA events
        prepared = pipe.prepare(tb)
        st = torch.cuda.Stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        pipe._decoder.launch(prepared[1], st, 0)
        e1.record(st)
        pipe._decoder.collect(prepared[1])
        torch.cuda.synchronize()
        ms_decode = e0.elapsed_time(e1) / len(tb)
        print(json.dumps({
            "format": "uncalled4_bam, device-side decoding", "transcripts": len(tasks),
            "distinct_transcripts": len(transcripts),
            "positions_per_transcript": args.positions, "reads_per_transcript": args.reads, "host_cores": os.cpu_count(),
            "batch": batch, "tested_positions": rows, "seconds": total, "tested_positions_per_s": rows / total,
            "ms_per_transcript": {"stage_compressed_members": 1e3 * t_stage, "launch_call": 1e3 * t_launch,
                                  "upload_decode_compare": 1e3 * t_gpu, "frames": 1e3 * t_frames,
                                  "upload_and_decode_kernels": ms_decode,
                                  "whole_run": 1e3 * total / len(tasks)},
            "staged_per_batch": info}), flush=True)
        pipe.close()


if __name__ == '__main__':
    main()
