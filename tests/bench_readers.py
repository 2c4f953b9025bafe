"""
Host-side throughput of the transcript readers: signal files -> CSR batch.

    python tests/bench_readers.py [--reads 400] [--positions 2000] [--transcripts 6]

Writes synthetic Uncalled4 BAMs and SQLite stores (tests/signal_files.py) to a temporary directory
and times, per transcript:
  native      readers.Uncalled4Reader / DbReader + readers.pack_rows (libncb: BGZF inflate, tag
              decoding, rows -> CSR)
  restated    oracle/readers.py, i.e. the reference's sequence of numpy steps (dense NaN tensor per
              transcript, run.py:692-718 / 766-836) followed by the dense packer.  Its BAM leg parses
              records with a pure-Python reader where the reference uses pysam (C), so the BAM
              "restated" time is split into parse and the reference's own per-read numpy work.
Prints one JSON line per format.  No GPU involved.
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

import signal_files as SF  # noqa: E402
from nanocompore_b200 import readers as R  # noqa: E402
from nanocompore_b200.config import PathConfig  # noqa: E402
from nanocompore_b200.engine import pack_csr  # noqa: E402
from oracle import bam_shim  # noqa: E402
from oracle import readers as OR  # noqa: E402

SAMPLES = ('kd1', 'kd2', 'wt1', 'wt2')


def encode_read(inten, dwell, klen, name, ref):
    covered = np.nonzero(~np.isnan(inten))[0]
    a, b = int(covered[0]), int(covered[-1]) + 1
    uc = np.where(np.isnan(inten[a:b]), -32768, inten[a:b]).astype(np.int64)
    return dict(ref=ref, name=name, flag=0, ur=[a, b + klen - 1], uc=uc[::-1].tolist(),
                ul=dwell[a:b].astype(np.int64)[::-1].tolist())


def make_files(tmp, n_tx, P, reads_per_sample, seed=1):
    rng = np.random.default_rng(seed)
    refs = [(f'tx{t}|bench', P) for t in range(n_tx)]
    bams, dbs = {}, {}
    for s in SAMPLES:
        reads, signal, rid = [], [], 0
        for t in range(n_tx):
            inten = np.rint(rng.normal(0, 3000, (reads_per_sample, P)))
            dwell = np.maximum(1, np.rint(rng.lognormal(3.7, 0.5, (reads_per_sample, P))))
            for k in range(reads_per_sample):
                start = int(rng.integers(0, P // 3))
                inten[k, :start] = np.nan
                gaps = (rng.random(P) < 0.02) & (np.arange(P) > start + 1) & (np.arange(P) < P - 1)
                inten[k, gaps] = np.nan
                reads.append(encode_read(inten[k], dwell[k], 5, f'{rng.integers(1 << 40):012x}', t))
                signal.append((t + 1, rid, inten[k], np.where(np.isnan(inten[k]), np.nan, dwell[k] / 4000.0)))
                rid += 1
        bams[s] = os.path.join(tmp, f'{s}.bam')
        SF.write_bam(bams[s], refs, reads, block_bytes=60000)
        dbs[s] = os.path.join(tmp, f'{s}.db')
        SF.write_db(dbs[s], [(n, i + 1) for i, (n, _) in enumerate(refs)], [(f'r{i}', i, 0.0) for i in range(rid)], signal)
    return refs, bams, dbs


def config(paths, key):
    return PathConfig(dict(data={'kd': {'kd1': {key: paths['kd1']}, 'kd2': {key: paths['kd2']}},
                                 'wt': {'wt1': {key: paths['wt1']}, 'wt2': {key: paths['wt2']}}},
                           depleted_condition='kd', kit='RNA002'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reads', type=int, default=400, help='reads per transcript (all samples)')
    ap.add_argument('--positions', type=int, default=2000)
    ap.add_argument('--transcripts', type=int, default=6)
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        refs, bams, dbs = make_files(tmp, args.transcripts, args.positions, args.reads // 4)
        cores = os.cpu_count()
        for kind, paths in (('uncalled4_bam', bams), ('sqlite_db', dbs)):
            cfg = config(paths, 'bam' if kind == 'uncalled4_bam' else 'db')
            sids = cfg.get_sample_ids()
            s2c, cids = cfg.sample_to_condition(), cfg.get_condition_ids()
            cond_of = {s: cids[s2c[s]] for s in sids}
            # ---- native
            t0 = time.perf_counter()
            reader = R.Uncalled4Reader(cfg) if kind == 'uncalled4_bam' else R.DbReader(cfg)
            t_open = time.perf_counter() - t0
            t0 = time.perf_counter()
            rows = [reader.read(name, P) for name, P in refs]
            t_read = time.perf_counter() - t0
            t0 = time.perf_counter()
            batch, _, _ = R.pack_rows(rows, pin=False)
            t_pack = time.perf_counter() - t0
            entries = batch.n_entries
            # ---- restated reference steps
            parse = 0.0
            t0 = time.perf_counter()
            dense = []
            if kind == 'uncalled4_bam':
                tp = time.perf_counter()
                shim = {s: bam_shim.AlignmentFile(p, 'rb') for s, p in paths.items()}
                parse = time.perf_counter() - tp
                for name, P in refs:
                    data, reads, samples = OR.uncalled4_get_data(name, P, shim, 'RNA002')
                    order = np.argsort(reads, kind='stable')
                    data, samples = data[:, order, :], samples[order]
                    ok = OR.reads_invalid_ratio(data[:, :, 0]) < 0.1
                    dense.append((data[:, ok], np.array([sids[s] for s in samples])[ok],
                                  np.array([cond_of[s] for s in samples])[ok]))
            else:
                for name, P in refs:
                    dense.append(OR.read_data_db(name, paths, sids, cond_of))
            t_ref_read = time.perf_counter() - t0 - parse
            t0 = time.perf_counter()
            ref_batch, _, _ = pack_csr(dense)
            t_ref_pack = time.perf_counter() - t0
            assert ref_batch.n_entries == entries
            assert np.array_equal(ref_batch.pos_offsets.numpy(), batch.pos_offsets.numpy())
            native = t_read + t_pack
            print(json.dumps({
                "format": kind, "transcripts": len(refs), "positions": args.positions, "reads": args.reads,
                "entries": entries, "host_cores": cores,
                "native": {"open_index_s": t_open, "read_s": t_read, "pack_s": t_pack,
                           "entries_per_s": entries / native, "MB_per_s_of_csr": entries * 9 / native / 1e6},
                "restated_reference": {"pure_python_bam_parse_s": parse, "read_dense_tensor_s": t_ref_read,
                                       "dense_pack_s": t_ref_pack,
                                       "entries_per_s_without_parse": entries / (t_ref_read + t_ref_pack)},
                "speedup_without_parse": (t_ref_read + t_ref_pack) / native}), flush=True)


if __name__ == '__main__':
    main()
