"""Throughput of the other BASELINE configs (2, 4, 5) on one GPU: supplementary to bench.py, which
times config 3.  Batches are device resident; each config runs a shard sized to fit a few seconds of
set-up and is timed with CUDA events over `reps` passes on 2 engines / streams.

    python tests/gpu_configs_bench.py            # one JSON line per config
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanocompore_b200 import _lib as L  # noqa: E402
from nanocompore_b200.engine import Engine, Results, pack_csr  # noqa: E402
from nanocompore_b200.synthetic import BASE_SEED, make_skew_tail, make_transcript  # noqa: E402

dev = torch.device('cuda', 0)


def run(name, batches, tests, reps=3, fdr=False):
    engines = [Engine(0, tests=tests, min_coverage=30, dwell_is_raw=True) for _ in range(2)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    dbs = [b.to(dev) for b in batches]
    res = [Results(b.n_positions, 4, dev) for b in dbs]
    def one_pass():
        for i, b in enumerate(dbs):
            s = i % 2
            with torch.cuda.stream(streams[s]):
                engines[s].compare_csr(b, res[i], stream=streams[s], wait=False)
    one_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream(dev)
    e0.record(main)
    for st in streams:
        st.wait_event(e0)
    for _ in range(reps):
        one_pass()
    for st in streams:
        ev = torch.cuda.Event()
        ev.record(st)
        main.wait_event(ev)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    positions = sum(b.n_positions for b in dbs)
    tested = sum(int(((r.status & L.NCB_POS_DROP_MASK) == 0).sum()) for r in res)
    out = dict(config=name, tests=list(tests), positions=positions, tested_positions=tested,
               entries=sum(b.n_entries for b in dbs), ms_per_pass=ms, tested_positions_per_s=tested / (ms / 1e3))
    if fdr:   # config 4: Benjamini-Hochberg over the three p-value columns of the shard, on the GPU
        cols = torch.cat([torch.cat([r.pvalues[L.PV[k]] for r in res]) for k in ('MW_INTENSITY', 'MW_DWELL', 'GMM_CHI2')])
        engines[0].bh_fdr(cols)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        q = engines[0].bh_fdr(cols)
        torch.cuda.synchronize()
        out.update(fdr_values=int(cols.numel()), fdr_ms=(time.perf_counter() - t0) * 1e3,
                   fdr_significant=int((q < 0.05).sum()))
    for e in engines:
        e.close()
    print(json.dumps(out), flush=True)


def packed(txs):
    return pack_csr([(t.data, t.samples, t.conditions) for t in txs])[0]


rng = np.random.default_rng(BASE_SEED + 2)
run('config2: 100 tx x 1000 pos x 50 reads/cond (whole config)',
    [packed([make_transcript(rng, 1000, 50) for _ in range(25)]) for _ in range(4)], ('KS',))
rng = np.random.default_rng(BASE_SEED + 4)
run('config4 shard: 200 tx x 1000 pos x 150 reads/cond',
    [packed([make_transcript(rng, 1000, 150) for _ in range(50)]) for _ in range(4)], ('GMM', 'MW'), fdr=True)
rng = np.random.default_rng(BASE_SEED + 5)
high = [packed([make_transcript(rng, 250, 5000)]) for _ in range(2)]
tail = [packed(make_skew_tail(rng, 25, n_positions=1000)) for _ in range(2)]
run('config5 shard: 2 tx x 250 pos x 5000 reads/cond + 50 low-coverage tx x 1000 pos', high + tail, ('GMM', 'KS'), reps=2)
# the same mix at the config's own proportions (1 high-coverage transcript per 25 of the tail) and
# whole 1000-position transcripts: enough exact-KS problems of ~10^6 lattice cells to fill the GPU
if '--large' in sys.argv:
    rng = np.random.default_rng(BASE_SEED + 55)
    high = [packed([make_transcript(rng, 1000, 5000)]) for _ in range(8)]
    tail = [packed(make_skew_tail(rng, 25, n_positions=1000)) for _ in range(8)]
    run('config5 shard at config proportions: 8 tx x 1000 pos x 5000 reads/cond + 200 low-coverage tx x 1000 pos',
        [b for pair in zip(high, tail) for b in pair], ('GMM', 'KS'), reps=1)
