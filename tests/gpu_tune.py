"""Times builds of libncb (build/tune/*.so) on the config-3 batch shape; tuning helper.

    python -m nanocompore_b200.build -DNCB_ACC_FOLD=1 --out=build/tune/fold.so     # a variant
    python tests/gpu_tune.py [n_transcripts] [lib ...]                              # default: build/tune/*.so

One JSON line per library: milliseconds per batch of the statistics kernels, the mixture kernels and
the exact-KS kernels (CUDA events, timed batches run on one stream), the milliseconds per batch when
batches are in flight on 4 handles / streams (what bench.py's `value` measures: the end of one batch's
kernels overlaps the start of the next's), and whether the sum of all p-values is bit-identical to the
first library's.  NCB_TUNE_FP32=1: float32 E-step.
"""
import glob
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_transcripts  # noqa: E402
from nanocompore_b200.engine import Engine, Results, pack_csr  # noqa: E402

n_tx = int(sys.argv[1]) if len(sys.argv) > 1 else 25
libs = sys.argv[2:] or sorted(glob.glob(os.path.join(ROOT, 'build', 'tune', '*.so')))
batches = []
for b in range(3):
    txs = make_transcripts(555 + b, n_tx)
    batches.append(pack_csr([(t.data, t.samples, t.conditions) for t in txs], wire='i16')[0].to('cuda'))
P = batches[0].n_positions
ref = None
for lib in libs:
    eng = Engine(0, tests=('GMM', 'KS'), min_coverage=30, dwell_is_raw=True, lib_path=lib,
                 **({'em_fp64': False} if os.environ.get('NCB_TUNE_FP32') == '1' else {}))
    res = Results(P, 4, torch.device('cuda'))
    for b in batches:
        eng.compare_csr(b, res, wait=False)
    eng.set_timing(True)
    for _ in range(2):
        for b in batches:
            eng.compare_csr(b, res, wait=False)
    t = eng.timing_totals()
    chk = torch.nan_to_num(res.pvalues, nan=2.0).sum().item()
    ref = chk if ref is None else ref
    rec = dict(lib=os.path.basename(lib), ms_position=t['ms_position'] / t['batches'],
               ms_stats=t['ms_stats'] / t['batches'], ms_gmm=t['ms_gmm'] / t['batches'],
               ms_ks=t['ms_ks'] / t['batches'], positions=P, same=bool(chk == ref))
    # steady state: 4 handles / streams, batches round-robin
    kw = dict(tests=('GMM', 'KS'), min_coverage=30, dwell_is_raw=True, lib_path=lib)
    if os.environ.get('NCB_TUNE_FP32') == '1':
        kw['em_fp64'] = False
    engs = [Engine(0, **kw) for _ in range(4)]
    streams = [torch.cuda.Stream() for _ in range(4)]
    outs = [Results(P, 4, torch.device('cuda')) for _ in range(4)]
    def sweep(rounds):
        for r in range(rounds):
            for i, b in enumerate(batches):
                k = (r * len(batches) + i) % 4
                engs[k].compare_csr(b, outs[k], stream=streams[k], wait=False)
    sweep(2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in streams:
        st.wait_event(e0)
    rounds = 6
    sweep(rounds)
    for st in streams:
        e = torch.cuda.Event()
        e.record(st)
        torch.cuda.current_stream().wait_event(e)
    e1.record()
    torch.cuda.synchronize()
    rec['ms_batch_4streams'] = e0.elapsed_time(e1) / (rounds * len(batches))
    print(json.dumps(rec), flush=True)
    for e in engs:
        e.close()
    eng.close()
