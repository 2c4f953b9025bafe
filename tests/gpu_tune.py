"""Times builds of libncb (build/tune/*.so) on the config-3 batch shape; tuning helper.

    python -m nanocompore_b200.build -DNCB_ACC_FOLD=1 --out=build/tune/fold.so     # a variant
    python tests/gpu_tune.py [n_transcripts] [lib ...]                              # default: build/tune/*.so

One JSON line per library: milliseconds per batch of the statistics kernels, the mixture kernels and
the exact-KS kernels (CUDA events, timed batches run on one stream), and whether the sum of all
p-values is bit-identical to the first library's.
"""
import glob
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_batch_transcripts  # noqa: E402
from nanocompore_b200.engine import Engine, Results, pack_csr  # noqa: E402

n_tx = int(sys.argv[1]) if len(sys.argv) > 1 else 25
libs = sys.argv[2:] or sorted(glob.glob(os.path.join(ROOT, 'build', 'tune', '*.so')))
batches = []
for b in range(3):
    txs = make_batch_transcripts(555 + b, n_tx)
    batches.append(pack_csr([(t.data, t.samples, t.conditions) for t in txs])[0].to('cuda'))
P = batches[0].n_positions
ref = None
for lib in libs:
    eng = Engine(0, tests=('GMM', 'KS'), min_coverage=30, dwell_is_raw=True, lib_path=lib)
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
    print(json.dumps(dict(lib=os.path.basename(lib), ms_position=t['ms_position'] / t['batches'],
                          ms_stats=t['ms_stats'] / t['batches'], ms_gmm=t['ms_gmm'] / t['batches'],
                          ms_ks=t['ms_ks'] / t['batches'], positions=P, same=bool(chk == ref))), flush=True)
    eng.close()
