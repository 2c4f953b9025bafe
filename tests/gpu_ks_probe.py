"""Times the exact-KS unit entry on the (n0, n1, h) triples of a high-coverage shard; tuning helper."""
import json, math, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanocompore_b200 import _lib as L
from nanocompore_b200.engine import Engine, Results, pack_csr
from nanocompore_b200.synthetic import make_transcript, BASE_SEED

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
npos = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rng = np.random.default_rng(BASE_SEED + 5)
txs = [make_transcript(rng, npos, reads) for _ in range(2)]
batch = pack_csr([(t.data, t.samples, t.conditions) for t in txs])[0].to('cuda')
eng = Engine(0, tests=('KS',), min_coverage=30, dwell_is_raw=True)
res = eng.compare_csr(batch, diagnostics=True).numpy()
d = res['diag_i64']
n0 = np.concatenate([d[L.DI['N0_INTENSITY']], d[L.DI['N0_DWELL']]]).astype(np.int32)
n1 = np.concatenate([d[L.DI['N1_INTENSITY']], d[L.DI['N1_DWELL']]]).astype(np.int32)
h = np.concatenate([d[L.DI['KS_H_INTENSITY']], d[L.DI['KS_H_DWELL']]]).astype(np.int64)
ok = (n0 > 0) & (n1 > 0)
n0, n1, h = n0[ok], n1[ok], h[ok]
g = np.gcd(n0, n1)
m, n = np.maximum(n0, n1), np.minimum(n0, n1)
mg = m // g
lcm = mg.astype(np.float64) * n
D = h / lcm
width = np.minimum(2.0 * h / mg + 1, n + 1)
skip = 2 * D * D * m.astype(np.float64) * n / (m + n) > 921
cells = np.where(skip | (n0 == n1), 0, m * width)
print(json.dumps(dict(problems=int(ok.sum()), equal=int((n0 == n1).sum()), skipped=int(skip.sum()),
                      n_median=float(np.median(n)), D_median=float(np.median(D)), width_median=float(np.median(width)),
                      cells_total=float(cells.sum()), cells_max=float(cells.max()))))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = eng.ks_exact_pvalues(n0, n1, h)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(json.dumps(dict(ks_unit_ms=dt * 1e3, cells_per_s=float(cells.sum()) / dt, p_min=float(np.nanmin(p)), p_max=float(np.nanmax(p)))))
