"""High-coverage positions (BASELINE config 5: 5000 reads/condition): where the time goes, and the
rate of the exact-KS wavefront kernels in lattice cells per second.  Tuning helper, not a test.

    python tests/gpu_ks_highcov_probe.py [transcripts] [reads per condition] [positions per transcript]"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanocompore_b200 import _lib as L
from nanocompore_b200.engine import Engine, Results, pack_csr
from nanocompore_b200.synthetic import make_transcript, BASE_SEED

n_tx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reads = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
n_pos = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
rng = np.random.default_rng(BASE_SEED + 5)
txs = [make_transcript(rng, n_pos, reads) for _ in range(n_tx)]
batch = pack_csr([(t.data, t.samples, t.conditions) for t in txs])[0].to('cuda')
eng = Engine(0, tests=('GMM', 'KS'), min_coverage=30, dwell_is_raw=True)
res = Results(batch.n_positions, 4, torch.device('cuda'), diagnostics=True)
eng.compare_csr(batch, res, diagnostics=True)
eng.set_timing(True)
for _ in range(2):
    eng.compare_csr(batch, res, diagnostics=True, wait=False)
t = eng.timing_totals()
nb = t['batches']
out = res.numpy()
di = out['diag_i64']
n0 = np.concatenate([di[L.DI['N0_INTENSITY']], di[L.DI['N0_DWELL']]]).astype(np.int64)
n1 = np.concatenate([di[L.DI['N1_INTENSITY']], di[L.DI['N1_DWELL']]]).astype(np.int64)
h = np.concatenate([di[L.DI['KS_H_INTENSITY']], di[L.DI['KS_H_DWELL']]]).astype(np.int64)
ok = (n0 > 0) & (n1 > 0) & (h > 0) & (n0 != n1)
m, n = np.maximum(n0, n1)[ok], np.minimum(n0, n1)[ok]
g = np.gcd(m, n)
mg = m // g
lcm = mg * n
d = h[ok] / lcm
skipped = 2.0 * d * d * m * n / (m + n) > 921.0
width = np.minimum(2.0 * h[ok] / mg + 1.0, n + 1.0)
cells = (m * width)[~skipped]
# the KS kernels alone, through the unit entry point
t0 = time.perf_counter()
p = eng.ks_exact_pvalues(n0.astype(np.int32), n1.astype(np.int32), h)
t_unit = time.perf_counter() - t0
t0 = time.perf_counter()
p = eng.ks_exact_pvalues(n0.astype(np.int32), n1.astype(np.int32), h)
t_unit = time.perf_counter() - t0
print(json.dumps(dict(reads_per_condition=reads, positions=batch.n_positions, entries=batch.n_entries,
                      ms_stats=t['ms_stats'] / nb, ms_gmm=t['ms_gmm'] / nb, ms_ks=t['ms_ks'] / nb,
                      problems=int(ok.sum()), equal_sizes=int(((n0 == n1) & (n0 > 0)).sum()),
                      skipped_underflow=int(skipped.sum()), walked=int(cells.size),
                      mean_cells=float(cells.mean()), max_cells=float(cells.max()), total_cells=float(cells.sum()),
                      median_width=float(np.median(width[~skipped])),
                      cells_per_s=float(cells.sum() / (t['ms_ks'] / nb / 1e3)),
                      unit_call_ms=t_unit * 1e3)), flush=True)
eng.close()
