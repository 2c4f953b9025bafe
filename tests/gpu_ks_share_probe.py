"""Times builds of libncb with different CTA / warp split thresholds (build/tune/share*.so, -DKS_CTA_SHARE=...)
on exact-KS problem sets of high-coverage shards; tuning helper.  Measured: 3072 and 8192 tie, 128-1024 and
32768 lose 15-45 %."""
import glob, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanocompore_b200 import _lib as L
from nanocompore_b200.engine import Engine, pack_csr
from nanocompore_b200.synthetic import make_transcript, BASE_SEED
sets = {}
for reads, npos in ((5000, 250), (5000, 2500), (1000, 400), (1000, 5000), (400, 10000)):
    rng = np.random.default_rng(BASE_SEED + 5)
    txs = [make_transcript(rng, npos, reads) for _ in range(2)]
    batch = pack_csr([(t.data, t.samples, t.conditions) for t in txs])[0].to('cuda')
    eng = Engine(0, tests=('KS',), min_coverage=30, dwell_is_raw=True)
    d = eng.compare_csr(batch, diagnostics=True).numpy()['diag_i64']
    n0 = np.concatenate([d[L.DI['N0_INTENSITY']], d[L.DI['N0_DWELL']]]).astype(np.int32)
    n1 = np.concatenate([d[L.DI['N1_INTENSITY']], d[L.DI['N1_DWELL']]]).astype(np.int32)
    h = np.concatenate([d[L.DI['KS_H_INTENSITY']], d[L.DI['KS_H_DWELL']]]).astype(np.int64)
    ok = (n0 > 0) & (n1 > 0)
    sets[(reads, npos)] = (n0[ok], n1[ok], h[ok])
    eng.close()
    del batch
for lib in sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'build', 'tune', 'share*.so'))):
    eng = Engine(0, tests=('KS',), lib_path=lib)
    out = {}
    for key, (n0, n1, h) in sets.items():
        for _ in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            eng.ks_exact_pvalues(n0, n1, h)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        out[f'{key[0]}x{key[1]}'] = round(dt * 1e3, 2)
    print(os.path.basename(lib), json.dumps(out), flush=True)
    eng.close()
