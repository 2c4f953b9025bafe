"""Prints EM iteration statistics of the config-3 workload (helper for DESIGN.md / profiles)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_batch_transcripts
from nanocompore_b200 import _lib as L
from nanocompore_b200.engine import Engine, pack_csr

txs = make_batch_transcripts(424242, 5)
batch, _, _ = pack_csr([(t.data, t.samples, t.conditions) for t in txs], pin=True)
eng = Engine(0, tests=('GMM', 'KS'), min_coverage=30, dwell_is_raw=True)
res = eng.compare_csr(batch.to('cuda'), diagnostics=True).numpy()
ok = (res['status'] & L.NCB_POS_DROP_MASK) == 0
it = res['diag_i64'][L.DI['EM_ITERS']][ok]
ngmm = res['diag_i64'][L.DI['N_GMM']][ok]
sig = ~np.isnan(res['pvalues'][L.PV['GMM_CHI2']][ok])
print(json.dumps(dict(tested=int(ok.sum()), em_iters_mean=float(it.mean()), em_iters_median=float(np.median(it)),
                      em_iters_p90=float(np.percentile(it, 90)), em_iters_max=int(it.max()),
                      frac_at_max_iter=float((it >= 100).mean()), reads_mean=float(ngmm.mean()),
                      frac_two_component=float(sig.mean()))))
