"""Times a shard of BASELINE config 5 (5000 reads/condition + low-coverage tail); tuning helper."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanocompore_b200.engine import Engine, Results, pack_csr
from nanocompore_b200.synthetic import make_transcript, make_skew_tail, BASE_SEED

rng = np.random.default_rng(BASE_SEED + 5)
high = [make_transcript(rng, 100, 5000) for _ in range(2)]
tail = make_skew_tail(rng, 50, n_positions=1000)
for name, txs in (('high', high), ('tail', tail), ('mixed', high + tail)):
    batch = pack_csr([(t.data, t.samples, t.conditions) for t in txs])[0].to('cuda')
    eng = Engine(0, tests=('GMM', 'KS'), min_coverage=30, dwell_is_raw=True)
    res = Results(batch.n_positions, 4, torch.device('cuda'))
    eng.compare_csr(batch, res)
    eng.set_timing(True)
    for _ in range(3):
        eng.compare_csr(batch, res, wait=False)
    t = eng.timing_totals()
    nb = t['batches']
    print(json.dumps(dict(case=name, positions=batch.n_positions, entries=batch.n_entries,
                          ms_stats=t['ms_stats'] / nb, ms_gmm=t['ms_gmm'] / nb, ms_ks=t['ms_ks'] / nb,
                          positions_per_s=batch.n_positions / ((t['ms_stats'] + t['ms_gmm'] + t['ms_ks']) / nb / 1e3))), flush=True)
    eng.close()
