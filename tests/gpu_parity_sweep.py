"""Wide parity sweep (not part of the test-suite): many seeded transcripts through the CUDA path
and through the oracle (a process pool on the host cores), every check of
tests/test_gpu_parity.py:check_against_oracle per transcript.  Prints a JSON summary.

    python tests/gpu_parity_sweep.py [n_transcripts] [positions] [processes] [large]

``large`` sweeps the CTA size classes (513-10 240 entries per position) and with them the wavefront
forms of the exact-KS kernels: 300-5 000 reads per condition, ragged (3'-anchored read spans).
"""
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

SHAPES = [(40, False), (100, False), (100, True), (150, False), (31, True), (230, False), (12, True)]
SHAPES_MOTOR = [(40, 0.08), (120, 0.03), (31, 0.08), (230, 0.03), (100, 0.10), (60, 0.0), (300, 0.06), (700, 0.04)]
SHAPES_CSR = [(40, False), (100, True), (150, False), (230, False), (505, True), (1100, False), (12, True),
              (2600, False)]
SHAPES_LARGE = [(300, False), (505, True), (700, False), (1100, False), (2040, True), (2600, False), (4000, False),
                (5000, False)]


MOTOR = len(sys.argv) > 4 and sys.argv[4] == 'motor'
CSR = len(sys.argv) > 4 and sys.argv[4] == 'csr'


def job_inputs(seed, P, n, q):
    from nanocompore_b200.synthetic import make_transcript
    from oracle import prep as oprep
    if MOTOR:      # q is the gap rate
        tx = make_transcript(np.random.default_rng(seed), P, n, gap_rate=q)
        return oprep.prepare_data(tx.data, tx.samples, tx.conditions, 7)
    tx = make_transcript(np.random.default_rng(seed), P, n, quantised=q)
    return oprep.prepare_data(tx.data, tx.samples, tx.conditions)


def gpu_results(jobs):
    import torch
    from nanocompore_b200.engine import Engine, make_labels
    eng = Engine(0, tests=('KS', 'MW', 'TT', 'GMM'), with_logit=not MOTOR, **(dict(min_coverage=0) if MOTOR else {}))
    out = []
    if CSR:
        from nanocompore_b200.engine import pack_csr
        for b in range(0, len(jobs), 8):
            txs = [job_inputs(*j)[:3] for j in jobs[b:b + 8]]
            batch, ptx, _ = pack_csr(txs, pin=True)
            host = eng.compare_csr(batch, diagnostics=True).numpy()
            for ti in range(len(txs)):
                sel = ptx == ti
                out.append({k: (np.ascontiguousarray(v[..., sel]) if v is not None else None) for k, v in host.items()})
        eng.close()
        return out
    for seed, P, n, q in jobs:
        data, samples, conditions, _ = job_inputs(seed, P, n, q)
        res = eng.compare_dense(torch.from_numpy(data).cuda(), torch.from_numpy(make_labels(samples, conditions)).cuda(),
                                diagnostics=True).numpy()
        out.append(res)
    eng.close()
    return out


def check(args):
    (seed, P, n, q), res = args
    os.environ['CUDA_VISIBLE_DEVICES'] = ''
    import test_gpu_parity as T
    data, samples, conditions, positions = job_inputs(seed, P, n, q)
    try:
        if MOTOR:
            ok, _ = T.check_three_variables(res, data, samples, conditions, positions)
        else:
            ok = T.check_against_oracle(res, data, samples, conditions)
        return seed, int(ok.sum()), None
    except AssertionError as e:
        return seed, 0, f"seed {seed} n {n} {'gap rate' if MOTOR else 'quantised'} {q}: {str(e)[:400]}"


def main():
    n_tx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    procs = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
    large = len(sys.argv) > 4 and sys.argv[4] == 'large'
    shapes = SHAPES_LARGE if large else SHAPES_MOTOR if MOTOR else SHAPES_CSR if CSR else SHAPES
    jobs = [((19000 if large else 29000 if MOTOR else 39000 if CSR else 9000) + i, P) + shapes[i % len(shapes)] for i in range(n_tx)]
    t0 = time.perf_counter()
    res = gpu_results(jobs)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=procs) as ex:
        outcomes = list(ex.map(check, zip(jobs, res)))
    failures = [msg for _, _, msg in outcomes if msg]
    print(json.dumps(dict(transcripts=n_tx, reads_per_condition=sorted({n for n, _ in shapes}), positions_checked=sum(k for _, k, _ in outcomes), failures=len(failures),
                          first_failures=failures[:5], gpu_s=t_gpu, oracle_s=time.perf_counter() - t0)), flush=True)
    sys.exit(1 if failures else 0)


if __name__ == '__main__':
    main()
