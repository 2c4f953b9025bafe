#!/usr/bin/env python
"""
bench.py -- tested positions/sec of the per-position comparison hot path (GMM + KS).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path

Workload = BASELINE.json configs[2] ("config 3"): synthetic transcripts of 2,000 positions x
2x2 replicates x 100 reads per condition, comparison_methods [GMM, KS], the coverage filter
(min_coverage 30) and log10(dwell) inside the timed path.  One step = one batch of
``--tx-per-step`` transcripts (default 50 -> 100,000 positions, ~180 MB of reads: larger than
the 126 MB L2, and every step reads a different batch).  K = 20 steps is the whole config.

Printed JSON (one line, rank 0):
  value     positions/s with the CSR batches already resident in HBM (device timed, CUDA events),
            batches issued round-robin on ``--slots`` engines/streams
  e2e       positions/s through the public API with pinned HOST batches: H2D copy of the reads,
            kernels and D2H copy of the result columns inside the timed region (``--slots`` engines on
            as many streams so that the copies of one batch overlap the kernels of the others); wall
            clock, K steps timed three times over the same batches, the median pass reported
  roofline  HBM roofline of the dominant kernel (position kernel), live CUDA-event timing
  cpu_baseline  the oracle port (oracle/comparator.py: numpy + SciPy per position, as the
            reference does) on one host core, on a bounded sample of the same workload

``--impl reference`` times that same CPU port on ALL host cores (one process per core over
transcripts, as the reference's ``devices: {cpu: C}`` does, run.py:104-131).  The reference is
pure Python and its GMM lives in the absent third-party package gmm-gpu, so the CPU arm is the
port: "reference algorithm, restated gmm_gpu".
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIG_INDEX = 3
N_POSITIONS = 2000
READS_PER_COND = 100
METHODS = ['GMM', 'KS']
MIN_COVERAGE = 30
METRIC = "tested positions/sec (GMM+KS)"
UNIT = "positions/s"
SAMPLE_IDS = {'kd1': 0, 'kd2': 1, 'wt1': 2, 'wt2': 3}


def workload_name(tx_per_step):
    return (f"config3: transcripts of {N_POSITIONS} positions x 2x2 replicates x {READS_PER_COND} "
            f"reads/condition, [GMM, KS], min_coverage {MIN_COVERAGE}; step = {tx_per_step} transcripts")


def make_batch_transcripts(seed, n_tx):
    from nanocompore_b200.synthetic import make_transcript
    rng = np.random.default_rng(seed)
    return [make_transcript(rng, N_POSITIONS, READS_PER_COND, name=f'tx{seed}_{i}') for i in range(n_tx)]


# ----------------------------------------------------------------------------------------
# CPU arm: the oracle port
# ----------------------------------------------------------------------------------------

def cpu_compare_transcript(tx):
    """One transcript through the CPU port: prep (run.py:511-601) + compare_transcript."""
    from oracle import prep as oprep
    from oracle.comparator import OracleComparator, OracleConfig
    data, samples, conditions, positions = oprep.prepare_data(tx.data, tx.samples, tx.conditions)
    data, positions = oprep.filter_low_cov_positions(data, positions, conditions, MIN_COVERAGE)
    cmp_ = OracleComparator(OracleConfig(comparison_methods=METHODS, sample_ids=SAMPLE_IDS))
    df = cmp_.compare_transcript(tx.name, data, samples, conditions, positions)
    return 0 if df is None else len(df)


def _cpu_worker(args):
    seed, n_positions = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass
    from nanocompore_b200.synthetic import make_transcript
    rng = np.random.default_rng(seed)
    tx = make_transcript(rng, n_positions, READS_PER_COND, name=f'cpu{seed}')
    t0 = time.perf_counter()
    rows = cpu_compare_transcript(tx)
    return rows, time.perf_counter() - t0


def cpu_baseline_single_core(budget_s=20.0, n_positions=500):
    """Bounded sample on one core: slices of config-3 transcripts until ~budget_s."""
    rows, spent, n = 0, 0.0, 0
    while spent < budget_s and n < 64:
        r, dt = _cpu_worker((977 + n, n_positions))
        rows += r
        spent += dt
        n += 1
    return dict(value=rows / spent, unit=UNIT, cores=1, kind="port",
                sample=f"{n} slices of {n_positions} positions x {2 * READS_PER_COND} reads "
                       f"({rows} tested positions, {spent:.1f} s)")


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_positions = args.ref_positions
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        def step(i):
            t0 = time.perf_counter()
            out = pool.map(_cpu_worker, [(100000 + i * cores + c, n_positions) for c in range(cores)])
            return sum(r for r, _ in out), time.perf_counter() - t0
        for i in range(args.warmup):
            step(i)
        rows, total = 0, 0.0
        for i in range(args.steps):
            r, dt = step(args.warmup + i)
            rows += r
            total += dt
    value = rows / total
    sample = (f"{args.steps} steps x {cores} slices of {n_positions} positions x "
              f"{2 * READS_PER_COND} reads, one process per core")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.tx_per_step), "cpu_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={self.QUERY}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from nanocompore_b200 import _lib as L
    from nanocompore_b200.engine import Engine, Results, pack_csr

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    # pinned batches are allocated below: place this rank next to its GPU first (NCB_NUMA=0 skips)
    from nanocompore_b200.numa import bind_rank_to_gpu_node
    numa = bind_rank_to_gpu_node(local_rank, apply=os.environ.get('NCB_NUMA', '1') != '0')
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL prints its version banner on stdout when the communicator comes up; stdout must
        # carry the one JSON line only, so it points at stderr until the first collective is done
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    K, W, T = args.steps, args.warmup, args.tx_per_step
    n_batches = K + W
    opts = dict(tests=METHODS, n_samples=4, depleted_condition=0, min_coverage=MIN_COVERAGE,
                dwell_is_raw=True, cluster_counts='hard')
    NSLOT = args.slots
    engines = [Engine(local_rank, **opts) for _ in range(NSLOT)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NSLOT)]

    # ---- synthetic batches: pinned host copies (e2e) and device copies (value) ------------
    host_batches, dev_batches = [], []
    t_gen = time.perf_counter()
    for b in range(n_batches):
        txs = make_batch_transcripts(20251019 + CONFIG_INDEX + 1000 * rank + b, T)
        batch, _, _ = pack_csr([(t.data, t.samples, t.conditions) for t in txs], pin=True)
        host_batches.append(batch)
        dev_batches.append(batch.to(dev))
    t_gen = time.perf_counter() - t_gen
    P = host_batches[0].n_positions
    h2d_bytes = int(np.mean([b.nbytes() for b in host_batches[W:]]))
    entries = int(np.mean([b.n_entries for b in host_batches[W:]]))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- value: device-resident inputs, NSLOT engines round-robin on their own streams -------
    res_dev = [Results(P, 4, dev) for _ in range(NSLOT)]
    tested_slot = [torch.zeros((), dtype=torch.int64, device=dev) for _ in range(NSLOT)]
    sampler = ClockSampler(local_rank)
    sampler.start()

    def value_step(i, b):
        s = i % NSLOT
        with torch.cuda.stream(streams[s]):
            engines[s].compare_csr(dev_batches[b], res_dev[s], stream=streams[s], wait=False)
            tested_slot[s] += ((res_dev[s].status & L.NCB_POS_DROP_MASK) == 0).sum()

    for i, b in enumerate(range(W)):
        value_step(i, b)
    barrier()
    for t_ in tested_slot:
        t_.zero_()
    launches0 = sum(e.kernel_launches() for e in engines)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream(dev)
    ev0.record(main)
    for st in streams:
        st.wait_event(ev0)
    for i, b in enumerate(range(W, n_batches)):
        value_step(i, b)
    for st in streams:
        done = torch.cuda.Event()
        done.record(st)
        main.wait_event(done)
    ev1.record(main)
    barrier()
    ms_value = max_over_ranks(ev0.elapsed_time(ev1))
    launches = sum(e.kernel_launches() for e in engines) - launches0
    tested_positions = int(sum(int(t_.item()) for t_ in tested_slot))
    total_tested = sum_over_ranks(float(tested_positions))
    value = total_tested / (ms_value / 1000.0)

    # ---- roofline pass: the same batches on ONE stream, so that every kernel is timed alone --
    eng = engines[0]
    eng.set_timing(True)
    for b in range(W, min(n_batches, W + 6)):
        eng.compare_csr(dev_batches[b], res_dev[0], wait=False)
    totals = eng.timing_totals()
    eng.set_timing(False)
    fp64_peak = eng.probe_fp64_peak()   # DFMA microbenchmark on this device, TFLOP/s

    # ---- e2e: pinned host batches through the public API, copies inside the timed region ---
    res_host = [Results(P, 4, None) for _ in range(NSLOT)]
    import gc
    gc.collect()
    gc.disable()            # a collection inside the wall-clock region would be charged to the path
    for b in range(W):
        engines[b % NSLOT].compare_csr(host_batches[b], res_host[b % NSLOT], stream=streams[b % NSLOT], wait=False)
    # the wall-clock region is short (K steps = tens of milliseconds) and a single stall of the host
    # (the nvidia-smi sampler attaching, a neighbour on the box) has been seen to triple it: the K
    # steps are timed E2E_REPEATS times over the same batches and the MEDIAN pass is reported; every
    # pass is listed in the JSON line
    E2E_REPEATS = 3
    e2e_passes = []
    for _ in range(E2E_REPEATS):
        barrier()
        t0 = time.perf_counter()
        tested_host = 0
        pending = [None] * NSLOT
        for i, b in enumerate(range(W, n_batches)):
            s = i % NSLOT
            if pending[s] is not None:
                engines[s].wait(streams[s])
                tested_host += int(((res_host[s].status & L.NCB_POS_DROP_MASK) == 0).sum())
            engines[s].compare_csr(host_batches[b], res_host[s], stream=streams[s], wait=False)
            pending[s] = b
        for s in range(NSLOT):
            if pending[s] is not None:
                engines[s].wait(streams[s])
                tested_host += int(((res_host[s].status & L.NCB_POS_DROP_MASK) == 0).sum())
        torch.cuda.synchronize(dev)
        e2e_passes.append(max_over_ranks((time.perf_counter() - t0) * 1000.0))
    ms_e2e = float(np.median(e2e_passes))
    gc.enable()
    clocks = sampler.stop()
    total_tested_host = sum_over_ranks(float(tested_host))
    e2e_value = total_tested_host / (ms_e2e / 1000.0)
    d2h_bytes = res_host[0].nbytes()
    total_launches = int(sum_over_ranks(float(launches)))

    # ---- end-of-run gather of the p-value columns on rank 0 (the only collective of the job;
    # outside the timed regions: it happens once per run, not per step)
    gathered = None
    if world > 1:
        from nanocompore_b200.sharding import gather_columns
        cols = res_dev[0].pvalues[[L.PV['KS_INTENSITY'], L.PV['KS_DWELL'], L.PV['GMM_CHI2']]].contiguous()
        keys = (torch.arange(P, device=dev, dtype=torch.int64) | (rank << 32))
        tg = time.perf_counter()
        out = gather_columns(cols, keys)
        torch.cuda.synchronize(dev)
        if rank == 0:
            gathered = {"rows": int(out[0].shape[1]), "columns": int(out[0].shape[0]),
                        "ms": (time.perf_counter() - tg) * 1000.0,
                        "sorted_by_key": bool((out[1][1:] > out[1][:-1]).all().item())}
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the mixture kernel), HBM bound -------------------
    # algorithmic bytes per launch of the mixture kernel: standardised values (8 B) + label
    # (1 B) of every read once, offsets + read count (12 B) and the columns it writes
    # (chi2 p, LOR, 2 x 4 cluster counts = 44 B) per position.  The statistics kernel moves
    # 9 B/read in and 8 B/read out plus 8 + 132 B per position.
    nb = max(totals['batches'], 1)
    gmm_bytes = entries * 9 + P * (12 + 44)
    stats_bytes = entries * (9 + 8) + P * (8 + 132)
    ms_gmm, ms_stats = totals['ms_gmm'] / nb, totals['ms_stats'] / nb
    peak, peak_src = measured_peaks()
    achieved = gmm_bytes / (ms_gmm / 1000.0) / 1e9
    roofline = {"bound": "hbm", "kernel": "mixture kernels of a batch: ncb::position_kernel_warp<256, 8, 3, 1> (79 % of their time) + <128, 8, 2, 1>",
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None,
                "algorithmic_bytes_per_launch": gmm_bytes, "ms_per_launch": ms_gmm,
                "statistics_kernel": {"kernel": "statistics kernels of a batch: ncb::position_kernel_warp<256, 8, 3, 0> + <128, 8, 2, 0>",
                                      "algorithmic_bytes_per_launch": stats_bytes, "ms_per_launch": ms_stats,
                                      "achieved": stats_bytes / (ms_stats / 1000.0) / 1e9,
                                      "frac": stats_bytes / (ms_stats / 1000.0) / 1e9 / peak},
                "ms_ks_pvalue_kernels": totals['ms_ks'] / nb,
                "ms_classify": totals['ms_classify'] / nb,
                "note": "both kernels are issue / fp64-pipe bound (fp64 EM, sort), not HBM bound: "
                        "DRAM traffic equals the algorithmic bytes and is ~1 % of peak; the fp64 entry "
                        "is the compute roofline of the mixture kernel; see DESIGN.md"}
    # compute roofline of the same kernel: fp64 operations per position counted by ncu
    # (profiles/fp64_ops.json, thread-level DFMA/DADD/DMUL of the --set full capture) x the
    # positions of a launch / the live launch time, against the fp64 rate measured by
    # ncb_probe_fp64_peak in this run
    ops_path = os.path.join(ROOT, 'profiles', 'fp64_ops.json')
    fp64 = {"peak": fp64_peak, "unit": "TFLOP/s", "peak_source": "measured (DFMA microbenchmark, this run)",
            "achieved": None, "frac": None}
    if os.path.exists(ops_path):
        try:
            with open(ops_path) as f:
                ops = json.load(f)
            flop = (2.0 * ops["dfma_per_position"] + ops["dadd_per_position"] + ops["dmul_per_position"]) * P
            slots = (ops["dfma_per_position"] + ops["dadd_per_position"] + ops["dmul_per_position"]) * P
            fp64["achieved"] = flop / (ms_gmm / 1000.0) / 1e12
            fp64["frac"] = fp64["achieved"] / fp64_peak
            # share of the fp64 pipe's issue slots (an add or a multiply takes the slot of an FMA)
            fp64["pipe_frac"] = 2.0 * slots / (ms_gmm / 1000.0) / 1e12 / fp64_peak
            fp64["flop_per_position"] = flop / P
            fp64["ops_source"] = ops.get("source")
        except Exception:
            pass
    roofline["fp64"] = fp64
    ncu_traffic = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(ncu_traffic):
        try:
            with open(ncu_traffic) as f:
                tr = json.load(f)
            # ncu capture at tr['positions'] positions per launch: scale per position
            roofline["traffic"] = tr["dram_bytes_per_position"] * P
            roofline["traffic_source"] = tr.get("source")
        except Exception:
            pass

    cpu = cpu_baseline_single_core(args.cpu_budget) if not args.no_cpu_baseline else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_value / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(T), "positions_per_step": P,
                   "tested_positions_per_step": tested_positions / K,
                   "reads_per_step": entries, "l2": "inputs larger than L2, distinct batch per step", "slots": NSLOT,
                   "parallelism": f"transcript sharding x{world}, no data-path collective"},
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / K,
                "passes_ms_per_step": [m / K for m in e2e_passes], "reported": "median pass"},
        "gpu_launches": total_launches,
        "final_gather": gathered,
        "clocks": clocks,
        "numa": numa,
        "setup_s": {"generate_and_pack": t_gen},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--tx-per-step', type=int, default=50)
    ap.add_argument('--cpu-budget', type=float, default=20.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--slots', type=int, default=4,
                    help="engines/streams of the end-to-end pipeline (copy-in, kernels, copy-out overlap)")
    ap.add_argument('--ref-positions', type=int, default=250,
                    help="positions per slice of the reference arm (one slice per core per step)")
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
