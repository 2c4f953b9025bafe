#!/usr/bin/env python
"""
bench.py -- tested positions/sec of the per-position comparison hot path (GMM + KS).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path

Workload = BASELINE.json configs[2] ("config 3"): synthetic transcripts of 2,000 positions x 2x2
replicates x 100 reads per condition, comparison_methods [GMM, KS], the coverage filter (min_coverage 30)
and log10(dwell) inside the timed path.  The values are what the Uncalled4 reader hands over (int16 tag
values in a float32 tensor, uncalled4.py:118-192; synthetic.make_transcript(uncalled4=True)), so the
batches travel as NCB_LAYOUT_CSR_I16_RUNS (int16 entries + the label runs of every position: 4 bytes per
entry; the same entries with a label byte each and as float32 are timed beside it).  One step = ``--batches-per-step`` batches of
``--tx-per-batch`` transcripts (default 32 x 50 -> 3.2 M positions: the timed region of K = 20 steps lasts
~1.5 s, the end-to-end job ~1.6 s per pass, three passes), taken round-robin from a pool of distinct pinned batches (a batch is ~68 MB, the pool ~0.8 GB: every
launch reads data that is not in the 126 MB L2).

Printed JSON (one line, rank 0):
  value     positions/s with the CSR batches already resident in HBM (device timed, CUDA events),
            batches issued round-robin on ``--slots`` engines/streams; weak scaling (every rank runs
            K steps of its own shard)
  e2e       positions/s of the whole JOB through the public API with pinned HOST batches: per batch the
            H2D copy of the reads, the kernels and the D2H copy of the result columns (``--slots``
            engines on as many streams, so the copies of one batch overlap the kernels of the others),
            the p-value columns kept for the correction, then -- still inside the timed region -- the
            Benjamini-Hochberg correction over the rows of ALL ranks (sharding.fdr_columns: a sample sort,
            every rank corrects one value range of every column, q-values back), the q-values copied to the host, and the
            gather of keys + q-values on rank 0 sorted by key (sharding.gather_columns).  Wall clock,
            max over ranks, the job timed three times and the MEDIAN pass reported.  The transcripts of
            the job form ONE global list; with several ranks its batches are handed out by a queue the
            ranks share (sharding.TaskQueue, list scheduling: a rank the host serves faster runs more of
            them; ``schedule`` on the line) -- ``--schedule static`` runs sharding.lpt_partition's shards.
  roofline  HBM roofline of the kernel group of a batch that takes longest (live CUDA-event timing of the
            statistics kernels and of each of the three mixture kernels; `kernels` lists all of them,
            `mixture_fit` the three mixture kernels as one), `traffic` from the committed ncu capture; the
            fp64 entries are the compute rooflines of the mixture kernels
  cpu_baseline  the reference's own code (baseline/_ref: comparisons.TranscriptComparator + the Worker
            prep functions, gmm_gpu restated by oracle/gmm.py) on one host core, bounded sample
  other_configs (N = 1)  BASELINE configs 2, 4 and 5 on the same GPU: value, e2e, roofline terms and a
            parity spot check against the oracle
  job_skew  a FIXED global job of the config-5 mix (high-coverage transcripts + low-coverage tail), LPT
            sharded over the N ranks and timed end to end: the strong-scaling figure

``--impl reference`` times the reference's code on ALL host cores (one process per core over
transcripts, as the reference's ``devices: {cpu: C}`` does, run.py:104-131), full 2,000-position
transcripts of the same generator.  Its GMM lives in the absent third-party package gmm-gpu, restated by
oracle/gmm.py: "reference code, restated gmm_gpu".  The numpy port (oracle/comparator.py) is timed next to
it as a second field.
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
PV_COLUMN = {'KS_intensity_pvalue': 'KS_INTENSITY', 'KS_dwell_pvalue': 'KS_DWELL',
             'MW_intensity_pvalue': 'MW_INTENSITY', 'MW_dwell_pvalue': 'MW_DWELL',
             'TT_intensity_pvalue': 'TT_INTENSITY', 'TT_dwell_pvalue': 'TT_DWELL',
             'GMM_chi2_pvalue': 'GMM_CHI2'}


def workload_name(tx_per_batch, batches_per_step):
    return (f"config3: transcripts of {N_POSITIONS} positions x 2x2 replicates x {READS_PER_COND} "
            f"reads/condition, [GMM, KS], min_coverage {MIN_COVERAGE}, Uncalled4-like int16 values; "
            f"step = {batches_per_step} batches x {tx_per_batch} transcripts")


def make_transcripts(seed, n_tx, n_positions=N_POSITIONS, reads=READS_PER_COND, prefix='tx'):
    from nanocompore_b200.synthetic import make_transcript
    rng = np.random.default_rng(seed)
    return [make_transcript(rng, n_positions, reads, uncalled4=True, name=f'{prefix}{seed}_{i}')
            for i in range(n_tx)]


# ----------------------------------------------------------------------------------------
# CPU arms: the reference's code (baseline/_ref or /root/reference) and the oracle port
# ----------------------------------------------------------------------------------------

_REF = {}


def reference_compare_transcript(tx, methods=METHODS, min_coverage=MIN_COVERAGE):
    """One transcript through the UNMODIFIED reference: Worker._prepare_data /
    _filter_low_cov_positions / _downsample (run.py:511-601) and
    TranscriptComparator.compare_transcript (comparisons.py:41-148) on the CPU."""
    key = (tuple(methods), min_coverage)
    if key not in _REF:
        from oracle import refharness as H
        ref = H.load_reference()
        from nanocompore.run import Worker
        from nanocompore.transcript import Transcript
        conf = H.reference_config(ref, comparison_methods=list(methods), min_coverage=min_coverage)

        class W:
            _conf = conf
            _device = 'cpu'

            def log(self, level, msg):
                pass
        w = W()
        _REF[key] = (Worker, Transcript, conf, w, ref.comparisons.TranscriptComparator(config=conf, worker=w))
    Worker, Transcript, conf, w, comparator = _REF[key]
    data, s_t, c_t, pos_t = Worker._prepare_data(w, tx.data.copy(), tx.samples, tx.conditions)
    data, pos_t = Worker._filter_low_cov_positions(w, data, pos_t, c_t, conf.get_min_coverage())
    data, s_t, c_t = Worker._downsample(w, data, s_t, c_t, conf.get_downsample_high_coverage())
    _, df = comparator.compare_transcript(Transcript(0, tx.name, 'A' * tx.data.shape[0]), data, s_t, c_t, pos_t, 'cpu')
    return df


def port_compare_transcript(tx, methods=METHODS, min_coverage=MIN_COVERAGE):
    """One transcript through the numpy port: prep (run.py:511-601) + compare_transcript."""
    from oracle import prep as oprep
    from oracle.comparator import OracleComparator, OracleConfig
    data, samples, conditions, positions = oprep.prepare_data(tx.data, tx.samples, tx.conditions)
    data, positions = oprep.filter_low_cov_positions(data, positions, conditions, min_coverage)
    cmp_ = OracleComparator(OracleConfig(comparison_methods=methods, sample_ids=SAMPLE_IDS))
    return cmp_.compare_transcript(tx.name, data, samples, conditions, positions)


def _cpu_worker(args):
    seed, n_positions, kind = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass
    tx = make_transcripts(seed, 1, n_positions, prefix='cpu')[0]
    t0 = time.perf_counter()
    df = reference_compare_transcript(tx) if kind == 'reference' else port_compare_transcript(tx)
    return (0 if df is None else len(df)), time.perf_counter() - t0


def cpu_kind():
    from oracle import refharness as H
    return 'reference' if H.reference_available() else 'port'


def cpu_baseline_single_core(budget_s=20.0, n_positions=500):
    """Bounded sample on one core: slices of config-3 transcripts until ~budget_s."""
    kind = cpu_kind()
    _cpu_worker((976, 32, kind))     # imports and first-call set-up are not part of the sample
    rows, spent, n = 0, 0.0, 0
    while spent < budget_s and n < 64:
        r, dt = _cpu_worker((977 + n, n_positions, kind))
        rows += r
        spent += dt
        n += 1
    return dict(value=rows / spent, unit=UNIT, cores=1, kind=kind,
                code="reference code (comparisons.py / run.py of nanocompore 2.0.0), restated gmm_gpu"
                if kind == 'reference' else "numpy port (oracle/comparator.py)",
                sample=f"{n} slices of {n_positions} positions x {2 * READS_PER_COND} reads "
                       f"({rows} tested positions, {spent:.1f} s)")


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_positions = args.ref_positions
    kind = cpu_kind()
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        def step(i, k):
            t0 = time.perf_counter()
            out = pool.map(_cpu_worker, [(100000 + i * cores + c, n_positions, k) for c in range(cores)])
            return sum(r for r, _ in out), time.perf_counter() - t0
        pool.map(_cpu_worker, [(99000 + c, 32, kind) for c in range(cores)])   # imports, first-call set-up
        for i in range(args.warmup):
            step(i, kind)
        rows, total = 0, 0.0
        for i in range(args.steps):
            r, dt = step(args.warmup + i, kind)
            rows += r
            total += dt
        # second field: the numpy port on the same kind of step (one step is enough for a ratio)
        port = None
        if kind == 'reference':
            pool.map(_cpu_worker, [(98000 + c, 32, 'port') for c in range(cores)])
            pr, pt = step(10 ** 6, 'port')
            port = {"value": pr / pt, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"1 step x {cores} transcripts of {n_positions} positions"}
    value = rows / total
    sample = (f"{args.steps} steps x {cores} transcripts of {n_positions} positions x "
              f"{2 * READS_PER_COND} reads, one process per core")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.tx_per_batch, args.batches_per_step), "cpu_step": sample,
                   "code": "unmodified nanocompore 2.0.0 comparisons.py / run.py from baseline/_ref, "
                           "gmm_gpu restated by oracle/gmm.py" if kind == 'reference' else "numpy port"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "port": port,
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
        sm, smax, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        # "under load": samples drawing more than half of the largest power seen
        load = [s for s, p in zip(sm, power) if power and p >= 0.5 * max(power)]
        return {"sm_mhz": float(np.median(load if load else sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "samples_under_load": len(load),
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class Rig:
    """The engines / streams of one rank and the two ways a list of batches is run through them."""

    def __init__(self, local_rank, n_slots, **opts):
        import torch
        from nanocompore_b200.engine import Engine
        self.torch = torch
        self.dev = torch.device('cuda', local_rank)
        self.n_slots = n_slots
        self.engines = [Engine(local_rank, **opts) for _ in range(n_slots)]
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(n_slots)]

    def configure(self, **opts):
        for e in self.engines:
            e.configure(**opts)

    def launches(self):
        return sum(e.kernel_launches() for e in self.engines)

    def resident(self, dev_batches, order, res_dev):
        """Device-resident batches ``order`` (indices into dev_batches) round-robin on the slots.
        Returns (device ms between the first launch and the last kernel, tested positions)."""
        torch = self.torch
        from nanocompore_b200 import _lib as L
        tested = [torch.zeros((), dtype=torch.int64, device=self.dev) for _ in range(self.n_slots)]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream(self.dev)
        torch.cuda.synchronize(self.dev)
        ev0.record(main)
        for st in self.streams:
            st.wait_event(ev0)
        for i, b in enumerate(order):
            s = i % self.n_slots
            with torch.cuda.stream(self.streams[s]):
                self.engines[s].compare_csr(dev_batches[b], res_dev[s], stream=self.streams[s], wait=False)
                tested[s] += ((res_dev[s].status & L.NCB_POS_DROP_MASK) == 0).sum()
        for st in self.streams:
            done = torch.cuda.Event()
            done.record(st)
            main.wait_event(done)
        ev1.record(main)
        torch.cuda.synchronize(self.dev)
        return ev0.elapsed_time(ev1), int(sum(int(t.item()) for t in tested))

    def end_to_end(self, host_batches, order, res_host, consume=None):
        """Pinned host batches through the public API: H2D, kernels, D2H per batch; ``consume(slot,
        index in order)`` runs when the results of a batch are on the host.  Does not synchronise at
        the end beyond the waits of the slots.  Returns tested positions."""
        from nanocompore_b200 import _lib as L
        tested = 0
        pending = [None] * self.n_slots
        for i, b in enumerate(order):
            s = i % self.n_slots
            if pending[s] is not None:
                self.engines[s].wait(self.streams[s])
                tested += int(np.count_nonzero((res_host[s].status.numpy() & L.NCB_POS_DROP_MASK) == 0))
                if consume:
                    consume(s, pending[s])
            self.engines[s].compare_csr(host_batches[b], res_host[s], stream=self.streams[s], wait=False)
            pending[s] = i
        for s in range(self.n_slots):
            if pending[s] is not None:
                self.engines[s].wait(self.streams[s])
                tested += int(np.count_nonzero((res_host[s].status.numpy() & L.NCB_POS_DROP_MASK) == 0))
                if consume:
                    consume(s, pending[s])
        return tested


def _end_to_end_queue(self, host_batches, claim, res_host, consume):
    """``end_to_end`` over batches handed out by ``claim() -> global batch index | None`` (a queue shared by
    the ranks): batch g is the pool batch g mod len(host_batches); ``consume(slot, i, g)`` runs when the results
    of the i-th batch THIS rank took (global index g) are on the host.  Returns (tested positions, [g, ...])."""
    from nanocompore_b200 import _lib as L
    tested, taken = 0, []
    pending = [None] * self.n_slots

    def finish(s):
        nonlocal tested
        self.engines[s].wait(self.streams[s])
        tested += int(np.count_nonzero((res_host[s].status.numpy() & L.NCB_POS_DROP_MASK) == 0))
        consume(s, *pending[s])
        pending[s] = None

    while True:
        s = len(taken) % self.n_slots
        if pending[s] is not None:
            finish(s)
        g = claim()
        if g is None:
            break
        self.engines[s].compare_csr(host_batches[g % len(host_batches)], res_host[s], stream=self.streams[s], wait=False)
        pending[s] = (len(taken), g)
        taken.append(g)
    for s in range(self.n_slots):
        if pending[s] is not None:
            finish(s)
    return tested, taken


Rig.end_to_end_queue = _end_to_end_queue


def spot_check(tx, res, first_row, methods, min_coverage, n_check=12):
    """Oracle (oracle/comparator.py) on ``n_check`` positions of transcript ``tx``, whose result rows
    start at ``first_row`` of the batch results ``res`` (numpy dict).  Returns (values compared,
    values off): p-values to 1e-5 relative on log10 p (1e-8 absolute near p = 1), NaN pattern equal,
    dropped rows equal."""
    from nanocompore_b200 import _lib as L
    from oracle import prep as oprep
    from oracle.comparator import OracleComparator, OracleConfig
    data, samples, conditions, positions = oprep.prepare_data(tx.data, tx.samples, tx.conditions)
    data, positions = oprep.filter_low_cov_positions(data, positions, conditions, min_coverage)
    if positions.size == 0:
        return 0, 0
    pick = np.unique(np.linspace(0, positions.size - 1, n_check).astype(int))
    df = OracleComparator(OracleConfig(comparison_methods=methods, sample_ids=SAMPLE_IDS)).compare_transcript(
        tx.name, data[pick], samples, conditions, positions[pick])
    checked = off = 0
    kept = set() if df is None else set(int(p) for p in df['pos'])
    for p in positions[pick]:
        dropped = (res['status'][first_row + int(p)] & L.NCB_POS_DROP_MASK) != 0
        checked += 1
        off += int(dropped == (int(p) in kept))
    if df is None:
        return checked, off
    rows = first_row + df['pos'].to_numpy().astype(np.int64)
    for col, key in PV_COLUMN.items():
        if col not in df.columns:
            continue
        want = df[col].to_numpy().astype(np.float64)
        got = res['pvalues'][L.PV[key]][rows]
        checked += want.size
        nan_w, nan_g = np.isnan(want), np.isnan(got)
        bad = nan_w != nan_g
        ok = ~nan_w & ~nan_g
        tiny = ok & ((want < 1e-300) | (got < 1e-300))
        bad |= tiny & ((want < 1e-300) != (got < 1e-300))
        ok &= ~tiny
        with np.errstate(divide='ignore'):
            bad[ok] |= ~np.isclose(np.log10(got[ok]), np.log10(want[ok]), rtol=1e-5, atol=1e-8)
        off += int(bad.sum())
    return checked, off


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from nanocompore_b200 import _lib as L
    from nanocompore_b200 import sharding
    from nanocompore_b200.engine import CsrBatch, Results, pack_csr

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
        # carry the one JSON line only, so it points at stderr until the collectives are warm
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

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def over_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return over_ranks(x, dist.ReduceOp.MAX if world > 1 else None)

    def sum_over_ranks(x):
        return over_ranks(x, dist.ReduceOp.SUM if world > 1 else None)

    K, W, T, B = args.steps, args.warmup, args.tx_per_batch, args.batches_per_step
    opts = dict(tests=METHODS, n_samples=4, depleted_condition=0, min_coverage=MIN_COVERAGE,
                dwell_is_raw=True, cluster_counts='hard')
    rig = Rig(local_rank, args.slots, **opts)
    NSLOT = rig.n_slots
    eng0 = rig.engines[0]
    bh = eng0.bh_fdr
    if world > 1:
        sharding.warm_up(dev, bh)

    # ---- the job: ONE global list of transcripts, LPT-sharded ------------------------------
    # world x K x B x T transcripts of config 3 (weak scaling: the list grows with the ranks).  The
    # transcripts of a rank's shard are drawn from its pool of distinct synthetic batches.
    n_global = world * K * B * T
    work = sharding.work_estimate(np.full(n_global, N_POSITIONS), 2 * READS_PER_COND)
    shards = sharding.lpt_partition(work, world)
    lpt = {"transcripts": n_global, "imbalance": sharding.lpt_imbalance(work, shards),
           "per_rank": [int(s.size) for s in shards]}
    my_tx = shards[rank]
    assert my_tx.size == K * B * T

    # ---- synthetic pool: pinned host copies (e2e) and device copies (value) ------------------
    n_pool = args.pool
    host_pool, dev_pool, pool_tx0 = [], [], []
    t_gen = time.perf_counter()
    for b in range(n_pool):
        txs = make_transcripts(20251019 + CONFIG_INDEX + 1000 * rank + b, T)
        batch, _, _ = pack_csr([(t.data, t.samples, t.conditions) for t in txs], pin=True, wire='i16r')
        host_pool.append(batch)
        dev_pool.append(batch.to(dev))
        pool_tx0.append(txs[0])
        del txs
    t_gen = time.perf_counter() - t_gen
    P = host_pool[0].n_positions
    h2d_batch = int(np.mean([b.nbytes() for b in host_pool]))
    entries = int(np.mean([b.n_entries for b in host_pool]))
    d2h_batch = eng0.results_host_bytes(P)
    order_w = [i % n_pool for i in range(W * B)]
    order_k = [(W * B + i) % n_pool for i in range(K * B)]

    # ---- value: device-resident inputs -----------------------------------------------------
    res_dev = [Results(P, 4, dev) for _ in range(NSLOT)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    rig.resident(dev_pool, order_w, res_dev)
    barrier()
    launches0 = rig.launches()
    barrier()
    ms_local, tested_positions = rig.resident(dev_pool, order_k, res_dev)
    barrier()
    ms_value = max_over_ranks(ms_local)
    launches = rig.launches() - launches0
    total_tested = sum_over_ranks(float(tested_positions))
    value = total_tested / (ms_value / 1000.0)

    # parity spot check of the headline workload: first transcript of two pool batches
    check = {"values": 0, "off": 0}
    for b in (0, n_pool - 1):
        r = eng0.compare_csr(dev_pool[b], res_dev[0]).numpy()
        c, o = spot_check(pool_tx0[b], r, 0, METHODS, MIN_COVERAGE)
        check["values"] += c
        check["off"] += o

    # ---- roofline pass: the same batches on ONE stream, so that every kernel is timed alone --
    eng0.set_timing(True)
    for b in range(min(n_pool, 6)):
        eng0.compare_csr(dev_pool[b], res_dev[0], wait=False)
    totals = eng0.timing_totals()
    eng0.set_timing(False)
    fp64_peak = eng0.probe_fp64_peak()   # DFMA microbenchmark on this device, TFLOP/s

    # ---- the float32 E-step mode (ncb_opts.em_fp64 = 0, the reference's dtype): same batches, resident;
    # how often it decides differently from the fp64 definition
    r64 = eng0.compare_csr(dev_pool[0], res_dev[0]).numpy()
    rig.configure(em_fp64=False)
    rig.resident(dev_pool, order_w, res_dev)
    ms32_local, tested32 = rig.resident(dev_pool, order_k, res_dev)
    ms32 = max_over_ranks(ms32_local)
    eng0.set_timing(True)
    for b in range(min(n_pool, 6)):
        eng0.compare_csr(dev_pool[b], res_dev[0], wait=False)
    totals32 = eng0.timing_totals()
    eng0.set_timing(False)
    r32 = eng0.compare_csr(dev_pool[0], res_dev[0]).numpy()
    rig.configure(em_fp64=True)
    ok64 = (r64['status'] & L.NCB_POS_DROP_MASK) == 0
    p64, p32 = r64['pvalues'][L.PV['GMM_CHI2']][ok64], r32['pvalues'][L.PV['GMM_CHI2']][ok64]
    gate_flips = int((np.isnan(p64) != np.isnan(p32)).sum())
    both = ~np.isnan(p64) & ~np.isnan(p32)
    count_flips = int((r64['counts'][:, :, ok64] != r32['counts'][:, :, ok64]).any(axis=(0, 1)).sum())
    with np.errstate(divide='ignore', invalid='ignore'):
        p_off = int((~np.isclose(np.log10(p64[both]), np.log10(p32[both]), rtol=1e-5, atol=1e-8)).sum())
    em_fp32 = {"value": sum_over_ranks(float(tested32)) / (ms32 / 1000.0), "unit": UNIT,
               "ms_per_step": ms32 / K, "ms_mixture_per_batch": totals32['ms_gmm'] / max(totals32['batches'], 1),
               "vs_fp64_definition": {"positions": int(ok64.sum()), "bic_gate_flips": gate_flips,
                                      "positions_with_other_counts": count_flips,
                                      "chi2_pvalues_off_1e-5": p_off}}

    # ---- e2e: the whole job from pinned host batches to corrected results --------------------
    res_host = [Results(P, 4, None) for _ in range(NSLOT)]
    pv_rows = [L.PV['KS_INTENSITY'], L.PV['KS_DWELL'], L.PV['GMM_CHI2']]
    n_rows_job = K * B * P
    # With several ranks the batches of the job are handed out by a queue shared by the ranks of the box
    # (sharding.TaskQueue: list scheduling over the global batch list, the reference's task queue of
    # run.py:104-131): the host does not serve the GPUs of a box evenly (section 6 of DESIGN.md), and a rank
    # whose copies are faster takes more batches.  A rank holds room for 1.5x its even share.
    n_batches_global = world * K * B
    queue = None
    if world > 1 and args.schedule == 'queue':
        name = f"/ncb_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getuid()}"
        ok = 1.0

        def open_queue(create):
            nonlocal queue, ok
            try:
                queue = sharding.TaskQueue(name, n_batches_global, create=create)
            except Exception as e:                   # no shared memory to be had: the static partition
                print(f"[bench] rank {rank}: task queue unavailable ({e}); static LPT shards", file=sys.stderr)
                ok = 0.0

        # every rank passes the same collectives whatever happens to its own open
        if rank == 0:
            open_queue(True)
        barrier()
        if rank != 0:
            open_queue(False)
        if over_ranks(ok, dist.ReduceOp.MIN) < 1.0:
            if queue is not None:
                queue.close()
            queue = None
    cap_batches = K * B if queue is None else min(n_batches_global, (3 * K * B) // 2 + NSLOT)
    job_p = torch.empty((len(pv_rows), cap_batches * P), dtype=torch.float64, device=dev)
    job_q_host = torch.empty((len(pv_rows), cap_batches * P), dtype=torch.float64, pin_memory=True)
    # global row key: transcript index in the global list << 32 | position
    pos_in_tx = torch.arange(N_POSITIONS, device=dev, dtype=torch.int64)
    keys = torch.empty(cap_batches * P, dtype=torch.int64, device=dev)
    tx_of_row = torch.from_numpy(np.repeat(my_tx, N_POSITIONS)).to(dev)
    keys[:n_rows_job] = (tx_of_row << 32) | pos_in_tx.repeat(my_tx.size)          # the static shard
    del tx_of_row
    # keys of the batch whose first transcript is number 0 (queue: batch g holds transcripts g T .. g T + T - 1)
    base_key = (torch.arange(T, device=dev, dtype=torch.int64).repeat_interleave(N_POSITIONS) << 32) | pos_in_tx.repeat(T)

    def job_pass(order, n_rows, use_queue=False):
        """Compare + correction + gather.  The batches are the pool indices ``order`` (this rank's static
        shard), or, with ``use_queue``, whatever this rank takes from the queue shared by the ranks.
        Returns (wall ms, tested, stage ms, gather info, batches this rank ran)."""
        def consume(slot, i, g=None):
            # the p-value columns of the batch join the job's columns on the device (the correction
            # runs there); in-stream, so the next copy-out of the slot cannot overtake it
            with torch.cuda.stream(rig.streams[slot]):
                for j, r in enumerate(pv_rows):
                    job_p[j, i * P:(i + 1) * P].copy_(res_host[slot].pvalues[r], non_blocking=True)
                if g is not None:
                    torch.add(base_key, (g * T) << 32, out=keys[i * P:(i + 1) * P])

        def claim():
            nonlocal n_taken
            if n_taken >= cap_batches:
                return None
            g = queue.claim()
            n_taken += g is not None
            return g

        n_taken = 0
        barrier()
        if use_queue:
            if rank == 0:
                queue.reset()
            barrier()
        t0 = time.perf_counter()
        if use_queue:
            tested, taken = rig.end_to_end_queue(host_pool, claim, res_host, consume)
            n_rows = len(taken) * P
        else:
            tested = rig.end_to_end(host_pool, order, res_host, consume)
            taken = order
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        q = sharding.fdr_columns(job_p[:, :n_rows], bh)
        for j in range(len(pv_rows)):      # row by row: a strided block would travel through pageable memory
            job_q_host[j, :n_rows].copy_(q[j], non_blocking=True)
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        gathered = None
        if world > 1:
            gathered = sharding.gather_columns(q, keys[:n_rows])
            torch.cuda.synchronize(dev)
        t3 = time.perf_counter()
        info = None
        if gathered is not None:
            info = {"rows": int(gathered[0].shape[1]), "columns": int(gathered[0].shape[0]),
                    "sorted_by_key": bool((gathered[1][1:] > gathered[1][:-1]).all().item())}
        del gathered, q
        return ((t3 - t0) * 1000.0, tested,
                {"compare_ms": (t1 - t0) * 1000.0, "fdr_ms": (t2 - t1) * 1000.0,
                 "gather_ms": (t3 - t2) * 1000.0}, info, len(taken))

    def per_rank(x):
        """``x`` of every rank, as a list."""
        if world == 1:
            return [float(x)]
        mine = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        out = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, mine)
        return [float(v) for v in out.tolist()]

    def critical_path(stages):
        """The stage times of the rank that finishes its compare stage LAST, plus the spread of that stage
        over the ranks.  (A rank that is through its batches early waits inside the first collective of
        the correction for the slowest one: the maximum over ranks of every stage separately would add
        that wait to the correction and the parts would not sum to the job.)"""
        if world == 1:
            return stages
        names = sorted(stages)
        mine = torch.tensor([stages[k] for k in names], dtype=torch.float64, device=dev)
        table = torch.empty((world, len(names)), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(table.view(-1), mine)
        table = table.cpu().numpy()
        c = names.index('compare_ms')
        slow = int(np.argmax(table[:, c]))
        out = {k: float(table[slow, i]) for i, k in enumerate(names)}
        out["critical_rank"] = slow
        out["compare_ms_fastest_rank"] = float(table[:, c].min())
        return out

    import gc
    gc.collect()
    gc.disable()            # a collection inside the wall-clock region would be charged to the path
    job_pass(order_w, len(order_w) * P)                 # warm-up: W steps + correction + gather
    E2E_REPEATS = 3
    passes = []
    if queue is not None:
        job_pass(None, 0, use_queue=True)               # and once through the queue
    for _ in range(E2E_REPEATS):
        ms, tested_host, stages, ginfo, n_ran = job_pass(order_k, n_rows_job, use_queue=queue is not None)
        passes.append((max_over_ranks(ms), tested_host, critical_path(stages), ginfo,
                       per_rank(n_ran), per_rank(stages["compare_ms"])))
    gc.enable()
    passes.sort(key=lambda x: x[0])
    ms_e2e, tested_host, stages, ginfo, batches_per_rank, compare_ms_per_rank = passes[len(passes) // 2]
    schedule = {"kind": "queue shared by the ranks (sharding.TaskQueue: list scheduling over the global batch list)"
                        if queue is not None else "static LPT shards",
                "batches_per_rank": [int(b) for b in batches_per_rank],
                "compare_ms_per_rank": compare_ms_per_rank}
    if queue is not None:
        barrier()
        queue.close()
    clocks = sampler.stop()
    total_tested_host = sum_over_ranks(float(tested_host))
    e2e_value = total_tested_host / (ms_e2e / 1000.0)
    total_launches = int(sum_over_ranks(float(launches)))
    fdr_h2d_batch = len(pv_rows) * 8 * P               # the p-value columns go back up for the correction
    # where the correction spends its time: one more pass with a synchronisation after every stage
    # (diagnostic, outside the timed passes)
    fdr_stages = {}
    if world > 1:
        sharding.fdr_columns(job_p[:, :min(n_rows_job, int(batches_per_rank[rank]) * P)], bh, timings=fdr_stages)
        fdr_stages = {k: max_over_ranks(v) for k, v in sorted(fdr_stages.items())}
    del job_p, keys

    # what the host of this box gives N ranks at once: the signal arrays of the same pinned batches copied
    # to the device with no kernel in between, all ranks together -- the ceiling of the end-to-end figure
    # when the link, not the kernels, bounds it
    def h2d_ceiling():
        stage = [torch.empty_like(host_pool[0].signal, device=dev) for _ in range(2)]
        n_rows = min(b.signal.shape[0] for b in host_pool)
        copies = 4 * n_pool
        def run():
            for i in range(copies):
                stage[i % 2][:n_rows].copy_(host_pool[i % n_pool].signal[:n_rows], non_blocking=True)
        run()
        barrier()
        t0 = time.perf_counter()
        run()
        torch.cuda.synchronize(dev)
        dt = max_over_ranks(time.perf_counter() - t0)
        nbytes = copies * n_rows * host_pool[0].signal.shape[1] * host_pool[0].signal.element_size()
        per_rank = nbytes / dt / 1e9
        return {"per_rank_GBps": per_rank, "aggregate_GBps": per_rank * world, "ranks": world,
                "bytes_per_rank": nbytes}
    h2d_probe = h2d_ceiling()
    e2e_h2d_GBps = (h2d_batch + fdr_h2d_batch) * B * K / (stages["compare_ms"] / 1000.0) / 1e9

    # float32 wire (9 bytes per entry) on the same data, one pass of the compare stage: what the
    # int16 wire buys end to end
    f32_pool = [CsrBatch(b.pos_offsets, b.signal_f32().pin_memory(), b.label, b.max_entries) for b in host_pool]
    rig.end_to_end(f32_pool, order_w[:NSLOT * 2], res_host)
    barrier()
    t0 = time.perf_counter()
    tested_f32 = rig.end_to_end(f32_pool, order_k, res_host)
    torch.cuda.synchronize(dev)
    ms_f32 = max_over_ranks((time.perf_counter() - t0) * 1000.0)
    e2e_f32 = {"value": sum_over_ranks(float(tested_f32)) / (ms_f32 / 1000.0), "unit": UNIT,
               "h2d_bytes_per_step": int(np.mean([b.nbytes() for b in f32_pool])) * B,
               "stage": "compare only (no correction / gather)", "ms_per_step": ms_f32 / K}
    del f32_pool
    # the same int16 entries with a label byte each (NCB_LAYOUT_CSR_I16, 5 bytes per entry): what the label
    # runs buy end to end
    lab_pool = [CsrBatch(b.pos_offsets, b.signal, b.label, b.max_entries) for b in host_pool]
    rig.end_to_end(lab_pool, order_w[:NSLOT * 2], res_host)
    barrier()
    t0 = time.perf_counter()
    tested_lab = rig.end_to_end(lab_pool, order_k, res_host)
    torch.cuda.synchronize(dev)
    ms_lab = max_over_ranks((time.perf_counter() - t0) * 1000.0)
    e2e_lab = {"value": sum_over_ranks(float(tested_lab)) / (ms_lab / 1000.0), "unit": UNIT,
               "h2d_bytes_per_step": int(np.mean([b.nbytes() for b in lab_pool])) * B,
               "stage": "compare only (no correction / gather)", "ms_per_step": ms_lab / K}
    del lab_pool

    # ---- a FIXED job of the config-5 mix, LPT-sharded: strong scaling --------------------------
    job_skew = run_skew_job(args, rig, world, rank, dev, barrier, max_over_ranks, sum_over_ranks)

    # ---- the other BASELINE configs on one GPU ---------------------------------------------
    other = None
    if world == 1 and not args.no_other_configs:
        other = run_other_configs(args, rig, dev)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline: every kernel group of a batch, the largest by live time on top ----------------
    # Algorithmic bytes per launch: the statistics kernels read the int16 entries +
    # label (5 B/read) and hand the standardised values to the mixture fit (8 B/read), 8 + 132 B per
    # position of offsets and result columns; each of the three split mixture kernels reads the
    # standardised values once (8 B/read; the final pass also the label) plus offsets / fit size (12 B)
    # and the 21-double hand-over record of a position (written by init, read + 13 doubles rewritten by
    # EM, read by the final pass, which writes 44 B of result columns).
    nb = max(totals['batches'], 1)
    peak, peak_src = measured_peaks()
    traffic_src, ncu_traffic, ncu_pos = None, {}, 1
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            tr = json.load(f)
        ncu_traffic, ncu_pos, traffic_src = tr["per_kernel"], tr["positions"], tr.get("source")
    except Exception:
        pass
    fp64_ops, ops_src = {}, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'fp64_ops.json')) as f:
            ops = json.load(f)
        ops_src = ops.get("source")
        for name, k in ops["kernels"].items():
            fp64_ops[name.split('(')[0].replace('void ', '')] = k["fp64_thread_ops_per_position"]
    except Exception:
        pass

    def group(kernels, algorithmic_bytes, ms):
        """One roofline record: HBM terms from the live time; traffic (ncu, scaled to this batch) and
        the fp64 terms when the committed ncu capture has these kernels."""
        ach = algorithmic_bytes / (ms / 1000.0) / 1e9 if ms > 0 else None
        rec = {"bound": "hbm", "kernel": ' + '.join(f"ncb::{k}" for k in kernels), "achieved": ach, "peak": peak,
               "peak_source": peak_src, "unit": "GB/s", "frac": None if ach is None else ach / peak,
               "traffic": None, "algorithmic_bytes_per_launch": algorithmic_bytes, "ms_per_launch": ms}
        if all(k in ncu_traffic for k in kernels):
            rec["traffic"] = sum(ncu_traffic[k] for k in kernels) * P / ncu_pos
        if kernels and all(k in fp64_ops for k in kernels) and ms > 0:
            dfma = sum(fp64_ops[k]["dfma"] for k in kernels)
            other = sum(fp64_ops[k]["dadd"] + fp64_ops[k]["dmul"] for k in kernels)
            flop = (2.0 * dfma + other) * P
            rec["fp64"] = {"achieved": flop / (ms / 1000.0) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                           "frac": flop / (ms / 1000.0) / 1e12 / fp64_peak,
                           # share of the fp64 pipe's issue slots (an add or a multiply takes the slot of an FMA)
                           "pipe_frac": 2.0 * (dfma + other) * P / (ms / 1000.0) / 1e12 / fp64_peak,
                           "flop_per_position": flop / P}
        return rec

    K_STATS = ["position_kernel_warp<256, 8, 3, 5>", "position_kernel_warp<128, 8, 2, 5>"]
    K_INIT, K_EM, K_FINAL = "gmm_init_kernel<8, 256, 8, 3>", "gmm_em_kernel<8, 256, 8, 2, 0, 3>", "gmm_final_kernel<8, 8, 3, 0>"
    ms_stats, ms_gmm = totals['ms_stats'] / nb, totals['ms_gmm'] / nb
    groups = {
        "statistics": group(K_STATS, entries * (5 + 8) + P * (8 + 132), ms_stats),
        "mixture_init": group([K_INIT], entries * 8 + P * (12 + 21 * 8), totals['ms_gmm_init'] / nb),
        "mixture_em": group([K_EM], entries * 8 + P * (12 + 21 * 8 + 13 * 8), totals['ms_gmm_em'] / nb),
        "mixture_final": group([K_FINAL], entries * 9 + P * (12 + 21 * 8 + 44), totals['ms_gmm_final'] / nb),
    }
    top = max(groups, key=lambda g: groups[g]["ms_per_launch"])
    roofline = dict(groups[top])
    roofline["group"] = top
    roofline["traffic_source"] = traffic_src
    roofline["kernels"] = groups
    # the mixture fit as a whole (what one fused kernel would have to move: every standardised value and
    # label once, 56 B per position), against the time of its three kernels
    roofline["mixture_fit"] = group([K_INIT, K_EM, K_FINAL], entries * 9 + P * (12 + 44), ms_gmm)
    roofline["ms_ks_pvalue_kernels"] = totals['ms_ks'] / nb
    roofline["ms_classify"] = totals['ms_classify'] / nb
    roofline["fp64_peak"] = {"value": fp64_peak, "unit": "TFLOP/s", "source": "measured (DFMA microbenchmark, this run)",
                             "ops_source": ops_src}
    roofline["note"] = ("no kernel of the path is HBM bound (DRAM at 1-4 % of peak): the statistics kernels are bound by "
                        "integer issue (sort, scans), the mixture kernels by the fp64 pipe -- their fp64 entries are the "
                        "compute rooflines; the three mixture kernels each re-read the standardised values (traffic 3x the "
                        "algorithmic bytes of a fused fit) in exchange for an instruction working set that fits the "
                        "instruction cache; see DESIGN.md section 4")

    cpu = cpu_baseline_single_core(args.cpu_budget) if not args.no_cpu_baseline else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_value / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(T, B), "positions_per_step": P * B,
                   "tested_positions_per_step": tested_positions / K,
                   "positions_per_batch": P, "reads_per_batch": entries, "wire": "int16 entries + label runs per position (NCB_LAYOUT_CSR_I16_RUNS, 4 bytes per entry)",
                   "l2": f"inputs larger than L2: pool of {n_pool} distinct batches, consecutive launches read different batches",
                   "slots": NSLOT, "timed_region_s": ms_value / 1000.0,
                   "parallelism": f"transcript sharding x{world} (LPT over one global list), no data-path collective"},
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (h2d_batch + fdr_h2d_batch) * B,
                "d2h_bytes_per_step": d2h_batch * B + len(pv_rows) * 8 * P * B, "ms_per_step": ms_e2e / K,
                "job_ms": ms_e2e, "passes_job_ms": [p[0] for p in passes], "reported": "median pass",
                "stages_ms": stages, "includes": "H2D, kernels, D2H per batch; BH correction over all ranks' rows; "
                                                 "q-values to the host; gather of keys + q-values on rank 0"},
        "h2d": {"probe_all_ranks_copying": h2d_probe, "per_rank_GBps_in_the_compare_stage_of_the_job": e2e_h2d_GBps,
                "share_of_probe": e2e_h2d_GBps / h2d_probe["per_rank_GBps"]},
        "e2e_f32_wire": e2e_f32,
        "e2e_label_bytes_wire": e2e_lab,
        "em_fp32": dict(em_fp32, speedup_of_mixture_kernels=ms_gmm / em_fp32["ms_mixture_per_batch"]),
        "parity_spot_check": check,
        "gpu_launches": total_launches,
        "lpt": lpt,
        "schedule": schedule,
        "final_gather": None if ginfo is None else dict(ginfo, ms=stages["gather_ms"],
                                                        share_of_job=stages["gather_ms"] / ms_e2e),
        "fdr": {"ms": stages["fdr_ms"], "share_of_job": stages["fdr_ms"] / ms_e2e,
                "rows": int(n_rows_job * world), "columns": len(pv_rows),
                "stages_ms_diagnostic_pass": fdr_stages or None},
        "job_skew": job_skew,
        "other_configs": other,
        "clocks": clocks,
        "numa": numa,
        "setup_s": {"generate_and_pack": t_gen},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def pack_pool(tx_lists, dev, methods_wire='i16r'):
    from nanocompore_b200.engine import pack_csr
    host, devb = [], []
    for txs in tx_lists:
        batch, _, _ = pack_csr([(t.data, t.samples, t.conditions) for t in txs], pin=True, wire=methods_wire)
        host.append(batch)
        devb.append(batch.to(dev))
    return host, devb


def measure_config(rig, dev, host, devb, order, n_samples=4, timing_batches=4):
    """value / e2e / kernel times of a list of batches on one GPU (the rig is already configured)."""
    import torch
    from nanocompore_b200.engine import Results
    P = max(b.n_positions for b in host)
    res_dev = [Results(P, n_samples, dev) for _ in range(rig.n_slots)]
    res_host = [Results(P, n_samples, None) for _ in range(rig.n_slots)]
    # every batch of a config has the same number of positions (the pools are built that way)
    assert all(b.n_positions == P for b in host)
    rig.resident(devb, order[:rig.n_slots * 2], res_dev)
    ms, tested = rig.resident(devb, order, res_dev)
    rig.end_to_end(host, order[:rig.n_slots * 2], res_host)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    tested_h = rig.end_to_end(host, order, res_host)
    torch.cuda.synchronize(dev)
    ms_e2e = (time.perf_counter() - t0) * 1000.0
    eng = rig.engines[0]
    eng.set_timing(True)
    for b in order[:timing_batches]:
        eng.compare_csr(devb[b], res_dev[0], wait=False)
    t = eng.timing_totals()
    eng.set_timing(False)
    nb = max(t['batches'], 1)
    n_pos = sum(host[b].n_positions for b in order)
    entries = sum(host[b].n_entries for b in order)
    peak, _ = measured_peaks()
    hbm_bytes = 5 * entries + 112 * n_pos          # (bytes of a read on the wire) N + 112 per position
    return {"positions": n_pos, "tested_positions": tested, "reads": entries,
            "value": tested / (ms / 1000.0), "ms": ms,
            "e2e": {"value": tested_h / (ms_e2e / 1000.0), "unit": UNIT, "ms": ms_e2e,
                    "h2d_bytes": sum(host[b].nbytes() for b in order),
                    "d2h_bytes": sum(eng.results_host_bytes(host[b].n_positions) for b in order)},
            "roofline": {"hbm": {"achieved": hbm_bytes / (ms / 1000.0) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": hbm_bytes / (ms / 1000.0) / 1e9 / peak,
                                 "algorithmic_bytes": hbm_bytes},
                         "ms_per_batch": {"statistics": t['ms_stats'] / nb, "mixture": t['ms_gmm'] / nb,
                                          "ks_pvalue": t['ms_ks'] / nb, "classify": t['ms_classify'] / nb}}}, res_dev


def run_other_configs(args, rig, dev):
    """BASELINE configs 2, 4 and 5 on one GPU, each with a parity spot check against the oracle."""
    import torch
    from nanocompore_b200.synthetic import make_skew_tail
    out = []
    eng = rig.engines[0]

    def record(name, methods, min_cov, tx_lists, repeat, extra=None):
        rig.configure(tests=methods, min_coverage=min_cov)
        host, devb = pack_pool(tx_lists, dev)
        order = [i % len(host) for i in range(len(host) * repeat)]
        rec, res_dev = measure_config(rig, dev, host, devb, order)
        checked = off = 0
        for b in (0, len(host) - 1):
            r = eng.compare_csr(devb[b], res_dev[0]).numpy()
            c, o = spot_check(tx_lists[b][0], r, 0, methods, min_cov, n_check=8)
            checked += c
            off += o
        rec.update({"config": name, "methods": methods, "parity_spot_check": {"values": checked, "off": off}})
        if extra:
            rec.update(extra(rec, devb, res_dev))
        out.append(rec)
        del host, devb, res_dev
        torch.cuda.empty_cache()

    # config 2: 100 transcripts x 1,000 positions x 50 reads/condition, KS only -- the whole config per
    # pass, 4 distinct copies of it (a copy is 50 MB: L2-resident if repeated)
    record("config2: 100 transcripts x 1000 positions x 50 reads/condition, [KS] (whole config per pass, 20 passes over 4 distinct copies)",
           ['KS'], MIN_COVERAGE,
           [make_transcripts(7000 + i, 100, 1000, 50, prefix='c2_') for i in range(4)], 5)

    # config 4: transcripts of 1,000 positions x 150 reads/condition, [GMM, MW] + BH over the shard's
    # p-value columns: a 2.4 M-position shard (6 distinct batches of 100 transcripts, 4 passes)
    def fdr_of_shard(rec, devb, res_dev):
        from nanocompore_b200 import _lib as L
        rows = [L.PV['GMM_CHI2'], L.PV['MW_INTENSITY'], L.PV['MW_DWELL']]
        n = rec["positions"]
        cols = torch.rand((3, n), dtype=torch.float64, device=dev)
        for j, r in enumerate(rows):          # real p-values of one batch, tiled over the shard
            src = res_dev[0].pvalues[r]
            cols[j] = src.repeat((n + src.numel() - 1) // src.numel())[:n]
        eng.bh_fdr(cols[0])
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for j in range(3):
            eng.bh_fdr(cols[j])
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        return {"bh_fdr": {"ms": ms, "values": 3 * n, "share_of_shard": ms / (ms + rec["ms"]),
                           "value_with_fdr": rec["tested_positions"] / ((ms + rec["ms"]) / 1000.0)}}
    record("config4 shard: transcripts of 1000 positions x 150 reads/condition, [GMM, MW] + BH FDR (2.4 M positions: 6 distinct batches x 100 transcripts, 4 passes)",
           ['GMM', 'MW'], MIN_COVERAGE,
           [make_transcripts(7100 + i, 100, 1000, 150, prefix='c4_') for i in range(6)], 4, fdr_of_shard)

    # config 5 at the config's own proportions (200 : 5000): per pass 2 transcripts at 5,000
    # reads/condition + 50 of the low-coverage tail; the two kinds in batches of their own
    rng = np.random.default_rng(7200)
    heavy = [make_transcripts(7201 + i, 1, 1000, 5000, prefix='c5h_') for i in range(2)]
    tail = [make_skew_tail(rng, 25, uncalled4=True) for _ in range(2)]
    rig.configure(tests=['GMM', 'KS'], min_coverage=MIN_COVERAGE)
    rec = {"config": "config5 shard: 2 transcripts x 1000 positions at 5000 reads/condition + 50 tail transcripts "
                     "(30-200 reads/condition), [GMM, KS], 4 passes; see job_skew for the LPT-sharded job",
           "methods": ['GMM', 'KS']}
    hh, hd = pack_pool(heavy, dev)
    th, td = pack_pool(tail, dev)
    rh, res_h = measure_config(rig, dev, hh, hd, [0, 1] * 4)
    checked = off = 0
    r = eng.compare_csr(hd[0], res_h[0]).numpy()
    c, o = spot_check(heavy[0][0], r, 0, ['GMM', 'KS'], MIN_COVERAGE, n_check=2)
    checked, off = checked + c, off + o
    rt, res_t = measure_config(rig, dev, th, td, [0, 1] * 4)
    r = eng.compare_csr(td[0], res_t[0]).numpy()
    c, o = spot_check(tail[0][0], r, 0, ['GMM', 'KS'], MIN_COVERAGE, n_check=8)
    checked, off = checked + c, off + o
    tested = rh["tested_positions"] + rt["tested_positions"]
    rec.update({"positions": rh["positions"] + rt["positions"], "tested_positions": tested,
                "reads": rh["reads"] + rt["reads"],
                "value": tested / ((rh["ms"] + rt["ms"]) / 1000.0), "ms": rh["ms"] + rt["ms"],
                "e2e": {"value": tested / ((rh["e2e"]["ms"] + rt["e2e"]["ms"]) / 1000.0), "unit": UNIT,
                        "ms": rh["e2e"]["ms"] + rt["e2e"]["ms"],
                        "h2d_bytes": rh["e2e"]["h2d_bytes"] + rt["e2e"]["h2d_bytes"],
                        "d2h_bytes": rh["e2e"]["d2h_bytes"] + rt["e2e"]["d2h_bytes"]},
                "roofline": {"high_coverage": rh["roofline"], "tail": rt["roofline"]},
                "high_coverage_positions_per_s": rh["value"], "tail_positions_per_s": rt["value"],
                "parity_spot_check": {"values": checked, "off": off}})
    out.append(rec)
    rig.configure(tests=METHODS, min_coverage=MIN_COVERAGE)
    return out


def run_skew_job(args, rig, world, rank, dev, barrier, max_over_ranks, sum_over_ranks):
    """A fixed global job of the config-5 mix -- ``--skew-heavy`` transcripts at 5,000 reads/condition
    and 25 x as many of the low-coverage tail (the config's 200 : 5000) -- LPT-sharded over the ranks
    by sharding.work_estimate, run end to end (pinned host batches in, results on the host, BH
    correction over all ranks' rows).  The same job at every N: strong scaling."""
    import torch
    from nanocompore_b200 import _lib as L
    from nanocompore_b200 import sharding
    from nanocompore_b200.engine import Results, pack_csr
    from nanocompore_b200.synthetic import make_transcript
    if args.skew_heavy <= 0:
        return None
    n_heavy, n_tail = args.skew_heavy, 25 * args.skew_heavy
    rng = np.random.default_rng(5005)
    # the global list: (positions, reads per condition); tail coverage ~ truncated geometric in [30, 200]
    reads = np.concatenate([np.full(n_heavy, 5000), np.minimum(200, 30 + rng.geometric(1.0 / 50, n_tail))])
    order = rng.permutation(reads.size)       # shuffled order, as the config says
    reads = reads[order]
    work = sharding.work_estimate(np.full(reads.size, 1000), 2 * reads)
    shards = sharding.lpt_partition(work, world)
    mine = shards[rank]
    # this rank's transcripts: heavy ones from a pool of 2 distinct synthetic transcripts, tail ones
    # generated one by one; heavy transcripts are batches of their own, tail ones go 25 to a batch
    t_gen = time.perf_counter()
    heavy_pool = [make_transcript(np.random.default_rng(5100 + i), 1000, 5000, uncalled4=True) for i in range(2)]
    batches, n_h = [], 0
    tail_items = []
    for t in mine:
        if reads[t] == 5000:
            tx = heavy_pool[n_h % 2]
            n_h += 1
            batches.append(pack_csr([(tx.data, tx.samples, tx.conditions)], pin=True, wire='i16r')[0])
        else:
            tx = make_transcript(np.random.default_rng(5200 + int(t)), 1000, int(reads[t]), uncalled4=True)
            tail_items.append((tx.data, tx.samples, tx.conditions))
            if len(tail_items) == 25:
                batches.append(pack_csr(tail_items, pin=True, wire='i16r')[0])
                tail_items = []
    if tail_items:
        batches.append(pack_csr(tail_items, pin=True, wire='i16r')[0])
    t_gen = time.perf_counter() - t_gen
    # heavy batches first: the long ones do not form the tail of the job
    batches.sort(key=lambda b: -b.n_entries)
    n_rows = sum(b.n_positions for b in batches)
    pv_rows = [L.PV['KS_INTENSITY'], L.PV['KS_DWELL'], L.PV['GMM_CHI2']]
    job_p = torch.full((3, max(n_rows, 1)), float('nan'), dtype=torch.float64, device=dev)
    res = [Results(max((b.n_positions for b in batches), default=1), 4, None) for _ in range(rig.n_slots)]
    row0 = np.concatenate([[0], np.cumsum([b.n_positions for b in batches])]).astype(np.int64)
    rig.configure(tests=['GMM', 'KS'], min_coverage=MIN_COVERAGE)

    def one_pass():
        tested = 0
        pending = [None] * rig.n_slots

        def finish(s):
            nonlocal tested
            i, view = pending[s]
            rig.engines[s].wait(rig.streams[s])
            tested += int(np.count_nonzero((view.status.numpy() & L.NCB_POS_DROP_MASK) == 0))
            with torch.cuda.stream(rig.streams[s]):
                for j, r in enumerate(pv_rows):
                    job_p[j, row0[i]:row0[i + 1]].copy_(view.pvalues[r], non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for i, b in enumerate(batches):
            s = i % rig.n_slots
            if pending[s] is not None:
                finish(s)
            view = res[s].carve(b.n_positions)     # the slot's pinned buffers serve batches of any size
            rig.engines[s].compare_csr(b, view, stream=rig.streams[s], wait=False)
            pending[s] = (i, view)
        for s in range(rig.n_slots):
            if pending[s] is not None:
                finish(s)
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        q = sharding.fdr_columns(job_p[:, :n_rows], rig.engines[0].bh_fdr)
        q_host = q.cpu()
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        del q, q_host
        return (t2 - t0) * 1000.0, (t1 - t0) * 1000.0, tested

    one_pass()
    runs = [one_pass() for _ in range(3)]
    runs = [(max_over_ranks(a), max_over_ranks(b), c) for a, b, c in runs]
    runs.sort(key=lambda x: x[0])
    ms, ms_compare, tested = runs[1]
    total_tested = sum_over_ranks(float(tested))
    rig.configure(tests=METHODS, min_coverage=MIN_COVERAGE)
    del job_p, batches
    torch.cuda.empty_cache()
    return {"workload": f"config-5 mix, fixed at every N: {n_heavy} transcripts x 1000 positions at 5000 reads/condition "
                        f"+ {n_tail} tail transcripts (30-200 reads/condition), shuffled, [GMM, KS] + BH FDR, end to end",
            "scaling": "strong", "transcripts": int(reads.size), "positions": int(reads.size) * 1000,
            "tested_positions": total_tested, "job_ms": ms, "compare_ms": ms_compare,
            "value": total_tested / (ms / 1000.0), "unit": UNIT,
            "lpt": {"imbalance": sharding.lpt_imbalance(work, shards),
                    "per_rank_transcripts": [int(s.size) for s in shards],
                    "per_rank_heavy": [int((reads[s] == 5000).sum()) for s in shards]},
            "passes_job_ms": [r[0] for r in runs], "generate_s": t_gen}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--tx-per-batch', type=int, default=50, help="transcripts per batch (one ncb_compare call)")
    ap.add_argument('--batches-per-step', type=int, default=32)
    ap.add_argument('--pool', type=int, default=12, help="distinct synthetic batches per rank")
    ap.add_argument('--cpu-budget', type=float, default=20.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true')
    ap.add_argument('--skew-heavy', type=int, default=16,
                    help="high-coverage transcripts of the fixed config-5 job (0 = skip it)")
    ap.add_argument('--schedule', default='queue', choices=['queue', 'static'],
                    help="N > 1: batches of the end-to-end job from a queue shared by the ranks, or static LPT shards")
    ap.add_argument('--slots', type=int, default=4,
                    help="engines/streams of the end-to-end pipeline (copy-in, kernels, copy-out overlap)")
    ap.add_argument('--ref-positions', type=int, default=N_POSITIONS,
                    help="positions per transcript of the reference arm (one transcript per core per step)")
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
