"""
Transcript sharding across the GPUs of one box.

The reference parallelises over transcripts with one OS process per device entry of the
``devices:`` config key, pulling from a queue (run.py:104-131), and corrects the p-values of the
whole result set afterwards (postprocessing.py:209-285).  Here: one process (rank) per GPU under
``torch.distributed``; transcripts are assigned up front by a deterministic
longest-processing-time (LPT) bin packing on a work estimate, every rank runs its shard with no
data-path collective (positions are independent), and the job ends with two exchanges over
NCCL / NVLink (gloo in the CPU tests):

  * ``fdr_columns``     Benjamini-Hochberg over the p-value columns of ALL ranks as a sample sort:
                        the value axis of a column is cut into one range per rank, an all-to-all
                        brings every range to its rank, which runs the GPU sort-scan
                        (``ncb_bh_fdr_segment``) with the ranks of the whole column; a second
                        all-to-all brings every rank the q-values of its own rows.  Every rank does
                        1 / world of every column (with whole columns on owner ranks, three ranks of
                        eight sorted 5 x 10^8 values each while five idled: 0.9 s of a 3.1 s job).
  * ``gather_columns``  the result columns of every rank on one rank, rows sorted by a global key
                        (so the table does not depend on the sharding).

``TaskQueue`` is the dynamic alternative to a fixed partition: the task list (heaviest first) is known to
every rank and a counter in shared memory hands out its next entry, so a rank the host serves faster takes more
tasks -- what the reference's task queue does for its worker processes (run.py:104-131).

Both exchanges use pre-sized buffers and collectives only (``all_gather_into_tensor`` /
``all_to_all_single``): one communicator-wide operation each, no per-peer send/recv lists.
``warm_up`` runs them once on a few elements so that NCCL sets its channels up before a timed job.
"""
import numpy as np
import torch
import torch.distributed as dist


def work_estimate(n_positions, reads_per_position, gmm=True, ks=True):
    """Relative cost of a transcript: positions x reads x (sort/statistics + EM + exact KS).

    The exact KS p-value walks a lattice band of ~n0 x 2 D n1 cells, and D shrinks like
    1 / sqrt(n) under the null: its cost per read grows like sqrt(reads) and takes over at high
    coverage.  The weights are fitted to the measured device time per position at 200 reads
    (33 ns, BASELINE config 3) and at 10 000 reads (8.6 us, the 5 000-per-condition positions of
    config 5): a ratio of 260, where reads x log(reads) alone predicts 55."""
    n = np.asarray(reads_per_position, dtype=np.float64)
    per_read = 1.0 + (8.0 if gmm else 0.0) + 0.2 * np.log2(np.maximum(n, 2.0))
    if ks:
        per_read = per_read + 1.6 * np.sqrt(n)
    return np.asarray(n_positions, dtype=np.float64) * n * per_read


def lpt_partition(work, world_size):
    """Deterministic LPT: heaviest transcript first onto the least loaded rank; ties broken by
    transcript index, then by rank.  Returns a list of index arrays, one per rank (each in
    ascending transcript order)."""
    work = np.asarray(work, dtype=np.float64)
    order = np.lexsort((np.arange(work.size), -work))
    load = np.zeros(world_size)
    shards = [[] for _ in range(world_size)]
    for t in order:
        r = int(np.argmin(load))       # first minimum: lowest rank wins ties
        shards[r].append(int(t))
        load[r] += work[t]
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


class TaskQueue:
    """Cursor over a task list shared by the ranks of one box (``ncb_counter_*``: a 64-bit word in POSIX shared
    memory, atomic fetch-add).  Every rank holds the same list (e.g. ``lpt_order(work)``: heaviest first, so that
    handing the next task to whichever rank is free is list scheduling); ``claim()`` returns the index of the
    next task nobody has taken, or None when the list is exhausted.

    ``name``: the same on every rank of the job and different between jobs that run at the same time (bench.py
    derives it from MASTER_PORT); rank 0 passes ``create=True`` BEFORE the others open it (a barrier in between)
    and calls ``reset`` between passes over the list (again with barriers on both sides)."""

    def __init__(self, name, n_tasks, create=False):
        from nanocompore_b200 import _lib as L
        import ctypes as C
        self._lib = L.load()
        self._h = C.c_void_p()
        self.name = name if name.startswith('/') else '/' + name
        self.n_tasks = int(n_tasks)
        self._owner = bool(create)
        rc = self._lib.ncb_counter_open(self.name.encode(), int(create), C.byref(self._h))
        if rc != L.NCB_OK:
            raise L.NcbError(rc, f"ncb_counter_open({self.name!r})")

    def claim(self, n=1):
        """First index of the next ``n`` tasks (the caller gets [i, min(i + n, n_tasks))), or None."""
        i = int(self._lib.ncb_counter_add(self._h, int(n)))
        return i if i < self.n_tasks else None

    def reset(self, value=0):
        self._lib.ncb_counter_set(self._h, int(value))

    def close(self):
        if self._h:
            self._lib.ncb_counter_close(self._h, int(self._owner))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def lpt_order(work):
    """Task indices, heaviest first (ties by index): the order a ``TaskQueue`` hands them out in."""
    work = np.asarray(work, dtype=np.float64)
    return np.lexsort((np.arange(work.size), -work))


def lpt_imbalance(work, shards):
    """max load / mean load of a partition (1.0 = perfectly even)."""
    work = np.asarray(work, dtype=np.float64)
    loads = np.array([work[s].sum() for s in shards])
    return float(loads.max() / loads.mean()) if loads.size and loads.mean() > 0 else 1.0


def _world(group):
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def _all_counts(n_local, device, group, world):
    mine = torch.tensor([n_local], dtype=torch.int64, device=device)
    counts = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    return [int(c) for c in counts.tolist()]


FDR_SAMPLE = 4096      # numbers every rank contributes to the splitter sample of a column


def sample_rows(n, k, device=None):
    """``k`` evenly spaced row indices of 0 .. n - 1 (first and last included), in integer arithmetic:
    a float32 ``linspace`` cannot represent the last row of a shard of more than 2^24 rows and rounds it
    past the end (64 M rows per rank in bench.py's job)."""
    i = torch.arange(k, dtype=torch.int64, device=device)
    return (i * (n - 1)) // max(k - 1, 1)


class _Stages:
    """Wall-clock milliseconds per stage of ``fdr_columns`` when a dict is passed as ``timings`` (every
    stage then ends with a device synchronisation: a diagnostic pass, not the timed one)."""

    def __init__(self, timings, device):
        import time
        self.t, self.dev, self.clock = timings, device, time.perf_counter
        self.last = self._now() if timings is not None else 0.0

    def _now(self):
        if self.dev.type == 'cuda':
            torch.cuda.synchronize(self.dev)
        return self.clock()

    def mark(self, name):
        if self.t is None:
            return
        now = self._now()
        self.t[name] = self.t.get(name, 0.0) + (now - self.last) * 1000.0
        self.last = now


def fdr_columns(cols, bh, group=None, timings=None):
    """Benjamini-Hochberg q-values of p-value columns whose rows are spread over the ranks.

    ``cols``: float64 tensor [C, n_local], the columns of this rank's rows (NaN = not tested);
    ``bh``: callable ``bh(p, above=0, total=None)`` giving the q-values of a 1-D float64 tensor of
    p-values, NaNs skipped, when they are one value range of a column of ``total`` numbers with
    ``above`` larger ones elsewhere (``Engine.bh_fdr`` on a GPU, ``oracle.fdr.fdr_bh_segment`` in the
    CPU tests).  Returns [C, n_local]: for every row of this rank the q-value it has in the
    correction of the WHOLE column (all ranks' rows), as
    ``Postprocessor._multipletests_filter_nan`` (postprocessing.py:255-285) gives on the full table.
    The result does not depend on how the rows are sharded.

    Every rank does 1 / world of the work of every column (a sample sort): splitters from a pooled
    sample cut the value axis into ``world`` ranges of about equal population, an all-to-all brings
    range r to rank r, which sorts it and computes q = p m / rank with the ranks of the whole column
    (it knows how many numbers lie above its range) and the running minimum inside the range; the
    minimum over the ranges above arrives with one small all-gather; a second all-to-all takes the
    q-values back to the rows they belong to.  Equal p-values always share a range."""
    world, rank = _world(group)
    C, n_local = cols.shape
    if world == 1:
        return torch.stack([bh(cols[c].contiguous()) for c in range(C)]) if C else cols.clone()
    device = cols.device
    inf = float('inf')
    out = torch.full_like(cols, float('nan'))
    st = _Stages(timings, device)
    for c in range(C):
        col = cols[c]
        rows = torch.nonzero(col == col).squeeze(1)          # rows that hold a number
        v = col[rows]
        st.mark('select_rows')
        # splitters: FDR_SAMPLE numbers of every rank (evenly spaced rows; NaN where a rank has none)
        if v.numel():
            mine = v[sample_rows(v.numel(), FDR_SAMPLE, device)]
        else:
            mine = torch.full((FDR_SAMPLE,), float('nan'), dtype=cols.dtype, device=device)
        pooled = torch.empty(world * FDR_SAMPLE, dtype=cols.dtype, device=device)
        dist.all_gather_into_tensor(pooled, mine, group=group)
        pooled = torch.sort(pooled[pooled == pooled]).values
        if pooled.numel():
            at = (torch.arange(1, world, device=device) * pooled.numel()) // world
            splitters = pooled[at]
        else:
            splitters = torch.zeros(world - 1, dtype=cols.dtype, device=device)
        st.mark('splitters')
        # range of a value: the number of splitters below it (equal values: one range)
        dest = torch.bucketize(v, splitters).to(torch.uint8 if world <= 256 else torch.int32)
        order = torch.argsort(dest, stable=True)
        send = v[order]
        rows = rows[order]
        counts = torch.bincount(dest.long(), minlength=world)
        st.mark('partition')
        table = torch.empty((world, world), dtype=torch.int64, device=device)    # [source, range]
        dist.all_gather_into_tensor(table.view(-1), counts, group=group)
        table = table.cpu()
        send_split = table[rank].tolist()
        recv_split = table[:, rank].tolist()
        per_range = table.sum(0)
        total = int(per_range.sum())
        if total == 0:
            continue                                          # the same decision on every rank
        above = int(per_range[rank + 1:].sum())
        st.mark('count_table')
        recv = torch.empty(sum(recv_split), dtype=cols.dtype, device=device)
        dist.all_to_all_single(recv, send, recv_split, send_split, group=group)
        st.mark('all_to_all_values')
        q = bh(recv, above=above, total=total) if recv.numel() else recv
        st.mark('sort_scan')
        low = q.min().reshape(1) if q.numel() else torch.full((1,), inf, dtype=cols.dtype, device=device)
        lows = torch.empty(world, dtype=cols.dtype, device=device)
        dist.all_gather_into_tensor(lows, low, group=group)
        if rank + 1 < world and q.numel():
            q = torch.clamp(q, max=lows[rank + 1:].min())     # running minimum over the ranges above
        back = torch.empty(sum(send_split), dtype=cols.dtype, device=device)
        st.mark('range_minima')
        dist.all_to_all_single(back, q.contiguous(), send_split, recv_split, group=group)
        st.mark('all_to_all_q')
        out[c, rows] = back
        st.mark('scatter_rows')
    return out


def gather_columns(local, keys_local, world_size=None, rank=None, dst=0, group=None):
    """Gathers per-position result columns of every rank on ``dst``.

    ``local``: float64 tensor [C, n_local] (columns x positions of this rank's shard);
    ``keys_local``: int64 tensor [n_local] global row keys (e.g. transcript index << 32 | pos).
    Ranks hold different numbers of positions: the counts are exchanged, every rank pads its block to
    the largest and the blocks travel in ONE ``all_gather_into_tensor`` (a ring / NVLS collective on
    the communicator's existing channels; a list ``gather`` is per-peer send/recv, whose channels NCCL
    sets up lazily: 170 ms per peer in round 1).  Rank ``dst`` returns (columns [C, n_total], keys
    [n_total]) sorted by key -- on the device the columns live on -- so the row order does not depend
    on the sharding.  Other ranks get None."""
    world_size = dist.get_world_size(group) if world_size is None else world_size
    rank = dist.get_rank(group) if rank is None else rank
    device = local.device
    C, n_local = local.shape
    counts = _all_counts(n_local, device, group, world_size)
    n_max = max(counts) if counts else 0
    block = torch.full((C + 1, n_max), float('nan'), dtype=torch.float64, device=device)
    block[:C, :n_local] = local
    # keys ride along as float64 bit patterns (a view, not a conversion)
    if n_local:
        block[C, :n_local] = keys_local.contiguous().view(torch.float64)
    # the output is the concatenation of the blocks along the first dimension
    gathered = torch.empty((world_size * (C + 1), n_max), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(gathered, block, group=group)
    gathered = gathered.view(world_size, C + 1, n_max)
    if rank != dst:
        return None
    cols = torch.cat([gathered[r, :C, :c] for r, c in enumerate(counts)], dim=1)
    keys = torch.cat([gathered[r, C, :c].contiguous().view(torch.int64) for r, c in enumerate(counts)])
    order = torch.argsort(keys, stable=True)
    return cols[:, order], keys[order]


def warm_up(device, bh=None, group=None):
    """Runs every collective of this module once on a handful of elements: NCCL builds the channels
    of a collective kind the first time it is used (hundreds of milliseconds with 8 ranks), which
    belongs to the start-up of a job, not to its end."""
    world, rank = _world(group)
    if world == 1:
        return
    cols = torch.rand((3, 4 + rank), dtype=torch.float64, device=device)
    keys = torch.arange(4 + rank, dtype=torch.int64, device=device) | (rank << 32)
    fdr_columns(cols, bh or (lambda p, above=0, total=None: p.clone()), group=group)
    gather_columns(cols, keys, group=group)
    if device.type == 'cuda':
        torch.cuda.synchronize(device)
