"""
Transcript sharding across the GPUs of one box.

The reference parallelises over transcripts with one OS process per device entry of the
``devices:`` config key, pulling from a queue (run.py:104-131), and corrects the p-values of the
whole result set afterwards (postprocessing.py:209-285).  Here: one process (rank) per GPU under
``torch.distributed``; transcripts are assigned up front by a deterministic
longest-processing-time (LPT) bin packing on a work estimate, every rank runs its shard with no
data-path collective (positions are independent), and the job ends with two exchanges over
NCCL / NVLink (gloo in the CPU tests):

  * ``fdr_columns``     Benjamini-Hochberg over the p-value columns of ALL ranks: column c goes to
                        its owner rank (c mod world) in one all-to-all, the owner runs the GPU
                        sort-scan (``ncb_bh_fdr``) over the whole column, a second all-to-all brings
                        every rank the q-values of its own rows.  No rank holds more than its own
                        rows plus the columns it owns.
  * ``gather_columns``  the result columns of every rank on one rank, rows sorted by a global key
                        (so the table does not depend on the sharding).

Both use pre-sized buffers and collectives only (``all_gather_into_tensor`` /
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


def fdr_columns(cols, bh, group=None):
    """Benjamini-Hochberg q-values of p-value columns whose rows are spread over the ranks.

    ``cols``: float64 tensor [C, n_local], the columns of this rank's rows (NaN = not tested);
    ``bh``: callable mapping a 1-D float64 tensor of p-values to its q-values, NaNs skipped
    (``Engine.bh_fdr`` on a GPU).  Returns [C, n_local]: for every row of this rank the q-value it
    has in the correction of the WHOLE column (all ranks' rows), as
    ``Postprocessor._multipletests_filter_nan`` (postprocessing.py:255-285) gives on the full table.
    The result does not depend on how the rows are sharded."""
    world, rank = _world(group)
    C, n_local = cols.shape
    if world == 1:
        return torch.stack([bh(cols[c].contiguous()) for c in range(C)]) if C else cols.clone()
    device = cols.device
    counts = _all_counts(n_local, device, group, world)
    owned = [c for c in range(C) if c % world == rank]            # my columns
    n_owned = [len(range(r, C, world)) for r in range(world)]     # columns owned by rank r
    # send: for destination r, the columns it owns (each [n_local]), column after column
    send = torch.cat([cols[c] for r in range(world) for c in range(r, C, world)]) if C else cols.reshape(0)
    send_split = [n_owned[r] * n_local for r in range(world)]
    recv_split = [len(owned) * counts[r] for r in range(world)]
    recv = torch.empty(sum(recv_split), dtype=cols.dtype, device=device)
    dist.all_to_all_single(recv, send.contiguous(), recv_split, send_split, group=group)
    # recv: source rank after source rank, [len(owned), counts[r]] each -> whole columns
    qback = torch.empty_like(recv)
    if owned:
        parts = torch.split(recv, recv_split)
        qparts = [torch.empty_like(p).view(len(owned), -1) for p in parts]
        for j in range(len(owned)):
            whole = torch.cat([p.view(len(owned), -1)[j] for p in parts])
            q = bh(whole)
            for qp, piece in zip(qparts, torch.split(q, counts)):
                qp[j] = piece
        qback = torch.cat([qp.reshape(-1) for qp in qparts])
    # reverse exchange: every rank gets the q-values of its own rows back, owner after owner
    got = torch.empty(sum(send_split), dtype=cols.dtype, device=device)
    dist.all_to_all_single(got, qback, send_split, recv_split, group=group)
    out = torch.empty_like(cols)
    for r, chunk in enumerate(torch.split(got, send_split)):
        for j, c in enumerate(range(r, C, world)):
            out[c] = chunk.view(n_owned[r], n_local)[j]
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
    fdr_columns(cols, bh or (lambda p: p.clone()), group=group)
    gather_columns(cols, keys, group=group)
    if device.type == 'cuda':
        torch.cuda.synchronize(device)
