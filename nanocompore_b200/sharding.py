"""
Transcript sharding across the GPUs of one box.

The reference parallelises over transcripts with one OS process per device entry of the
``devices:`` config key, pulling from a queue (run.py:104-131).  Here: one process (rank) per
GPU under ``torch.distributed``; transcripts are assigned up front by a deterministic
longest-processing-time (LPT) bin packing on a work estimate, every rank runs its shard with
no data-path collective (positions are independent), and the fixed-width result columns are
gathered once at the end (NCCL on GPUs; gloo in the CPU tests).
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


def gather_columns(local, keys_local, world_size=None, rank=None, dst=0, group=None):
    """Gathers per-position result columns of every rank on ``dst``.

    ``local``: float64 tensor [C, n_local] (columns x positions of this rank's shard);
    ``keys_local``: int64 tensor [n_local] global row keys (e.g. transcript index << 32 | pos).
    Ranks hold different numbers of positions: counts are exchanged first, columns are padded
    to the maximum, gathered, and rank ``dst`` returns (columns [C, n_total], keys [n_total])
    sorted by key, so the row order does not depend on the sharding.  Other ranks get None.
    """
    world_size = dist.get_world_size(group) if world_size is None else world_size
    rank = dist.get_rank(group) if rank is None else rank
    device = local.device
    n_local = torch.tensor([local.shape[1]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world_size)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts) if counts else 0
    C = local.shape[0]
    padded = torch.full((C + 1, n_max), float('nan'), dtype=torch.float64, device=device)
    padded[:C, :local.shape[1]] = local
    # keys ride along as exactly representable float64 bit patterns
    padded[C, :local.shape[1]] = keys_local.view(torch.float64) if keys_local.numel() else keys_local.to(torch.float64)
    gathered = [torch.empty_like(padded) for _ in range(world_size)] if rank == dst else None
    dist.gather(padded, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    cols = torch.cat([g[:C, :c] for g, c in zip(gathered, counts)], dim=1)
    keys = torch.cat([g[C, :c].contiguous().view(torch.int64) for g, c in zip(gathered, counts)])
    order = torch.argsort(keys, stable=True)
    return cols[:, order], keys[order]
