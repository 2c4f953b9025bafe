"""Transcript sharding (LPT) and the end-of-run gather, on CPU with gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nanocompore_b200 import sharding
from nanocompore_b200.sharding import (fdr_columns, gather_columns, lpt_imbalance, lpt_partition, warm_up,
                                       work_estimate)


def test_lpt_partition_is_deterministic_and_balanced():
    rng = np.random.default_rng(0)
    P = rng.integers(500, 4000, size=400)
    n = np.where(rng.random(400) < 0.05, 10000, rng.integers(60, 400, size=400))
    w = work_estimate(P, n)
    a = lpt_partition(w, 8)
    b = lpt_partition(w, 8)
    assert all((x == y).all() for x, y in zip(a, b))
    assert sorted(np.concatenate(a).tolist()) == list(range(400))
    loads = np.array([w[s].sum() for s in a])
    # balanced up to the granularity of a transcript: with the exact-KS term a 10 000-read
    # transcript can weigh more than a rank's fair share, and then it is a rank's whole load
    assert loads.max() <= max(1.05 * loads.mean(), 1.0001 * w.max())
    light = n < 10000
    a = lpt_partition(w[light], 8)
    loads = np.array([w[light][s].sum() for s in a])
    assert loads.max() / loads.mean() < 1.05
    assert all((np.diff(s) > 0).all() for s in a)


def test_work_estimate_follows_the_measured_cost_ratio():
    # device time per position: 33 ns at 200 reads, 8.6 us at 10 000 reads (profiles/r1_highcov_ks.jsonl)
    ratio = work_estimate(1, 10000) / work_estimate(1, 200)
    assert 200 < ratio < 320
    assert work_estimate(1, 10000, ks=False) / work_estimate(1, 200, ks=False) < 60
    assert work_estimate(10, 200) == 10 * work_estimate(1, 200)


def test_lpt_partition_edge_cases():
    assert [s.tolist() for s in lpt_partition([], 2)] == [[], []]
    assert [s.tolist() for s in lpt_partition([5.0], 3)] == [[0], [], []]
    assert [s.tolist() for s in lpt_partition([1, 1, 1, 1], 2)] == [[0, 2], [1, 3]]


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    work = np.array([10, 30, 20, 5, 25, 15, 40], dtype=float)
    shard = lpt_partition(work, world)[rank]
    # rows of the shard: 3 positions per transcript, 2 result columns
    keys = torch.tensor([(int(t) << 32) | p for t in shard for p in range(3)], dtype=torch.int64)
    cols = torch.stack([keys.to(torch.float64) * 0.5, torch.full((keys.numel(),), float(rank), dtype=torch.float64)])
    if rank == 1:
        cols[0, 0] = float('nan')
    res = gather_columns(cols, keys)
    if rank == 0:
        c, k = res
        torch.save({'cols': c, 'keys': k}, out)
    else:
        assert res is None
    dist.destroy_process_group()


def test_gather_columns_gloo_world2(tmp_path):
    out = str(tmp_path / 'gathered.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    keys = got['keys'].numpy()
    assert (np.diff(keys) > 0).all() and keys.size == 21
    expect = np.array([(t << 32) | p for t in range(7) for p in range(3)])
    assert (keys == expect).all()
    c0 = got['cols'][0].numpy()
    nan = np.isnan(c0)
    assert nan.sum() == 1
    assert (c0[~nan] == keys[~nan] * 0.5).all()


def _columns_of(rank, n_cols):
    rng = np.random.default_rng(100 + rank)
    n = 37 + 11 * rank                      # ranks hold different numbers of rows
    cols = rng.random((n_cols, n)) ** 3
    cols[rng.random((n_cols, n)) < 0.1] = np.nan
    if n_cols:
        cols[0, :5] = cols[0, 5:10]         # ties
    return cols


def _fdr_worker(rank, world, port, out, n_cols):
    from oracle.fdr import fdr_bh_segment
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    # (the GPU path passes Engine.bh_fdr)
    bh = lambda p, above=0, total=None: torch.from_numpy(fdr_bh_segment(p.numpy(), above, total))      # noqa: E731
    warm_up(torch.device('cpu'), bh)
    q = fdr_columns(torch.from_numpy(_columns_of(rank, n_cols)), bh)
    torch.save(q, f'{out}.{rank}')
    dist.destroy_process_group()


@pytest.mark.parametrize('world,n_cols', [(2, 1), (2, 3), (2, 5), (3, 2)])
def test_fdr_columns_gloo_equals_the_correction_of_the_whole_table(tmp_path, world, n_cols):
    """postprocessing.py:255-285 corrects a column over ALL result rows: sharded over the ranks the
    q-values must be those of the concatenated column (oracle/fdr.py), NaNs skipped, however the
    value ranges fall."""
    from oracle.fdr import multipletests_filter_nan
    out = str(tmp_path / 'q.pt')
    mp.spawn(_fdr_worker, args=(world, _free_port(), out, n_cols), nprocs=world, join=True)
    whole = np.concatenate([_columns_of(r, n_cols) for r in range(world)], axis=1)
    want = np.stack([multipletests_filter_nan(whole[c]) for c in range(n_cols)])
    got = np.concatenate([torch.load(f'{out}.{r}').numpy() for r in range(world)], axis=1)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and not np.isnan(want).all()
    assert np.allclose(got, want, rtol=1e-13, atol=0, equal_nan=True)


def test_segment_correction_of_value_ranges_composes_to_the_whole_column():
    """oracle.fdr.fdr_bh_segment (the contract of ncb_bh_fdr_segment): ranges corrected one by one with
    the ranks of the whole column, then the minimum of the ranges above, give the whole-column q-values."""
    from oracle.fdr import fdr_bh_segment, multipletests_filter_nan
    rng = np.random.default_rng(5)
    p = rng.random(500) ** 2
    p[rng.random(500) < 0.1] = np.nan
    p[:20] = p[20:40]
    want = multipletests_filter_nan(p)
    cuts = [0.0, 0.05, 0.3, 0.31, 2.0]
    total = int((~np.isnan(p)).sum())
    got = np.full(p.size, np.nan)
    lows = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        sel = (p > lo) & (p <= hi) if lo > 0 else (p <= hi)
        above = int((p > hi).sum())
        q = fdr_bh_segment(p[sel], above, total)
        got[sel] = q
        lows.append((hi, np.nanmin(q) if q.size else np.inf))
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        sel = (p > lo) & (p <= hi) if lo > 0 else (p <= hi)
        carry = min([m for h, m in lows if h > hi], default=np.inf)
        got[sel] = np.minimum(got[sel], carry)
    assert np.allclose(got, want, rtol=1e-13, atol=0, equal_nan=True)


def test_fdr_columns_single_process():
    from oracle.fdr import fdr_bh_segment, multipletests_filter_nan
    cols = _columns_of(0, 3)
    q = fdr_columns(torch.from_numpy(cols), lambda p, above=0, total=None: torch.from_numpy(fdr_bh_segment(p.numpy(), above, total)))
    assert np.allclose(q.numpy(), np.stack([multipletests_filter_nan(c) for c in cols]), equal_nan=True)


def test_lpt_imbalance():
    w = np.array([4.0, 3.0, 2.0, 1.0])
    shards = lpt_partition(w, 2)
    assert lpt_imbalance(w, shards) == 1.0
    assert lpt_imbalance(w, [np.array([0, 1]), np.array([2, 3])]) == pytest.approx(1.4)


def test_sample_rows_stays_inside_large_shards():
    """The splitter sample of fdr_columns on shards beyond 2^24 rows (bench.py: 64 M rows per rank): a
    float32 linspace rounds the last index past the end there."""
    for n in (1, 2, 4095, 4096, 4097, (1 << 24) + 1, 63_999_999, 64_000_000, (1 << 31) + 5):
        pick = sharding.sample_rows(n, sharding.FDR_SAMPLE)
        assert pick.dtype == torch.int64 and pick.numel() == sharding.FDR_SAMPLE
        assert int(pick[0]) == 0 and int(pick[-1]) == n - 1
        assert bool((pick[1:] >= pick[:-1]).all()) and int(pick.max()) < n


def _claim_all(name, n_tasks, out):
    import time
    from nanocompore_b200.sharding import TaskQueue
    q = TaskQueue(name, n_tasks)
    mine = []
    while True:
        i = q.claim()
        if i is None:
            break
        mine.append(i)
        if len(mine) % 64 == 0:
            time.sleep(0)
    q.close()
    out.put(mine)


def test_task_queue_hands_every_task_to_exactly_one_process():
    """sharding.TaskQueue (ncb_counter_*: fetch-add on a word in POSIX shared memory): four processes claim from
    one list; every task is taken once, the counter can be reset for another pass, the name disappears with its
    owner."""
    import multiprocessing as mp
    import os
    from nanocompore_b200.sharding import TaskQueue, lpt_order
    name = f"/ncb_test_queue_{os.getpid()}"
    n_tasks = 20000
    owner = TaskQueue(name, n_tasks, create=True)
    ctx = mp.get_context('spawn')
    for _ in range(2):                                   # two passes over the list
        owner.reset()
        out = ctx.Queue()
        procs = [ctx.Process(target=_claim_all, args=(name, n_tasks, out)) for _ in range(4)]
        for p in procs:
            p.start()
        got = [out.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        taken = np.sort(np.concatenate([np.asarray(g, dtype=np.int64) for g in got]))
        assert np.array_equal(taken, np.arange(n_tasks))
    assert owner.claim() is None
    owner.close()
    assert not os.path.exists('/dev/shm' + name)
    with pytest.raises(Exception):
        TaskQueue(name, n_tasks)                         # nobody created it
    assert list(lpt_order([1.0, 5.0, 5.0, 2.0])) == [1, 2, 3, 0]
