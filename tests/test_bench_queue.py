"""bench.py's end-to-end loop over a shared batch queue (Rig.end_to_end_queue), with stand-in engines: every
batch a rank takes is launched once, consumed once with its (local index, global index) pair and in launch
order per slot, and a slot is never relaunched before its previous batch was consumed.  No GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


class _Status:
    def __init__(self):
        self.a = np.zeros(5, np.int32)

    def numpy(self):
        return self.a


class _Res:
    def __init__(self):
        self.status = _Status()


class _Engine:
    def __init__(self, log, slot):
        self.log, self.slot, self.busy = log, slot, None

    def compare_csr(self, batch, res, stream=None, wait=True):
        assert self.busy is None, "slot relaunched before its batch was consumed"
        self.busy = batch
        self.log.append(('launch', self.slot, batch))

    def wait(self, stream):
        assert self.busy is not None
        self.log.append(('wait', self.slot, self.busy))


def test_every_claimed_batch_is_launched_and_consumed_once():
    log = []
    rig = bench.Rig.__new__(bench.Rig)
    rig.n_slots = 3
    rig.engines = [_Engine(log, s) for s in range(3)]
    rig.streams = [None] * 3
    res = [_Res() for _ in range(3)]
    pool = ['b0', 'b1', 'b2', 'b3']
    handed = iter([7, 2, 9, 4, 11, 13, 20])          # what the shared queue gives THIS rank
    consumed = []

    def consume(slot, i, g):
        consumed.append((slot, i, g))
        rig.engines[slot].busy = None

    tested, taken = rig.end_to_end_queue(pool, lambda: next(handed, None), res, consume)
    assert taken == [7, 2, 9, 4, 11, 13, 20] and tested == 7 * 5
    assert sorted(consumed, key=lambda c: c[1]) == [(i % 3, i, g) for i, g in enumerate(taken)]
    launches = [(s, b) for kind, s, b in log if kind == 'launch']
    assert launches == [(i % 3, pool[g % 4]) for i, g in enumerate(taken)]
    # an empty queue: nothing launched, nothing consumed
    tested, taken = rig.end_to_end_queue(pool, lambda: None, res, consume)
    assert (tested, taken) == (0, [])
