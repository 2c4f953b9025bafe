"""
Records the known-answer cases of the reference's own tests of the prep functions
(/root/reference/tests/test_run.py:70-193: test_prepare_data_no_motor, test_prepare_data_with_motor,
test_downsample) into tests/golden/prep_ref.npz: the UNMODIFIED reference test functions are executed
(their own assertions must hold) with ``Worker._prepare_data`` / ``Worker._downsample`` wrapped by a
recorder, so the file holds the literal inputs of those tests and what the reference returned.

    python tests/golden/make_golden_prep.py          # build container only (/root/reference)

tests/test_prep.py feeds the recorded inputs to nanocompore_b200.prep and oracle.prep.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.refharness import REFERENCE_ROOT, load_reference      # noqa: E402


def load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def to_np(x):
    import torch
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def main():
    load_reference()
    import nanocompore.run as run
    pkg = types.ModuleType('tests')
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'tests')]
    sys.modules['tests'] = pkg
    load_module('tests.common', os.path.join(REFERENCE_ROOT, 'tests', 'common.py'))
    test_run = load_module('tests.test_run', os.path.join(REFERENCE_ROOT, 'tests', 'test_run.py'))

    calls = []

    def record(name):
        orig = getattr(run.Worker, name)

        def wrapper(self, *args):
            ins = [to_np(a).copy() if not np.isscalar(a) else a for a in args]
            out = orig(self, *args)
            calls.append((name, self._conf.get_motor_dwell_offset(), ins, [to_np(o).copy() for o in out]))
            return out
        setattr(run.Worker, name, wrapper)

    record('_prepare_data')
    record('_downsample')
    done = []
    for fn in ('test_prepare_data_no_motor', 'test_prepare_data_with_motor', 'test_downsample'):
        getattr(test_run, fn)()          # the reference's own assertions
        done.append(fn)
    out = {'tests': np.array(done), 'n_calls': np.array(len(calls))}
    for i, (name, motor, ins, outs) in enumerate(calls):
        out[f'call{i}/function'] = np.array(name)
        out[f'call{i}/motor_dwell_offset'] = np.array(motor)
        for j, a in enumerate(ins):
            out[f'call{i}/in{j}'] = np.asarray(a)
        for j, a in enumerate(outs):
            out[f'call{i}/out{j}'] = a
    np.savez_compressed(os.path.join(HERE, 'prep_ref.npz'), **out)
    print('prep_ref:', done, [c[0] for c in calls])


if __name__ == '__main__':
    main()
