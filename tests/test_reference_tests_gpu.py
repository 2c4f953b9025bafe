"""
The reference's OWN tests of the path (tests/test_comparisons.py of nanocompore 2.0.0, unmodified, from
/root/reference or the staged copy baseline/_ref) executed against THIS repo's
``nanocompore_b200.comparisons`` -- ``TranscriptComparator`` and the module helpers
(``calculate_lors``, ``get_contingency_matrices``, ``get_soft_contingency_matrices``): the module is
loaded with ``nanocompore.comparisons`` resolving to the B200 mirror, and the device string the tests
pass (``'cpu'``) rewritten to ``'cuda:0'`` -- this implementation has no CPU path.  Nothing else of the
test source changes.  Config, MockWorker and the fixtures' paths are the reference's.
"""
import importlib.util
import os
import sys
import types

import pytest

from oracle import refharness

REF_TESTS = None if refharness.reference_root() is None else os.path.join(refharness.reference_root(), 'tests')
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(REF_TESTS is None or not os.path.isfile(os.path.join(REF_TESTS, 'test_comparisons.py')),
                                 reason="the reference's tests are not present (python baseline/stage_reference.py)")]

REFERENCE_TESTS = [
    'test_add_shift_stats', 'test_add_shift_stats_with_motor_dwell', 'test_nonparametric_test_KS',
    'test_combine_context_pvalues', 'test_calculate_lor', 'test_get_contigency_matrices',
    'test_get_contigency_matrices_with_gaps', 'test_get_soft_contigency_matrices',
    'test_get_soft_contigency_matrices_with_gaps', 'test_get_cluster_counts', 'test_get_soft_cluster_counts',
    'test_gmm_test', 'test_gmm_test_split_single_component', 'test_gmm_test_split_two_components',
]


def load_reference_test_module(comparisons_module, device):
    """tests/test_comparisons.py of the reference as a module whose ``nanocompore.comparisons`` is
    ``comparisons_module``."""
    refharness.load_reference()
    saved = {k: sys.modules.get(k) for k in ('nanocompore.comparisons', 'tests', 'tests.common')}
    try:
        sys.modules['nanocompore.comparisons'] = comparisons_module
        pkg = types.ModuleType('tests')
        pkg.__path__ = [REF_TESTS]
        sys.modules['tests'] = pkg
        spec = importlib.util.spec_from_file_location('tests.common', os.path.join(REF_TESTS, 'common.py'))
        common = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(common)
        sys.modules['tests.common'] = common
        with open(os.path.join(REF_TESTS, 'test_comparisons.py')) as f:
            source = f.read()
        assert source.count("'cpu'") >= 4
        source = source.replace("'cpu'", repr(device))
        mod = types.ModuleType('reference_test_comparisons')
        mod.__file__ = os.path.join(REF_TESTS, 'test_comparisons.py')
        exec(compile(source, mod.__file__, 'exec'), mod.__dict__)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.fixture(scope='module')
def reference_tests():
    import nanocompore_b200.comparisons as ours
    mod = load_reference_test_module(ours, 'cuda:0')
    assert mod.TranscriptComparator is ours.TranscriptComparator
    return mod


def test_every_reference_test_of_the_path_is_listed(reference_tests):
    found = sorted(n for n in vars(reference_tests) if n.startswith('test_'))
    assert found == sorted(REFERENCE_TESTS), found


@pytest.mark.parametrize('name', REFERENCE_TESTS)
def test_reference_test_passes_on_the_b200_comparator(reference_tests, name):
    getattr(reference_tests, name)()
