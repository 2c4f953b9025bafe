"""
Drop-in proof through the reference's own control flow (missing #3): the UNMODIFIED
``Uncalled4Worker.run`` of the reference (run.py:364-491) processes the reference's fixture BAMs into a
real results database

  * with its own ``TranscriptComparator`` on the CPU (checks the harness: the 3519 rows that
    tests/test_run.py:63-64 pins as a 3520-line TSV), and
  * on a GPU box, with ``nanocompore.run.TranscriptComparator`` replaced by this repo's class and the
    worker's device set to ``cuda:0``: same rows, same k-mers, shift statistics and KS p-values equal to
    the CPU run's, GMM columns present for every row.

Everything else in the loop is the reference's code (tests/ref_worker.py says what is shimmed).
"""
import numpy as np
import pytest

import ref_worker
from oracle import refharness

needs_reference = pytest.mark.skipif(not refharness.reference_available(),
                                     reason="neither /root/reference nor baseline/_ref is present")

SHIFT = [f'c{c}_{s}_{v}' for c in (1, 2) for s in ('mean', 'median', 'sd') for v in ('intensity', 'dwell')]
COUNTS = [f'{s}_{k}' for s in ('kd1', 'kd2', 'wt1', 'wt2') for k in ('mod', 'unmod')]


@needs_reference
def test_reference_worker_loop_with_its_own_comparator(tmp_path):
    cols, messages = ref_worker.run_worker(tmp_path)
    assert len(cols['pos']) == 3519
    assert set(SHIFT + COUNTS + ['GMM_chi2_pvalue', 'GMM_LOR', 'KS_intensity_pvalue', 'KS_dwell_pvalue', 'kmer']) <= set(cols)
    assert not [m for m in messages if m[0] == 'error']


@needs_reference
@pytest.mark.gpu
def test_reference_worker_loop_with_the_b200_comparator(tmp_path):
    from nanocompore_b200.comparisons import TranscriptComparator
    (tmp_path / 'cpu').mkdir()
    (tmp_path / 'gpu').mkdir()
    want, _ = ref_worker.run_worker(tmp_path / 'cpu')
    got, messages = ref_worker.run_worker(tmp_path / 'gpu', comparator_cls=TranscriptComparator, device='cuda:0')
    assert not [m for m in messages if m[0] == 'error'], messages
    # the reference warns about the same positions (comparisons.py:108-120)
    assert len([m for m in messages if m[0] == 'warning' and 'standard deviation' in m[1]]) == 1
    assert len(got['pos']) == 3519
    for key in ('name', 'pos', 'kmer'):
        assert np.array_equal(got[key], want[key]), key
    assert set(want) <= set(got)
    for key in SHIFT:
        np.testing.assert_allclose(got[key], want[key], rtol=3e-6, atol=1e-6, err_msg=key)
    for key in ('KS_intensity_pvalue', 'KS_dwell_pvalue'):
        # the CPU run hands float32 tensors to SciPy 1.18, which then works in float32 (SURVEY 8c,
        # "float32 trap"); the kernels give the float64 value
        np.testing.assert_allclose(got[key], want[key], rtol=2e-6, err_msg=key)
    # mixture columns: "reference code, restated gmm_gpu" on the CPU side, the kernels' fp64 EM here
    gate = np.isnan(got['GMM_chi2_pvalue']) == np.isnan(want['GMM_chi2_pvalue'])
    assert gate.mean() > 0.97
    for key in COUNTS:
        assert (got[key] >= 0).all() and got[key].sum() > 0


@needs_reference
def test_cpu_device_is_refused():
    """The reference's default is ``devices: {'cpu': 2}``: this implementation has no CPU path and says
    so instead of running such a worker on GPU 0."""
    import torch
    from nanocompore_b200.comparisons import NanocomporeError, TranscriptComparator
    ref = refharness.load_reference()
    conf = refharness.reference_config(ref)
    cmp_ = TranscriptComparator(conf, refharness.RefWorker())
    data = torch.zeros((2, 4, 2))
    ids = torch.zeros(4, dtype=torch.int16)
    with pytest.raises(NanocomporeError, match='CUDA'):
        cmp_.compare_transcript(object(), data, ids, ids, torch.arange(2), 'cpu')
