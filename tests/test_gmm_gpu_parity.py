"""
Mixture columns against the reference's REAL fitter, gmm-gpu (pyproject.toml:29 of nanocompore 2.0.0:
``gmm-gpu==0.2.8``).  The package is neither vendored under /root/reference nor installed in this image
and cannot be fetched (no network), so in this repository's own runs every test of this file is
SKIPPED -- which is what "parity unpinned against gmm_gpu" in DESIGN.md section 3 means.  Wherever the
package is importable (a user's nanocompore environment with a GPU) the file turns into the missing pin:

  * the unmodified reference ``TranscriptComparator._gmm_test_split`` (comparisons.py:338-392) with the
    real ``gmm_gpu.gmm.GMM`` on the synthetic fixtures of the parity tests,
  * against ``nanocompore_b200.comparisons.TranscriptComparator`` on the same tensors,
  * on the BIC gate (NaN pattern of ``GMM_chi2_pvalue``), the per-sample cluster counts, ``GMM_LOR`` and the
    chi-squared p-value.

gmm_gpu starts its EM from a seeded random initialisation (run.py:774-780, comparisons.py:341-345); this
implementation from a deterministic 2-means start (oracle/gmm.py).  Two EM runs from different starts
agree where the likelihood has one dominant mode, so the bars are: on well-separated two-cluster
positions the gate and the partition agree on >= 95 % of positions, and wherever the two partitions ARE
equal the LOR and the p-value agree to 1e-6 (they are functions of the counts alone).  The measured rates
are printed for the record whatever they are.
"""
import json

import numpy as np
import pytest

gmm_module = pytest.importorskip(
    'gmm_gpu.gmm', reason="gmm-gpu (the reference's mixture fitter) is not installed: GMM parity with the "
                          "reference stays unpinned (DESIGN.md section 3)")

from oracle import refharness  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refharness.reference_available(),
                                 reason="the reference's code is not present (python baseline/stage_reference.py)")]


def _fixture(seed, n_positions, reads, separation):
    """Standardised-scale two-condition data with a modified cluster in condition 1 on half of the
    positions: (P, R, 2) float32, samples, conditions."""
    rng = np.random.default_rng(seed)
    R = 4 * reads
    samples = np.repeat(np.arange(4), reads)
    conditions = (samples >= 2).astype(np.int64)
    data = rng.normal(0.0, 1.0, (n_positions, R, 2)).astype(np.float32)
    modified = np.arange(n_positions) % 2 == 0
    shift = (rng.random((n_positions, R)) < 0.6) & (conditions == 1)[None, :] & modified[:, None]
    data[..., 0] += np.where(shift, separation, 0.0).astype(np.float32)
    data[..., 1] += np.where(shift, 0.5 * separation, 0.0).astype(np.float32)
    return data, samples, conditions


def _reference_columns(data, samples, conditions):
    import torch
    ref = refharness.load_reference()
    ref.comparisons.GMM = gmm_module.GMM         # the harness binds oracle.gmm.GMM when the package is absent
    conf = refharness.reference_config(ref, comparison_methods=['GMM'])
    cmp_ = ref.comparisons.TranscriptComparator(config=conf, worker=refharness.RefWorker())
    return cmp_._gmm_test_split(torch.from_numpy(data).to('cuda:0'), torch.from_numpy(samples).to('cuda:0'),
                                torch.from_numpy(conditions).to('cuda:0'), 'cuda:0')


def _our_columns(data, samples, conditions):
    import torch
    from nanocompore_b200.comparisons import TranscriptComparator
    from nanocompore_b200.config import PathConfig
    conf = PathConfig({'data': {'kd': {'kd1': {'bam': ''}, 'kd2': {'bam': ''}},
                                'wt': {'wt1': {'bam': ''}, 'wt2': {'bam': ''}}},
                       'depleted_condition': 'kd', 'comparison_methods': ['GMM']})
    cmp_ = TranscriptComparator(config=conf, worker=refharness.RefWorker())
    return cmp_._gmm_test_split(torch.from_numpy(data).to('cuda:0'), torch.from_numpy(samples).to('cuda:0'),
                                torch.from_numpy(conditions).to('cuda:0'), 'cuda:0')


@pytest.mark.parametrize('reads,separation', [(60, 4.0), (100, 3.0)])
def test_mixture_columns_against_gmm_gpu(reads, separation):
    data, samples, conditions = _fixture(20260 + reads, 200, reads, separation)
    want = _reference_columns(data, samples, conditions)
    got = _our_columns(data, samples, conditions)
    wp, gp = np.asarray(want['GMM_chi2_pvalue'], dtype=np.float64), np.asarray(got['GMM_chi2_pvalue'], dtype=np.float64)
    gate_same = np.isnan(wp) == np.isnan(gp)
    count_keys = [k for k in want if k.endswith('_mod') or k.endswith('_unmod')]
    assert count_keys, "the reference returned no cluster counts"
    same_counts = np.ones(wp.size, dtype=bool)
    for k in count_keys:
        same_counts &= np.asarray(want[k]) == np.asarray(got[k])
    both = ~np.isnan(wp) & ~np.isnan(gp) & same_counts
    print(json.dumps({"reads_per_sample": reads, "separation": separation, "positions": int(wp.size),
                      "bic_gate_agreement": float(gate_same.mean()),
                      "partition_agreement": float(same_counts.mean()),
                      "positions_compared_on_p_and_lor": int(both.sum())}))
    assert gate_same.mean() >= 0.95
    modified = np.arange(wp.size) % 2 == 0
    assert same_counts[modified].mean() >= 0.95
    assert np.allclose(np.log10(gp[both]), np.log10(wp[both]), rtol=1e-6, atol=1e-8)
    assert np.allclose(np.asarray(got['GMM_LOR'], dtype=np.float64)[both],
                       np.asarray(want['GMM_LOR'], dtype=np.float64)[both], rtol=0, atol=2e-3)   # the reference rounds the LOR to 3 decimals
