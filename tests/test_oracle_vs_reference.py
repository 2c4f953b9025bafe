"""
Pins the oracle (oracle/) against outputs of the UNMODIFIED reference, recorded in
tests/golden/*.npz by tests/golden/make_golden.py (which imports /root/reference in the build
container).  CPU only.

Tolerances and why:
  * float32 shift statistics: the reference reduces in float32 in torch's order; the oracle's
    convention is "exact sum rounded once" -> a few float32 ulps (rtol 2e-6).
  * KS / MW / TT p-values: the reference hands float32 tensors to SciPy 1.18, which then works
    in float32 (SURVEY 8c "float32 trap"); the oracle feeds float64 -> p-values agree to float32
    precision (rtol 2e-5 on p above 1e-30).
  * GMM columns: produced by the reference's code calling oracle/gmm.py ("restated gmm_gpu"),
    so they must match the oracle's own pipeline exactly (counts, LOR) / to 1e-9 (p).
"""
import os

import numpy as np
import pytest

from oracle import prep as oprep
from oracle.comparator import OracleComparator, OracleConfig

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SHIFT = [f"c{c}_{s}_{d}" for c in (1, 2) for s in ('mean', 'median', 'sd') for d in ('intensity', 'dwell')]


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def result_frame(z, prefix):
    cols = [str(c) for c in z[f'{prefix}/__columns__']]
    return {c: z[f'{prefix}/{c}'] for c in cols}


SHIFT_MOTOR = [f"c{c}_{s}_motor_dwell" for c in (1, 2) for s in ('mean', 'median', 'sd')]


def check_frame(df, ref, methods, sample_labels):
    assert len(df) == len(ref['pos'])
    np.testing.assert_array_equal(df['pos'].to_numpy(), ref['pos'])
    for col in SHIFT + [c for c in SHIFT_MOTOR if c in ref]:
        np.testing.assert_allclose(df[col].to_numpy(), ref[col], rtol=2e-6, atol=1e-6, equal_nan=True,
                                   err_msg=col)
    for t in ('KS', 'MW', 'TT'):
        if t not in methods:
            continue
        for v in ('intensity', 'dwell'):
            col = f'{t}_{v}_pvalue'
            got, want = df[col].to_numpy(dtype=np.float64), ref[col].astype(np.float64)
            assert (np.isnan(got) == np.isnan(want)).all(), col
            m = ~np.isnan(want) & (want > 1e-30)
            np.testing.assert_allclose(got[m], want[m], rtol=2e-5, err_msg=col)
    if 'GMM' in methods:
        got, want = df['GMM_chi2_pvalue'].to_numpy(), ref['GMM_chi2_pvalue']
        assert (np.isnan(got) == np.isnan(want)).all()
        np.testing.assert_allclose(got, want, rtol=1e-9, equal_nan=True)
        np.testing.assert_array_equal(df['GMM_LOR'].to_numpy(), ref['GMM_LOR'])
        for label in sample_labels:
            for kind in ('mod', 'unmod'):
                np.testing.assert_array_equal(df[f'{label}_{kind}'].to_numpy(), ref[f'{label}_{kind}'],
                                              err_msg=f'{label}_{kind}')


def test_reference_own_tests_pass_with_restated_gmm():
    import json
    with open(os.path.join(GOLDEN, 'reference_tests.json')) as f:
        rec = json.load(f)
    assert rec['returncode'] == 0 and '14 passed' in rec['summary'], rec


@pytest.mark.parametrize('case', range(6))
def test_synthetic_cases(case):
    """cases 4 and 5 have motor_dwell_offset > 0 (third variable; 2- and 3-variable mixtures)."""
    z = load('synthetic_ref.npz')
    ref = result_frame(z, f'case{case}/result')
    methods = ['GMM', 'KS', 'MW', 'TT']
    motor = int(z[f'case{case}/motor_dwell_offset'])
    assert ('c1_mean_motor_dwell' in ref) == (motor > 0)
    cmp_ = OracleComparator(OracleConfig(comparison_methods=methods, motor_dwell_offset=motor))
    df = cmp_.compare_transcript(case, z[f'case{case}/prepared'], z[f'case{case}/samples'],
                                 z[f'case{case}/conditions'], z[f'case{case}/prepared_positions'])
    check_frame(df, ref, methods, ['kd1', 'kd2', 'wt1', 'wt2'])


def test_prepare_data_matches_reference():
    """run.py:532-601 incl. the float32 log10 (the oracle's convention is fp64 log10 rounded
    once; numpy's float32 log10 may differ in the last bit)."""
    z = load('synthetic_ref.npz')
    for case in range(6):
        data, samples, conditions, positions = oprep.prepare_data(
            z[f'case{case}/raw'], z[f'case{case}/samples'], z[f'case{case}/conditions'],
            int(z[f'case{case}/motor_dwell_offset']))
        data, positions = oprep.filter_low_cov_positions(data, positions, conditions, 1)
        np.testing.assert_array_equal(positions, z[f'case{case}/prepared_positions'])
        want = z[f'case{case}/prepared']
        np.testing.assert_array_equal(data[:, :, 0], want[:, :, 0])
        # libm's float32 log10 is not correctly rounded (only ~40 % of the values are bit-equal to
        # the correctly rounded ones here); the two never differ by more than 2 float32 ulps
        ok = ~np.isnan(want[:, :, 1:])
        assert (np.isnan(data[:, :, 1:]) == ~ok).all()
        ulps = np.abs(data[:, :, 1:][ok].view(np.int32).astype(np.int64) -
                      want[:, :, 1:][ok].view(np.int32).astype(np.int64))
        assert ulps.max() <= 2


def test_fixture_config1():
    """BASELINE config 1: the reference's fixture BAMs, default methods [GMM, KS], min_coverage 1.
    3519 result rows = the 3520-line TSV pinned by the reference's tests/test_run.py:63-64."""
    z = load('fixture_cfg1.npz')
    assert int(z['total_rows']) == 3519
    labels = [str(s) for s in z['sample_labels']]
    sample_ids = dict(zip(labels, [int(i) for i in z['sample_ids']]))
    rows = 0
    for ti in range(int(z['n_transcripts'])):
        ref = result_frame(z, f'tx{ti}/result')
        data, samples, conditions, positions = oprep.prepare_data(
            z[f'tx{ti}/raw'], z[f'tx{ti}/sample_ids'], z[f'tx{ti}/condition_ids'])
        data, positions = oprep.filter_low_cov_positions(data, positions, conditions, 1)
        cmp_ = OracleComparator(OracleConfig(
            comparison_methods=['GMM', 'KS'], sample_ids=sample_ids,
            depleted_condition_id=int(z['depleted_condition_id'])))
        df = cmp_.compare_transcript(ti, data, samples, conditions, positions)
        check_frame(df, ref, ['GMM', 'KS'], labels)
        rows += len(df)
    assert rows == 3519


def test_helpers_match_reference():
    import torch
    from nanocompore_b200 import comparisons as C
    z = load('helpers_ref.npz')
    got = C.nanstd(torch.from_numpy(z['nanstd/in']), 1).numpy()
    np.testing.assert_allclose(got, z['nanstd/out'], rtol=1e-6, equal_nan=True)
    got = C.calculate_lors(torch.from_numpy(z['lor/in'])).numpy()
    np.testing.assert_array_equal(got, z['lor/out'])
    for ctx in (1, 2, 3):
        np.testing.assert_allclose(C.cross_corr_matrix(z['xcorr/in'], ctx), z[f'xcorr/{ctx}'], rtol=1e-12)
    w = C.harmomic_series(2)
    assert w == [1 / 3, 1 / 2, 1, 1 / 2, 1 / 3]
    cor = C.cross_corr_matrix(z['xcorr/in'], 2)
    pv = z['xcorr/in']
    got = np.array([C.combine_pvalues_hou(np.nan_to_num(pv[i:i + 5], nan=1), w, cor) for i in range(50)])
    np.testing.assert_allclose(got, z['hou/out'], rtol=1e-12)
    # oracle/context.py restates the same functions
    from oracle import context as octx
    np.testing.assert_allclose(octx.cross_corr_matrix(pv, 2), cor, rtol=1e-12)
    got = np.array([octx.combine_pvalues_hou(np.nan_to_num(pv[i:i + 5], nan=1), w, cor) for i in range(50)])
    np.testing.assert_allclose(got, z['hou/out'], rtol=1e-12)
