"""
The drop-in boundary: nanocompore_b200.comparisons.TranscriptComparator.compare_transcript on a
GPU against result frames recorded from the UNMODIFIED reference (tests/golden/*.npz), and the
behaviours the reference's own tests pin at this boundary (tests/test_comparisons.py,
tests/test_run.py:63-64).
"""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from nanocompore_b200.comparisons import TranscriptComparator
from nanocompore_b200.config import PathConfig
from oracle import prep as oprep
from test_oracle_vs_reference import SHIFT, SHIFT_MOTOR, load, result_frame

pytestmark = pytest.mark.gpu

DATA = {'kd': {'kd1': {}, 'kd2': {}}, 'wt': {'wt1': {}, 'wt2': {}}}


class Transcript:
    def __init__(self, id, name):
        self.id, self.name = id, name


class Worker:
    def __init__(self):
        self.messages = []

    def log(self, level, msg):
        self.messages.append((level, msg))


def config(**kw):
    return PathConfig(dict(data=DATA, depleted_condition='kd', **kw))


def compare(cfg, data, samples, conditions, positions, worker=None, **kw):
    cmp_ = TranscriptComparator(cfg, worker or Worker(), **kw)
    return cmp_.compare_transcript(Transcript(3, 'tx'), torch.from_numpy(np.ascontiguousarray(data)),
                                   torch.as_tensor(samples), torch.as_tensor(conditions),
                                   torch.as_tensor(positions), 'cuda:0')[1]


def check_against_reference(df, ref, methods, labels):
    assert len(df) == len(ref['pos'])
    np.testing.assert_array_equal(df['pos'].to_numpy(), ref['pos'])
    for col in SHIFT + [c for c in SHIFT_MOTOR if c in ref]:
        np.testing.assert_allclose(df[col].to_numpy(), ref[col], rtol=2e-6, atol=1e-6, equal_nan=True, err_msg=col)
    for t in ('KS', 'MW', 'TT'):
        if t in methods:
            for v in ('intensity', 'dwell'):
                col = f'{t}_{v}_pvalue'
                got, want = df[col].to_numpy(dtype=np.float64), ref[col].astype(np.float64)
                assert (np.isnan(got) == np.isnan(want)).all(), col
                m = ~np.isnan(want) & (want > 1e-30)
                # the reference's SciPy ran in float32 here, hence 2e-5
                np.testing.assert_allclose(got[m], want[m], rtol=2e-5, err_msg=col)
    if 'GMM' in methods:
        got, want = df['GMM_chi2_pvalue'].to_numpy(), ref['GMM_chi2_pvalue']
        assert (np.isnan(got) == np.isnan(want)).all()
        np.testing.assert_allclose(got, want, rtol=1e-9, equal_nan=True)
        np.testing.assert_array_equal(df['GMM_LOR'].to_numpy(), ref['GMM_LOR'])
        for label in labels:
            for kind in ('mod', 'unmod'):
                np.testing.assert_array_equal(df[f'{label}_{kind}'].to_numpy(), ref[f'{label}_{kind}'])
                assert df[f'{label}_{kind}'].dtype == np.uint32


def test_fixture_config1_rows_and_values():
    """BASELINE config 1: reference fixture BAMs, [GMM, KS], min_coverage 1 -> 3519 rows."""
    z = load('fixture_cfg1.npz')
    labels = [str(s) for s in z['sample_labels']]
    rows = 0
    for ti in range(int(z['n_transcripts'])):
        ref = result_frame(z, f'tx{ti}/result')
        data, samples, conditions, positions = oprep.prepare_data(
            z[f'tx{ti}/raw'], z[f'tx{ti}/sample_ids'], z[f'tx{ti}/condition_ids'])
        data, positions = oprep.filter_low_cov_positions(data, positions, conditions, 1)
        df = compare(config(), data, samples, conditions, positions)
        check_against_reference(df, ref, ['GMM', 'KS'], labels)
        assert list(df.columns) == [str(c) for c in z[f'tx{ti}/result/__columns__']] or \
            set(df.columns) == set(str(c) for c in z[f'tx{ti}/result/__columns__'])
        rows += len(df)
    assert rows == 3519


def test_fixture_config1_against_golden_db():
    """The CUDA result frame against the numbers a real nanocompore run (real SciPy in float64, older
    build: its condition 1 is today's condition 2) stored in the reference's golden results database
    (tests/golden/golden_db.npz, see tests/test_oracle_golden_db.py for what it can pin): KS p-values
    to 1e-9 and shift statistics to 2e-6 on >= 99 % of the 3437 rows (the rest predate a rule of the
    current reader), per-sample read totals of the mixture fit on >= 99.5 %."""
    z = load('fixture_cfg1.npz')
    g = load('golden_db.npz')
    cols = [str(c) for c in g['columns']]
    db_tx = [str(t) for t in g['transcripts']]
    n = ks_ok = shift_ok = reads_ok = 0
    for ti in range(int(z['n_transcripts'])):
        db = g[f"tx{db_tx.index(str(z[f'tx{ti}/name']))}"]
        data, samples, conditions, positions = oprep.prepare_data(
            z[f'tx{ti}/raw'], z[f'tx{ti}/sample_ids'], z[f'tx{ti}/condition_ids'])
        data, positions = oprep.filter_low_cov_positions(data, positions, conditions, 1)
        df = compare(config(), data, samples, conditions, positions)
        pos = df['pos'].to_numpy()
        row = {int(p): i for i, p in enumerate(db[:, 0])}
        keep = np.array([i for i, p in enumerate(pos) if int(p) in row])
        d = db[[row[int(pos[i])] for i in keep]]
        col = lambda c: d[:, cols.index(c)]                          # noqa: E731
        ks = np.ones(len(keep), dtype=bool)
        for v in ('intensity', 'dwell'):
            ks &= np.isclose(df[f'KS_{v}_pvalue'].to_numpy(dtype=np.float64)[keep], col(f'KS_{v}_pvalue'),
                             rtol=1e-9, atol=0, equal_nan=True)
        sh = np.ones(len(keep), dtype=bool)
        for ours_c, theirs_c in ((1, 2), (2, 1)):
            for s in ('mean', 'median', 'sd'):
                for v in ('intensity', 'dwell'):
                    sh &= np.isclose(df[f'c{ours_c}_{s}_{v}'].to_numpy(dtype=np.float64)[keep],
                                     col(f'c{theirs_c}_{s}_{v}'), rtol=2e-6, atol=1e-6, equal_nan=True)
        ours = np.stack([df[f'{s}_{k}'].to_numpy()[keep].astype(np.float64) for s in ('wt1', 'wt2', 'kd1', 'kd2')
                         for k in ('mod', 'unmod')], 1)
        theirs = np.stack([col(f'{s}_{k}') for s in ('WT1', 'WT2', 'KD1', 'KD2') for k in ('mod', 'unmod')], 1)
        n += len(keep)
        ks_ok += int(ks.sum())
        shift_ok += int(sh.sum())
        reads_ok += int((ours.reshape(-1, 4, 2).sum(2) == theirs.reshape(-1, 4, 2).sum(2)).all(1).sum())
    assert n == 3437
    assert ks_ok >= 0.99 * n and shift_ok >= 0.99 * n and reads_ok >= 0.995 * n, (ks_ok, shift_ok, reads_ok, n)


@pytest.mark.parametrize('case', range(6))
def test_synthetic_against_reference(case):
    """cases 4, 5: motor_dwell_offset > 0 -- 18 shift statistics, 2- and 3-variable mixtures."""
    z = load('synthetic_ref.npz')
    ref = result_frame(z, f'case{case}/result')
    methods = ['GMM', 'KS', 'MW', 'TT']
    motor = int(z[f'case{case}/motor_dwell_offset'])
    df = compare(config(comparison_methods=methods, motor_dwell_offset=motor), z[f'case{case}/prepared'],
                 z[f'case{case}/samples'], z[f'case{case}/conditions'], z[f'case{case}/prepared_positions'])
    assert ('c2_sd_motor_dwell' in df.columns) == (motor > 0)
    check_against_reference(df, ref, methods, ['kd1', 'kd2', 'wt1', 'wt2'])


def test_no_positions_returns_none():
    cmp_ = TranscriptComparator(config(), Worker())
    t = Transcript(1, 'x')
    out = cmp_.compare_transcript(t, torch.zeros((0, 4, 2)), torch.zeros(4), torch.zeros(4),
                                  torch.zeros(0, dtype=torch.int16), 'cuda:0')
    assert out == (t, None)


def test_columns_dtypes_and_float64_input():
    # the reference's tests hand float64 tensors on the CPU (tests/test_comparisons.py:145, 469)
    rng = np.random.default_rng(42)
    data = rng.standard_normal((5, 100, 2))
    data[:, 50:, :] += 3
    samples = np.repeat([0, 1, 2, 3], 25)
    conditions = np.repeat([0, 1], 50)
    df = compare(config(), data.astype(np.float64), samples, conditions, np.arange(5, dtype=np.int16))
    assert list(df.columns[:2]) == ['transcript_id', 'pos'] and list(df.columns[2:14]) == SHIFT
    assert list(df.columns[14:]) == ['GMM_chi2_pvalue', 'GMM_LOR', 'kd1_mod', 'kd1_unmod', 'kd2_mod',
                                     'kd2_unmod', 'wt1_mod', 'wt1_unmod', 'wt2_mod', 'wt2_unmod',
                                     'KS_intensity_pvalue', 'KS_dwell_pvalue']
    assert df['GMM_chi2_pvalue'].dtype == np.float64 and df['GMM_LOR'].dtype == np.float32
    assert df['c1_mean_intensity'].dtype == np.float32
    # two blobs 3 sigma apart: tests/test_comparisons.py:487-504
    assert (df['GMM_chi2_pvalue'] < 0.01).all()
    assert (df['kd1_mod'] + df['kd1_unmod'] == 25).all()


def test_single_component_gives_nan_pvalue():
    # tests/test_comparisons.py:464-484
    rng = np.random.default_rng(42)
    data = rng.standard_normal((1, 100, 2))
    df = compare(config(comparison_methods=['GMM']), data, np.repeat([0, 1, 2, 3], 25),
                 np.repeat([0, 1], 50), np.array([7], dtype=np.int16))
    assert np.isnan(df['GMM_chi2_pvalue'].iloc[0]) and np.isnan(df['GMM_LOR'].iloc[0])


def test_bad_std_positions_are_dropped_with_a_warning():
    rng = np.random.default_rng(1)
    data = rng.standard_normal((6, 40, 2)).astype(np.float32)
    data[2, :, 0] = 5.0                 # zero variance -> std 0 -> dropped (comparisons.py:108-120)
    w = Worker()
    df = compare(config(comparison_methods=['KS']), data, np.repeat([0, 1, 2, 3], 10),
                 np.repeat([0, 1], 20), np.arange(10, 16, dtype=np.int16), worker=w)
    assert df['pos'].tolist() == [10, 11, 13, 14, 15]
    assert any(level == 'warning' and '[12]' in msg for level, msg in w.messages)


def test_auto_routes_by_coverage():
    rng = np.random.default_rng(2)
    n = 520
    data = rng.standard_normal((4, 2 * n, 2)).astype(np.float32)
    data[:2, 50:n, :] = np.nan          # positions 0,1: 50 + 50 reads -> KS;
    data[:2, n + 50:, :] = np.nan       # positions 2,3: 1040 reads -> GMM
    df = compare(config(comparison_methods=['auto']), data, np.repeat([0, 1, 2, 3], n // 2),
                 np.repeat([0, 1], n), np.arange(4, dtype=np.int16))
    assert df['auto_test'].tolist() == ['KS', 'KS', 'GMM', 'GMM']
    assert df['KS_intensity_pvalue'].notna().tolist() == [True, True, False, False]
    assert (df['auto_pvalue'][:2].to_numpy() == df['KS_intensity_pvalue'][:2].to_numpy()).all()
    ok = df['GMM_chi2_pvalue'].notna().to_numpy()
    assert (df['auto_pvalue'].to_numpy()[2:][ok[2:]] == df['GMM_chi2_pvalue'].to_numpy()[2:][ok[2:]]).all()


def test_sequence_context_columns_match_the_oracle():
    from oracle.context import combine_context_pvalues
    z = load('synthetic_ref.npz')
    df = compare(config(comparison_methods=['KS'], sequence_context=2, sequence_context_weights='harmonic'),
                 z['case0/prepared'], z['case0/samples'], z['case0/conditions'], z['case0/prepared_positions'])
    assert list(df.columns[-2:]) == ['KS_intensity_pvalue_context_2', 'KS_dwell_pvalue_context_2']
    for var in ('intensity', 'dwell'):
        want = combine_context_pvalues(df['pos'].to_numpy(), df[f'KS_{var}_pvalue'].to_numpy(), 2, 'harmonic')
        got = df[f'KS_{var}_pvalue_context_2'].to_numpy()
        assert (np.isnan(got) == np.isnan(want)).all()
        m = ~np.isnan(want)
        np.testing.assert_allclose(np.log10(got[m]), np.log10(want[m]), rtol=1e-5, atol=1e-8)


def test_unsupported_options_raise():
    with pytest.raises(NotImplementedError):
        compare(config(comparison_methods=['GOF']), np.zeros((2, 8, 2), np.float32), np.zeros(8), np.zeros(8),
                np.arange(2, dtype=np.int16))


def test_batch_call_equals_per_transcript_calls():
    """compare_transcripts: one packed batch, one libncb call, same frames (fixture config 1 + synthetic)."""
    z = load('fixture_cfg1.npz')
    items = []
    for ti in range(int(z['n_transcripts'])):
        data, samples, conditions, positions = oprep.prepare_data(
            z[f'tx{ti}/raw'], z[f'tx{ti}/sample_ids'], z[f'tx{ti}/condition_ids'])
        data, positions = oprep.filter_low_cov_positions(data, positions, conditions, 1)
        items.append((Transcript(ti, f'tx{ti}'), torch.from_numpy(data), torch.as_tensor(samples),
                      torch.as_tensor(conditions), torch.as_tensor(positions)))
    s = load('synthetic_ref.npz')
    items.append((Transcript(7, 'syn'), torch.from_numpy(s['case1/prepared']), torch.as_tensor(s['case1/samples']),
                  torch.as_tensor(s['case1/conditions']), torch.as_tensor(s['case1/prepared_positions'])))
    items.append((Transcript(8, 'empty'), torch.zeros((0, 4, 2)), torch.zeros(4), torch.zeros(4),
                  torch.zeros(0, dtype=torch.int16)))
    cmp_ = TranscriptComparator(config(), Worker())
    batch = cmp_.compare_transcripts(items, 'cuda:0')
    assert batch[-1][1] is None and len(batch) == len(items)
    rows = 0
    for (tr, df), item in zip(batch[:-1], items[:-1]):
        one = cmp_.compare_transcript(*item, 'cuda:0')[1]
        assert tr is item[0] and list(df.columns) == list(one.columns)
        pd.testing.assert_frame_equal(df.reset_index(drop=True), one.reset_index(drop=True))
        rows += len(df)
    assert rows == 3519 + len(s['case1/result/pos'])
