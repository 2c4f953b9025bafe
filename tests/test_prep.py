"""nanocompore_b200.prep against the reference's recorded outputs (tests/golden) and the cases
of the reference's own tests (tests/test_run.py:70-238)."""
import numpy as np
import torch

from nanocompore_b200 import prep
from oracle import prep as oprep
from test_oracle_vs_reference import load


def test_prepare_and_filter_match_the_reference_goldens():
    z = load('synthetic_ref.npz')
    for case in range(6):
        motor = int(z[f'case{case}/motor_dwell_offset'])
        t, s, c, pos = prep.prepare_data(z[f'case{case}/raw'].copy(), z[f'case{case}/samples'],
                                         z[f'case{case}/conditions'], motor)
        t, pos = prep.filter_low_cov_positions(t, pos, c, 1)
        assert t.dtype == torch.float32 and s.dtype == torch.int16 and pos.dtype == torch.int16
        np.testing.assert_array_equal(pos.numpy(), z[f'case{case}/prepared_positions'])
        want = z[f'case{case}/prepared']
        np.testing.assert_array_equal(t.numpy()[:, :, 0], want[:, :, 0])
        ok = ~np.isnan(want[:, :, 1:])
        ulps = np.abs(t.numpy()[:, :, 1:][ok].view(np.int32).astype(np.int64) - want[:, :, 1:][ok].view(np.int32).astype(np.int64))
        assert ulps.max() <= 2            # libm's float32 log10 vs the correctly rounded one
        # and bit-equal to the oracle's convention
        od, _, _, opos = oprep.prepare_data(z[f'case{case}/raw'], z[f'case{case}/samples'], z[f'case{case}/conditions'], motor)
        od, opos = oprep.filter_low_cov_positions(od, opos, z[f'case{case}/conditions'], 1)
        assert np.array_equal(t.numpy(), od, equal_nan=True)


def test_prepare_data_motor_dwell_column():
    # tests/test_run.py:113-150: the motor column is the dwell `offset` positions downstream
    data = np.full((6, 2, 2), np.nan, dtype=np.float32)
    data[:, :, 0] = 100.0
    data[:, 0, 1] = [0.01, 0.02, 0.03, 0.04, 0.05, 0.06]
    t, s, c, pos = prep.prepare_data(data.copy(), [0, 1], [0, 1], motor_dwell_offset=2)
    assert t.shape == (6, 2, 3)
    np.testing.assert_allclose(t[:4, 0, 2].numpy(), np.log10(np.array([0.03, 0.04, 0.05, 0.06]) + 1e-10), rtol=1e-6)
    assert torch.isnan(t[4:, 0, 2]).all() and torch.isnan(t[:, 1, 2]).all()
    od = oprep.prepare_data(data, [0, 1], [0, 1], 2)[0]
    assert np.array_equal(t.numpy(), od, equal_nan=True)


def test_filter_low_cov_positions():
    # tests/test_run.py:196-238
    data = torch.full((3, 6, 2), float('nan'))
    cond = torch.tensor([0, 0, 0, 1, 1, 1])
    data[0, :, :] = 1.0                  # 3 + 3
    data[1, [0, 1, 3], :] = 1.0          # 2 + 1
    data[2, [0, 1, 2, 3, 4], :] = 1.0    # 3 + 2
    out, pos = prep.filter_low_cov_positions(data, torch.tensor([5, 6, 7]), cond, 2)
    assert pos.tolist() == [5, 7] and out.shape[0] == 2


def test_downsample_keeps_the_most_complete_reads_per_condition():
    # tests/test_run.py:153-193
    rng = np.random.default_rng(0)
    P, R = 20, 12
    data = torch.from_numpy(rng.standard_normal((P, R, 2)).astype(np.float32))
    cond = torch.tensor([0] * 6 + [1] * 6)
    samples = torch.arange(R) % 4
    holes = [0, 3, 1, 7, 2, 9, 5, 0, 4, 4, 8, 1]     # missing positions per read
    for r, h in enumerate(holes):
        data[:h, r, 0] = float('nan')
    out, s, c = prep.downsample(data, samples, cond, 3)
    assert out.shape == (P, 6, 2) and c.tolist() == [0, 0, 0, 1, 1, 1]
    kept = [r for r in range(R) if any(torch.equal(torch.nan_to_num(out[:, k]), torch.nan_to_num(data[:, r])) for k in range(6))]
    assert kept == [0, 2, 4, 7, 8, 11]        # fewest holes per condition; the tie 4/4 keeps read order
    od, os_, oc = oprep.downsample(data.numpy(), samples.numpy(), cond.numpy(), 3)
    assert np.array_equal(out.numpy(), od, equal_nan=True) and (oc == c.numpy()).all()
    same, s2, c2 = prep.downsample(data, samples, cond, 12)
    assert same is data


def test_reference_own_known_answers():
    """The literal inputs of the reference's tests/test_run.py:70-193 (test_prepare_data_no_motor,
    test_prepare_data_with_motor, test_downsample) and what the UNMODIFIED reference returned for
    them, recorded by tests/golden/make_golden_prep.py while those tests passed."""
    z = load('prep_ref.npz')
    assert [str(t) for t in z['tests']] == ['test_prepare_data_no_motor', 'test_prepare_data_with_motor',
                                            'test_downsample']
    seen = []
    for i in range(int(z['n_calls'])):
        fn = str(z[f'call{i}/function'])
        motor = int(z[f'call{i}/motor_dwell_offset'])
        want = [z[f'call{i}/out{j}'] for j in range(4 if fn == '_prepare_data' else 3)]
        if fn == '_prepare_data':
            args = [z[f'call{i}/in{j}'] for j in range(3)]
            got = prep.prepare_data(args[0].copy(), args[1], args[2], motor)
            ora = oprep.prepare_data(args[0].copy(), args[1], args[2], motor)
        else:
            args = [z[f'call{i}/in{j}'] for j in range(3)]
            max_reads = int(z[f'call{i}/in3'])
            got = prep.downsample(torch.from_numpy(args[0]), torch.from_numpy(args[1]), torch.from_numpy(args[2]),
                                  max_reads)
            ora = oprep.downsample(args[0], args[1], args[2], max_reads)
        assert len(got) == len(want) == len(ora)
        for g, o, w in zip(got, ora, want):
            g = g.numpy() if isinstance(g, torch.Tensor) else np.asarray(g)
            assert g.shape == w.shape and np.asarray(o).shape == w.shape, fn
            if w.dtype.kind == 'f':
                # log10 in float32: libm (reference) vs correctly rounded (here), <= 2 ulps
                np.testing.assert_allclose(g, w, rtol=3e-7, atol=1e-7, equal_nan=True, err_msg=fn)
                np.testing.assert_allclose(np.asarray(o), w, rtol=3e-7, atol=1e-7, equal_nan=True, err_msg=fn)
            else:
                np.testing.assert_array_equal(g, w, err_msg=fn)
                np.testing.assert_array_equal(np.asarray(o), w, err_msg=fn)
        seen.append((fn, motor))
    assert seen == [('_prepare_data', 0), ('_prepare_data', 2), ('_downsample', 0)]
