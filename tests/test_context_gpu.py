"""Sequence-context combination on the GPU (``ncb_context_combine``, csrc/context.cu) against the
oracle restatement of comparisons.py:436-474, 627-705 (oracle/context.py; pinned against the
reference's own functions and its v1 golden numbers in tests/test_oracle_vs_reference.py).

Tolerance: 1e-5 relative on log10 p (1e-8 absolute near p = 1), the p-value tolerance of the path;
the kernel sums in a different order than NumPy, so it is not bitwise."""
import numpy as np
import pytest
import torch

from nanocompore_b200 import _lib as L
from oracle.context import combine_context_pvalues, combine_pvalues_hou, cross_corr_matrix, harmonic_series

pytestmark = pytest.mark.gpu

TINY = np.finfo(np.float64).tiny


def assert_close_log10(got, want, rtol=1e-5, atol=1e-8):
    got, want = np.asarray(got, float), np.asarray(want, float)
    assert (np.isnan(got) == np.isnan(want)).all()
    m = ~np.isnan(want)
    lg, lw = np.log10(got[m]), np.log10(want[m])
    assert (np.abs(lg - lw) <= atol + rtol * np.abs(lw)).all(), np.abs(lg - lw).max()


def make_track(rng, n_rows, span, nan_frac=0.1, small_frac=0.05):
    pos = np.sort(rng.choice(np.arange(100, 100 + span), size=n_rows, replace=False))
    p = rng.uniform(1e-6, 1.0, n_rows)
    small = rng.random(n_rows) < small_frac
    p[small] = 10.0 ** rng.uniform(-300, -3, small.sum())
    p[rng.random(n_rows) < nan_frac] = np.nan
    return pos, p


@pytest.mark.parametrize('context', [1, 2, 3, 5, 8])
@pytest.mark.parametrize('mode', ['uniform', 'harmonic'])
def test_context_combination_matches_the_oracle(engine_factory, context, mode):
    eng = engine_factory(tests=('KS',))
    rng = np.random.default_rng(100 + context)
    tracks = [make_track(rng, n, span) for n, span in
              [(400, 400), (350, 900), (60, 64), (3 * context + 3, 3 * context + 3), (1500, 2200)]]
    offsets = np.concatenate([[0], np.cumsum([len(t[0]) for t in tracks])])
    pos = np.concatenate([t[0] for t in tracks])
    cols = np.stack([np.concatenate([t[1] for t in tracks]),
                     np.concatenate([np.where(np.isnan(t[1]), 0.5, t[1] ** 0.5) for t in tracks])])
    weights = harmonic_series(context) if mode == 'harmonic' else [1] * (2 * context + 1)
    got, status = eng.context_combine(offsets, pos, cols, context, weights)
    assert (status == L.NCB_CTX_OK).all()
    for c in range(cols.shape[0]):
        for t in range(len(tracks)):
            sl = slice(offsets[t], offsets[t + 1])
            want = combine_context_pvalues(pos[sl], cols[c, sl], context, mode)
            assert_close_log10(got[c, sl], want)


def test_rows_in_any_order_and_device_memory(engine_factory):
    eng = engine_factory(tests=('KS',))
    rng = np.random.default_rng(7)
    pos, p = make_track(rng, 300, 500)
    perm = rng.permutation(pos.size)
    want = combine_context_pvalues(pos, p, 2, 'harmonic')
    w = harmonic_series(2)
    got, status = eng.context_combine([0, pos.size], pos[perm], p[perm], 2, w)
    assert status.shape == (1, 1) and status[0, 0] == L.NCB_CTX_OK
    assert_close_log10(got[0], want[perm])
    dev = torch.from_numpy(p[perm]).cuda()
    got_dev, _ = eng.context_combine([0, pos.size], pos[perm], dev, 2, w)
    assert got_dev.is_cuda
    np.testing.assert_array_equal(got_dev.cpu().numpy()[0], got[0])


def test_all_ones_underflow_and_nan_centres(engine_factory):
    eng = engine_factory(tests=('KS',))
    pos = np.arange(50)
    ones = np.ones(50)
    ones[[3, 17]] = np.nan
    # a window of ones combines to 1; a NaN centre stays NaN (comparisons.py:464-465, 675-676)
    got, status = eng.context_combine([0, 50], pos, ones, 2, [1] * 5)
    assert status[0, 0] == L.NCB_CTX_OK
    assert np.isnan(got[0, [3, 17]]).all() and (np.delete(got[0], [3, 17]) == 1.0).all()
    # chi2.sf underflows to 0 -> the smallest normal number (comparisons.py:702-704)
    rng = np.random.default_rng(1)
    p = rng.uniform(0.01, 1.0, 50)
    p[20:25] = 1e-300
    want = combine_context_pvalues(pos, p, 2)
    got, _ = eng.context_combine([0, 50], pos, p, 2, [1] * 5)
    assert want[22] == TINY and got[0, 22] == TINY
    assert_close_log10(got[0], want)


def test_reference_errors_become_status_codes(engine_factory):
    eng = engine_factory(tests=('KS',))
    rng = np.random.default_rng(2)
    # transcript 0: too short for context 3 (< 3c + 3 positions between first and last);
    # transcript 1: a p-value of 0; transcript 2: a p-value above 1; transcript 3: fine; 4: empty
    pos = np.concatenate([np.arange(11), np.arange(40), np.arange(40), np.arange(40)])
    p = rng.uniform(0.01, 1.0, pos.size)
    p[11 + 5] = 0.0
    p[51 + 7] = 1.5
    offsets = [0, 11, 51, 91, 131, 131]
    got, status = eng.context_combine(offsets, pos, p, 3, [1] * 7)
    assert status[0].tolist() == [L.NCB_CTX_TOO_SHORT, L.NCB_CTX_INVALID_PVALUE, L.NCB_CTX_INVALID_PVALUE,
                                  L.NCB_CTX_OK, L.NCB_CTX_OK]
    assert np.isnan(got[0, :91]).all()
    assert_close_log10(got[0, 91:], combine_context_pvalues(pos[91:], p[91:], 3))
    with pytest.raises(RuntimeError):
        cross_corr_matrix(np.ones(11), 3)
    # empty inputs
    got, status = eng.context_combine([0], np.zeros(0, np.int64), np.zeros((2, 0)), 1, [1] * 3)
    assert got.shape == (2, 0) and status.shape == (2, 0)
    with pytest.raises(Exception):
        eng.context_combine([0, 4], np.arange(4), np.ones(4), L.NCB_CTX_MAX + 1, [1] * (2 * L.NCB_CTX_MAX + 3))


def test_constant_track_gives_nan_like_numpy(engine_factory):
    # every shifted window has zero variance: np.corrcoef is NaN, and so is every combination
    # that is not a window of ones
    eng = engine_factory(tests=('KS',))
    pos = np.arange(30)
    p = np.full(30, 0.25)
    with np.errstate(all='ignore'):
        want = combine_context_pvalues(pos, p, 1)
    got, status = eng.context_combine([0, 30], pos, p, 1, [1] * 3)
    assert status[0, 0] == L.NCB_CTX_OK
    assert np.isnan(want).all() and np.isnan(got).all()


def test_hou_reduces_to_fisher_for_independent_pvalues(engine_factory):
    # the docstring's own check (comparisons.py:667-669): equal weights and zero correlation give
    # Fisher's method -- and the v1 known answers of tests/test_comparisons.py:189-206 go through
    # the oracle, which is pinned against them
    from scipy.stats import combine_pvalues
    got = combine_pvalues_hou([0.1, 0.02, 0.1, 0.02, 0.3], [1] * 5, np.zeros((5, 5)))
    assert abs(got - combine_pvalues([0.1, 0.02, 0.1, 0.02, 0.3], method='fisher')[1]) < 1e-12
