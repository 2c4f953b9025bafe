"""
Shift statistics, outlier filter and standardisation -- numpy restatement.

Follows /root/reference/nanocompore/comparisons.py:
  * ``nanstd``                     622-624
  * ``_add_shift_stats``           406-433
  * outlier filter + standardise   95-120

Test infrastructure only (see oracle/__init__.py).

float32 convention: torch's ``nanmean``/``nansum`` reduce in float32 in an
implementation-defined order.  Here (and in the CUDA kernel) a reduction is the
exact sum (float64 accumulation) rounded once to float32; every other float32
operation (subtract, multiply, divide, sqrt) is performed in float32 exactly where
the reference performs it.
"""
import numpy as np

F32 = np.float32


def nansum32(x, axis):
    """float32 nansum with an exactly-rounded result (float64 accumulation)."""
    return np.nansum(x.astype(np.float64), axis=axis).astype(F32)


def nancount(x, axis):
    return (~np.isnan(x)).sum(axis=axis)


def nanmean32(x, axis):
    """torch.Tensor.nanmean: nansum / count, both float32 (comparisons.py:96,101,416)."""
    n = nancount(x, axis).astype(F32)
    with np.errstate(invalid='ignore', divide='ignore'):
        return (nansum32(x, axis) / n).astype(F32)


def nanstd32(x, axis):
    """
    comparisons.py:622-624
        nonnans = (~X.isnan()).sum(dim)
        (((X - X.nanmean(dim).unsqueeze(1)) ** 2).nansum(dim) / nonnans) ** 0.5
    population SD, float32 op by op.
    """
    assert x.dtype == F32
    n = nancount(x, axis).astype(F32)
    mean = np.expand_dims(nanmean32(x, axis), axis)
    d = (x - mean).astype(F32)
    sq = (d * d).astype(F32)
    with np.errstate(invalid='ignore', divide='ignore'):
        var = (nansum32(sq, axis) / n).astype(F32)
        return np.sqrt(var).astype(F32)


def lower_nanmedian(x, axis=1):
    """torch.nanmedian returns the LOWER of the two middle values (comparisons.py:417,422)."""
    assert axis == 1
    P = x.shape[0]
    srt = np.sort(x, axis=1)                 # NaNs sort to the end
    n = nancount(x, 1)                       # (P, V)
    idx = np.maximum(n - 1, 0) // 2          # lower middle
    out = np.take_along_axis(srt, idx[:, None, :], axis=1)[:, 0, :]
    out = out.astype(F32)
    out[n == 0] = np.nan
    return out


def shift_stats(data, conditions):
    """
    comparisons.py:406-433.  ``data`` (P, R, V) float32 raw (pre-filter) values,
    ``conditions`` (R,).  Returns dict of column name -> float32 (P,) in the
    reference's insertion order.
    """
    data = np.asarray(data, dtype=F32)
    cond0 = (np.asarray(conditions) == 0)[None, :, None]
    out = {}
    stats = {}
    for c, mask in ((0, cond0), (1, ~cond0)):
        cd = np.where(mask, data, F32(np.nan)).astype(F32)
        stats[c] = {'mean': nanmean32(cd, 1),
                    'median': lower_nanmedian(cd, 1),
                    'sd': nanstd32(cd, 1)}
    dims = {'intensity': 0, 'dwell': 1}
    if data.shape[2] > 2:
        dims['motor_dwell'] = 2
    for c in (0, 1):
        for stat in ('mean', 'median', 'sd'):
            for dim, di in dims.items():
                out[f"c{c+1}_{stat}_{dim}"] = stats[c][stat][:, di]
    return out


def outlier_standardise(data):
    """
    comparisons.py:95-120.  Returns (standardised data (P,R,V) float32,
    outlier mask (P,R), bad_std mask (P,), (mean1,std1,mean2,std2)).
    The input is not modified.
    """
    data = np.array(data, dtype=F32, copy=True)
    std1 = nanstd32(data, 1)
    mean1 = nanmean32(data, 1)
    with np.errstate(invalid='ignore', divide='ignore'):
        z = ((data - mean1[:, None, :]).astype(F32) / std1[:, None, :]).astype(F32)
        outliers = (np.abs(z) > F32(3)).any(2)
    data[outliers] = np.nan
    std2 = nanstd32(data, 1)
    mean2 = nanmean32(data, 1)
    with np.errstate(invalid='ignore', divide='ignore'):
        std_data = ((data - mean2[:, None, :]).astype(F32) / std2[:, None, :]).astype(F32)
    bad = np.isnan(std2).any(1) | (std2 == 0).any(1)
    return std_data, outliers, bad, (mean1, std1, mean2, std2)
