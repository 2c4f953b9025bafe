"""
Worker-side data preparation -- numpy restatement of
/root/reference/nanocompore/run.py:
  * ``_prepare_data``               532-601 (log10 dwell, motor column, drop empty positions)
  * ``_filter_low_cov_positions``   511-515
  * ``_downsample``                 518-529

Test infrastructure only (see oracle/__init__.py).

``log10``: numpy's float32 log10 is not specified to be correctly rounded; the
shared convention (oracle and kernel) is float64 log10 rounded once to float32.
"""
import numpy as np

F32 = np.float32
INTENSITY_POS, DWELL_POS, MOTOR_DWELL_POS = 0, 1, 2


def log10_dwell32(dwell):
    dwell = np.asarray(dwell, dtype=F32)
    x = (dwell + F32(1e-10)).astype(F32)                  # run.py:573, float32 add
    with np.errstate(invalid='ignore', divide='ignore'):
        return np.log10(x.astype(np.float64)).astype(F32)


def prepare_data(data, sample_ids, condition_ids, motor_offset=0):
    """run.py:532-601.  data (P,R,2) float32 raw (dwell in seconds).  Not in place."""
    data = np.array(data, dtype=F32, copy=True)
    data[:, :, DWELL_POS] = log10_dwell32(data[:, :, DWELL_POS])
    P = data.shape[0]
    if motor_offset > 0:
        data = np.pad(data, ((0, 0), (0, 0), (0, 1)), mode='constant', constant_values=np.nan)
        end = P - motor_offset
        if end > 0:
            data[:end, :, MOTOR_DWELL_POS] = data[motor_offset:, :, DWELL_POS]
    valid = np.any(~np.isnan(data[:, :, INTENSITY_POS]), axis=1)
    positions = np.arange(P, dtype=np.int16)[valid]
    return (data[valid], np.asarray(sample_ids, dtype=np.int16),
            np.asarray(condition_ids, dtype=np.int16), positions)


def filter_low_cov_positions(data, positions, conditions, min_cov):
    """run.py:511-515."""
    conditions = np.asarray(conditions)
    c0 = (~np.isnan(data[:, conditions == 0, INTENSITY_POS])).sum(1)
    c1 = (~np.isnan(data[:, conditions == 1, INTENSITY_POS])).sum(1)
    ok = (c0 >= min_cov) & (c1 >= min_cov)
    return data[ok], positions[ok]


def downsample(data, samples, conditions, max_reads):
    """run.py:518-529.  The reference's argsort(descending=True) is not stable, so the
    choice among equally-covered reads is implementation-defined; the shared definition
    is a STABLE descending sort (ties keep read order)."""
    samples = np.asarray(samples)
    conditions = np.asarray(conditions)
    if data.shape[1] > max_reads:
        nvalid = (~np.isnan(data).any(2)).sum(0)
        order = np.argsort(-nvalid, kind='stable')
        selected = np.zeros(order.shape[0], dtype=bool)
        for c in (0, 1):
            sel = order[conditions[order] == c][:max_reads]
            selected[sel] = True
        data = data[:, selected, :]
        samples = samples[selected]
        conditions = conditions[selected]
    return data, samples, conditions
