"""
Mirror of the three ``Worker`` data-preparation steps that feed ``compare_transcript``
(/root/reference/nanocompore/run.py:511-601), same names, arguments and results:

  * ``_prepare_data``               run.py:532-601  log10(dwell + 1e-10), optional motor-dwell column,
                                                    drop positions without any read, tensors on the device
  * ``_filter_low_cov_positions``   run.py:511-515  keep positions with >= min_cov reads in BOTH conditions
  * ``_downsample``                 run.py:518-529  per condition keep the max_reads most complete reads

They are thin tensor utilities (torch is the plumbing here).  On the batch path the first two
are not needed at all: libncb takes raw dwell times (``dwell_is_raw``) and applies the coverage
filter inside the statistics kernel (``min_coverage``), see ``Engine``.

``WorkerPrepMixin`` gives a reference ``Worker`` subclass the three methods under the
reference's names (they read ``self._conf`` and ``self._device`` like the originals).
"""
import numpy as np
import torch

INTENSITY_POS, DWELL_POS, MOTOR_DWELL_POS = 0, 1, 2


def prepare_data(data, sample_ids, condition_ids, motor_dwell_offset=0, device='cpu'):
    """run.py:532-601.  ``data`` (P, R, 2) float32 numpy with NaN, dwell in seconds; modified in
    place like the reference does.  Returns (tensor (P', R, V), samples, conditions, positions).

    The logarithm is the correctly rounded float32 log10 (fp64 log10, rounded once): the convention
    shared with the kernels and the oracle; libm's float32 log10 differs by at most 2 ulps."""
    dwell = data[:, :, DWELL_POS] + np.float32(1e-10)
    with np.errstate(invalid='ignore', divide='ignore'):
        data[:, :, DWELL_POS] = np.log10(dwell.astype(np.float64)).astype(np.float32)
    n_positions = data.shape[0]
    if motor_dwell_offset > 0:
        data = np.pad(data, ((0, 0), (0, 0), (0, 1)), mode='constant', constant_values=np.nan)
        end = n_positions - motor_dwell_offset
        if end > 0:
            data[:end, :, MOTOR_DWELL_POS] = data[motor_dwell_offset:, :, DWELL_POS]
    valid_positions = np.any(~np.isnan(data[:, :, INTENSITY_POS]), axis=1)
    tensor = torch.tensor(data[valid_positions], dtype=torch.float32, device=device)
    samples = torch.tensor(np.asarray(sample_ids), dtype=torch.int16, device=device)
    conditions = torch.tensor(np.asarray(condition_ids), dtype=torch.int16, device=device)
    positions = torch.arange(n_positions, dtype=torch.int16, device=device)[torch.from_numpy(valid_positions).to(device)]
    return tensor, samples, conditions, positions


def filter_low_cov_positions(data, positions, conditions, min_cov):
    """run.py:511-515."""
    valid = ~torch.isnan(data[:, :, INTENSITY_POS])
    cov0 = valid[:, conditions == 0].sum(1)
    cov1 = valid[:, conditions == 1].sum(1)
    keep = (cov0 >= min_cov) & (cov1 >= min_cov)
    return data[keep], positions[keep]


def downsample(data, samples, conditions, max_reads):
    """run.py:518-529.  Reads are ranked by the number of positions where all their variables are
    valid; per condition the first ``max_reads`` are kept.  The reference's ``argsort(descending=
    True)`` is not stable, so its choice among equally complete reads is implementation-defined;
    here the sort is stable (equally complete reads keep their read order), like oracle/prep.py."""
    if data.shape[1] > max_reads:
        complete = (~torch.isnan(data).any(2)).sum(0)
        order = torch.argsort(complete, descending=True, stable=True)
        selected = torch.zeros(order.shape[0], dtype=torch.bool, device=data.device)
        for cond in (0, 1):
            chosen = order[conditions[order] == cond][:max_reads]
            selected[chosen] = True
        data = data[:, selected, :]
        samples = samples[selected]
        conditions = conditions[selected]
    return data, samples, conditions


class WorkerPrepMixin:
    """``class Worker(WorkerPrepMixin, nanocompore.run.Worker)`` swaps the three steps in."""

    def _prepare_data(self, data, sample_ids, condition_ids):
        return prepare_data(data, sample_ids, condition_ids,
                            self._conf.get_motor_dwell_offset(), self._device)

    def _filter_low_cov_positions(self, data, positions, conditions, min_cov):
        return filter_low_cov_positions(data, positions, conditions, min_cov)

    def _downsample(self, data, samples, conditions, max_reads):
        return downsample(data, samples, conditions, max_reads)
