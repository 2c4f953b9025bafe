"""
Whole per-transcript comparison -- restatement of
/root/reference/nanocompore/comparisons.py:41-215 (``compare_transcript``,
``_get_test_masks``, ``_auto_test_mask``, ``_run_test``, ``_merge_results``) and
285-335 / 395-403 (``_gmm_test`` motor-dwell split).

Test infrastructure only (see oracle/__init__.py).  It is also the "port" CPU
baseline timed by bench.py: like the reference it loops over positions in Python
and calls SciPy per position.
"""
import numpy as np
import pandas as pd

from oracle import stats as ostats
from oracle.nonparametric import nonparametric_test
from oracle.gmm_test import gmm_test_split
from oracle.gmm import GMM
from oracle.context import combine_context_pvalues

INTENSITY, DWELL, MOTOR = 0, 1, 2


class OracleConfig:
    """The config getters read on the path (config.py:323-330, 368-459), dict backed."""

    def __init__(self, comparison_methods=('GMM', 'KS'), sample_ids=None,
                 depleted_condition_id=0, motor_dwell_offset=0, sequence_context=0,
                 sequence_context_weights='uniform', cluster_counts='hard',
                 with_logit=False):
        self.comparison_methods = list(comparison_methods)
        self.sample_ids = sample_ids or {'kd1': 0, 'kd2': 1, 'wt1': 2, 'wt2': 3}
        self.depleted_condition_id = depleted_condition_id
        self.motor_dwell_offset = motor_dwell_offset
        self.sequence_context = sequence_context
        self.sequence_context_weights = sequence_context_weights
        self.cluster_counts = cluster_counts
        self.with_logit = with_logit


class OracleComparator:
    def __init__(self, config: OracleConfig, log=None, gmm_cls=GMM):
        self.cfg = config
        self.log = log or (lambda level, msg: None)
        self.gmm_cls = gmm_cls

    def compare_transcript(self, transcript_id, data, samples, conditions, positions):
        """Returns a DataFrame with the reference's columns, or None (comparisons.py:79-81)."""
        cfg = self.cfg
        data = np.asarray(data, dtype=np.float32)
        samples = np.asarray(samples)
        conditions = np.asarray(conditions)
        positions = np.asarray(positions)
        if len(positions) == 0:
            return None
        results = pd.DataFrame({'transcript_id': transcript_id, 'pos': positions})
        for k, v in ostats.shift_stats(data, conditions).items():
            results[k] = v

        std_data, _, bad, _ = ostats.outlier_standardise(data)
        if bad.any():
            self.log("warning", f"positions {positions[bad].tolist()} will be skipped")
            std_data = std_data[~bad]
            positions = positions[~bad]
            results = results.loc[~bad]
        n_positions = len(positions)

        methods = cfg.comparison_methods
        has_auto = 'auto' in methods
        auto_mask = None
        if has_auto:
            cov = (~np.isnan(std_data[:, :, INTENSITY])).sum(1)
            auto_mask = np.where(cov < 500, 'KS', 'GMM')          # comparisons.py:165-167
        base = [m for m in dict.fromkeys(methods) if m != 'auto']
        masks = {t: np.ones(n_positions, dtype=bool) for t in base}
        if auto_mask is not None:
            for t in sorted(set(auto_mask) - set(base)):
                masks[t] = auto_mask == t

        results = results.copy()
        for test, mask in masks.items():
            tr = self._run_test(test, std_data[mask], samples, conditions)
            if has_auto and 'auto_pvalue' not in results:
                results['auto_pvalue'] = np.full(n_positions, np.nan)
                results['auto_test'] = auto_mask
            for column, values in tr.items():
                results.loc[mask, column] = values
                if has_auto and column == {'KS': 'KS_intensity_pvalue',
                                           'GMM': 'GMM_chi2_pvalue'}.get(test):
                    rows = auto_mask == test
                    results.loc[rows, 'auto_pvalue'] = results.loc[rows, column]

        ctx = cfg.sequence_context
        if ctx > 0:
            for col in [c for c in results.columns if 'pvalue' in c]:
                results[f"{col}_context_{ctx}"] = combine_context_pvalues(
                    results['pos'].values, results[col].values, ctx,
                    cfg.sequence_context_weights)
        return results

    def _run_test(self, test, test_data, samples, conditions):
        if test.upper() in ('MW', 'KS', 'TT'):
            return nonparametric_test(test, test_data, conditions, self.log)
        if test.upper() == 'GMM':
            return self._gmm_test(test_data, samples, conditions)
        raise NotImplementedError(f"Test {test} is not implemented!")

    def _split(self, d, samples, conditions):
        cfg = self.cfg
        return gmm_test_split(d, samples, conditions, cfg.sample_ids,
                              cfg.depleted_condition_id, cfg.cluster_counts,
                              gmm_cls=self.gmm_cls, with_logit=cfg.with_logit)

    def _gmm_test(self, test_data, samples, conditions):
        # comparisons.py:285-335
        if self.cfg.motor_dwell_offset == 0:
            return self._split(test_data, samples, conditions)
        valid = ~np.isnan(test_data)
        with_motor = (valid[:, :, INTENSITY] & valid[:, :, MOTOR]).sum(1)
        total = valid[:, :, INTENSITY].sum(1)
        use_motor = with_motor >= 0.9 * total
        out = {}
        r3 = self._split(test_data[use_motor], samples, conditions) if use_motor.any() else {}
        r2 = self._split(test_data[~use_motor][:, :, :MOTOR], samples, conditions) \
            if (~use_motor).any() else {}
        for col in set(r3) | set(r2):
            proto = r3[col] if col in r3 else r2[col]
            merged = np.empty(len(use_motor), dtype=np.asarray(proto).dtype)
            if col in r3:
                merged[use_motor] = r3[col]
            if col in r2:
                merged[~use_motor] = r2[col]
            out[col] = merged
        return out
