"""
The slice of the reference's ``Config`` (nanocompore/config.py:217-493) that the comparison path
reads, over the same YAML keys and with the same defaults (config.py:196-214).  The reference's
own ``Config`` object works unchanged with ``nanocompore_b200.comparisons.TranscriptComparator``
(the comparator only calls the getters below); this class exists so that the path can be driven
without the reference installed (its ``Config`` needs the ``schema`` package).
"""

DEFAULT_MIN_COVERAGE = 30
DEFAULT_DOWNSAMPLE_HIGH_COVERAGE = 5000
DEFAULT_COMPARISON_METHODS = ['GMM', 'KS']
DEFAULT_KIT = 'RNA002'
DEFAULT_MAX_INVALID_KMERS_FREQ = 0.1
DEFAULT_CORRECTION_METHOD = 'fdr_bh'
HARD_ASSIGNMENT = 'hard'
SOFT_ASSIGNMENT = 'soft'

_REMAP = {'GAUSSIAN_MIXTURE_MODEL': 'GMM', 'KOLMOGOROV_SMIRNOV': 'KS', 'T_TEST': 'TT',
          'MANN_WHITNEY': 'MW'}
_VALID = {'GMM', 'GOF', 'KS', 'TT', 'MW', 'auto', 'AUTO', 'GAUSSIAN_MIXTURE_MODEL',
          'GOODNESS_OF_FIT', 'KOLMOGOROV_SMIRNOV', 'T_TEST', 'MANN_WHITNEY'}


class PathConfig:
    def __init__(self, config: dict):
        if 'data' not in config or len(config['data']) != 2:
            raise ValueError("'data' must define exactly two conditions")
        if config.get('depleted_condition') not in config['data']:
            raise ValueError("The condition set in 'depleted_condition' is not defined in 'data'.")
        methods = config.get('comparison_methods', DEFAULT_COMPARISON_METHODS)
        if len(methods) == 0 or any(m not in _VALID for m in methods):
            raise ValueError(f"invalid comparison_methods {methods}")
        ctx = config.get('sequence_context', 0)
        if not (isinstance(ctx, int) and 0 <= ctx <= 4):
            raise ValueError('sequence_context must be >= 0 and <= 4')
        if config.get('motor_dwell_offset', 0) < 0:
            raise ValueError('motor_dwell_offset must be >= 0')
        # the reference's schema accepts exactly one value (config.py:175): refused here, when the
        # configuration is read, not when the correction runs at the end of a job
        if config.get('correction_method', DEFAULT_CORRECTION_METHOD) != DEFAULT_CORRECTION_METHOD:
            raise ValueError(f"correction_method must be {DEFAULT_CORRECTION_METHOD!r} "
                             f"(got {config.get('correction_method')!r})")
        self._config = config

    def get_data(self):
        return self._config['data']

    def get_comparison_methods(self):
        return [_REMAP.get(m, m) for m in self._config.get('comparison_methods', DEFAULT_COMPARISON_METHODS)]

    def get_motor_dwell_offset(self):
        return self._config.get('motor_dwell_offset', 0)

    def get_sequence_context(self):
        return self._config.get('sequence_context', 0)

    def get_sequence_context_weights(self):
        return self._config.get('sequence_context_weights', 'uniform')

    def get_correction_method(self):
        return self._config.get('correction_method', DEFAULT_CORRECTION_METHOD)

    def get_cluster_counts(self):
        return self._config.get('cluster_counts', HARD_ASSIGNMENT)

    def get_min_coverage(self):
        return self._config.get('min_coverage', DEFAULT_MIN_COVERAGE)

    def get_downsample_high_coverage(self):
        return self._config.get('downsample_high_coverage', DEFAULT_DOWNSAMPLE_HIGH_COVERAGE)

    def get_sample_labels(self):
        return [sample for samples in self.get_data().values() for sample in samples]

    def get_condition_labels(self):
        return list(self.get_data().keys())

    def get_depleted_condition(self):
        return self._config['depleted_condition']

    def get_sample_ids(self):
        labels = self.get_sample_labels()
        return dict(zip(labels, range(len(labels))))

    def get_condition_ids(self):
        labels = self.get_condition_labels()
        return dict(zip(labels, range(len(labels))))

    def sample_to_condition(self):
        return {sample: cond for cond, samples in self.get_data().items() for sample in samples}

    # ---- keys read by the transcript readers (run.py:657-855) --------------------------------
    def get_kit(self):
        """config.py:244-250; returns the kit NAME ('RNA002' / 'RNA004'), which is also what
        ``readers.kit_geometry`` accepts from the reference's ``Kit`` enum (``kit.name``)."""
        return self._config.get('kit', DEFAULT_KIT)

    def get_max_invalid_kmers_freq(self):
        return self._config.get('max_invalid_kmers_freq', DEFAULT_MAX_INVALID_KMERS_FREQ)

    def _sample_files(self, key):
        """(sample, condition, path) triples in the order of the ``data`` section."""
        out = []
        for condition, replicates in self.get_data().items():
            out.extend((sample, condition, entry[key]) for sample, entry in replicates.items())
        return out

    def get_sample_condition_bam_data(self):       # config.py: same name, same order
        return self._sample_files('bam')

    def get_sample_condition_db_data(self):
        return self._sample_files('db')

    def is_multi_replicate(self):
        return all(len(reps) > 1 for reps in self.get_data().values())
