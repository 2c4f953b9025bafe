"""
Mirror of the reference's ``nanocompore/comparisons.py`` for the per-position two-condition
comparison, running on the B200 kernels of libncb.

Same names, signatures, config getters, result columns and error behaviour as the reference
(/root/reference/nanocompore/comparisons.py).  What "same results" covers and what it does not:
the shift statistics, the outlier filter, KS / MW / TT, the contingency arithmetic (LOR,
chi-squared, counts from a given partition) and the BH correction are pinned on outputs of the
unmodified reference; the MIXTURE FIT itself (``GMM_chi2_pvalue``, ``GMM_LOR``, the BIC gate and
the ``*_mod`` / ``*_unmod`` counts) is NOT verified against the reference's fitter: gmm-gpu 0.2.8
is neither vendored with the reference nor installable here.  The kernels implement this
repository's own definition (oracle/gmm.py: scikit-learn's full-covariance EM, stop rule and BIC,
started from a deterministic farthest-pair 2-means -- ``random_seed`` is accepted and ignored),
so positions whose fit depends on the initialisation can get another p-value or another side of
the BIC gate than upstream.  tests/test_gmm_gpu_parity.py makes the comparison wherever
``gmm_gpu`` is importable; until it has run, the mixture columns are "definition-pinned", not
drop-in.

  Interface mirrored:

  * ``TranscriptComparator(config, worker, random_seed=42, dtype=torch.float32)``   :29-38
  * ``TranscriptComparator.compare_transcript(transcript, data, samples, conditions,
    positions, device) -> (transcript, DataFrame | None)``                          :41-148
  * module helpers ``nanstd``, ``calculate_lors``, ``get_contingency_matrices``,
    ``get_soft_contingency_matrices``, ``cross_corr_matrix``, ``harmomic_series``,
    ``combine_pvalues_hou``                                                         :622-744

What differs underneath: the reference walks the positions in Python and calls torch / SciPy /
gmm_gpu per position; here the whole tensor goes through ONE libncb call (shift statistics,
outlier filter, standardisation, KS / MW / TT, both Gaussian mixtures, contingency, counts,
LOR, chi-squared) and this module only assembles the DataFrame.  There is no CPU fallback:
with no GPU or no libncb.so the constructor's first use raises.

``motor_dwell_offset > 0`` (a third variable, comparisons.py:285-335) is supported: positions
where at least 90 % of the reads carry a motor dwell are fitted in 3 variables, the others in 2,
inside the same libncb call.  Not supported (raises NotImplementedError): the ``GOF`` test
(out of scope).
"""
import numpy as np
import pandas as pd
import torch

from nanocompore_b200 import _lib as L
from nanocompore_b200.engine import Engine, Results, make_labels, tests_mask

INTENSITY = 0
DWELL = 1
MOTOR = 2

HARD_ASSIGNMENT = 'hard'
SOFT_ASSIGNMENT = 'soft'

def shift_columns(with_motor):
    """Column name -> stats row, in the reference's insertion order (comparisons.py:421-433)."""
    dims = [('intensity', 'INTENSITY'), ('dwell', 'DWELL')] + ([('motor_dwell', 'MOTOR')] if with_motor else [])
    return [(f"c{c + 1}_{stat}_{dim}", L.ST[f"C{c + 1}_{stat.upper()}_{key}"])
            for c in (0, 1) for stat in ('mean', 'median', 'sd') for dim, key in dims]


SHIFT_COLUMNS = shift_columns(False)

_NONPARAMETRIC = {'KS': ('KS_INTENSITY', 'KS_DWELL'), 'MW': ('MW_INTENSITY', 'MW_DWELL'),
                  'TT': ('TT_INTENSITY', 'TT_DWELL')}
_AUTO_PVALUE = {'KS': 'KS_intensity_pvalue', 'GMM': 'GMM_chi2_pvalue'}


class NanocomporeError(Exception):
    """Same role as nanocompore.common.NanocomporeError."""


class TranscriptComparator:
    """
    Compare a set of positions (reference: comparisons.py:29-619).
    """

    def __init__(self, config, worker, random_seed=42, dtype=torch.float32, with_logit=False):
        self._config = config
        # The reference seeds gmm_gpu's random initialisation; this implementation initialises
        # the mixtures deterministically (oracle/gmm.py), so the seed is accepted and unused.
        self._random_seed = random_seed
        self._dtype = dtype
        self._motor_dwell_offset = self._config.get_motor_dwell_offset()
        self._worker = worker
        self._with_logit = with_logit
        self._engines = {}
        self._context_engine = None     # the engine of the last comparison: it also combines contexts

    # ------------------------------------------------------------------ engine
    def _engine(self, device):
        """One libncb handle per CUDA device, created on first use (no CUDA work at import or
        construction time, so fork and spawn start methods both work: run.py:364-372)."""
        if isinstance(device, torch.device):
            device = str(device)
        if not (isinstance(device, str) and device.startswith('cuda')):
            # the reference's default is `devices: {'cpu': 2}` (config.py:196-214): there is no CPU
            # path here, and running such a worker silently on GPU 0 would oversubscribe it
            raise NanocomporeError(
                f"device {device!r}: nanocompore_b200 runs the comparison on CUDA devices only; "
                "set `devices: {'cuda:0': <workers>, ...}` in the configuration")
        dev = torch.device(device)
        index = dev.index if dev.index is not None else 0
        if index not in self._engines:
            methods = [m for m in self._config.get_comparison_methods()]
            for m in methods:
                if m.upper() == 'GOF':
                    raise NotImplementedError("Test GOF is not implemented!")
            self._engines[index] = Engine(
                index,
                tests=tests_mask(methods, self._with_logit),
                n_samples=max(self._config.get_sample_ids().values()) + 1,
                depleted_condition=self._depleted_condition_id(),
                min_coverage=0,            # Worker._filter_low_cov_positions already ran
                dwell_is_raw=False,        # Worker._prepare_data already took the log10
                cluster_counts=self._config.get_cluster_counts())
        self._context_engine = self._engines[index]
        return self._engines[index]

    def _depleted_condition_id(self):
        depleted = self._config.get_depleted_condition()
        return int(self._config.get_condition_ids()[depleted])

    # ------------------------------------------------------------------ the hot path
    def compare_transcript(self, transcript, data, samples, conditions, positions, device):
        """
        Compare the two conditions for the given transcript (comparisons.py:41-148).

        data: (positions, reads, vars) float tensor, NaN = missing; samples, conditions: (reads,);
        positions: (positions,).  Returns (transcript, DataFrame) or (transcript, None) when
        there are no positions.
        """
        n_positions = len(positions)
        if n_positions == 0:
            return (transcript, None)
        with_motor = self._motor_dwell_offset > 0
        if data.shape[2] != (3 if with_motor else 2):
            raise NanocomporeError(
                f"data has {data.shape[2]} variables, motor_dwell_offset = {self._motor_dwell_offset}")

        engine = self._engine(device)
        self._worker.log("trace", "Start comparison kernels")
        dev_data = torch.as_tensor(data).to(device=engine.device, dtype=torch.float32).contiguous()
        labels = torch.from_numpy(make_labels(_to_numpy(samples), _to_numpy(conditions))).to(engine.device)
        res = engine.compare_dense(dev_data, labels).numpy()
        return transcript, self._assemble(transcript, positions, res['status'], res['pvalues'],
                                          res['stats'], res['counts'], with_motor)

    # ------------------------------------------------------------------ the reference's stages
    # compare_transcript above runs every stage in one libncb call.  The methods below are the
    # stages under the names the reference gives them (its own tests call them directly:
    # tests/test_comparisons.py:21-154, 307-504); each is one libncb call restricted to that stage.
    def _stage(self, data, labels, device, **opts):
        """One ``ncb_compare`` on ``data`` (positions, reads, vars) with the engine's options
        replaced by ``opts`` for the call."""
        engine = self._engine(device)
        keep = dict(tests=engine.opts.tests, input_standardised=bool(engine.opts.input_standardised))
        engine.configure(**opts)
        try:
            dev_data = torch.as_tensor(data).to(device=engine.device, dtype=torch.float32).contiguous()
            return engine.compare_dense(dev_data, torch.from_numpy(labels).to(engine.device)).numpy()
        finally:
            engine.configure(**keep)

    def _stage_device(self, device=None):
        if isinstance(device, str) and device.startswith('cuda'):
            return device
        if self._engines:
            return f'cuda:{next(iter(self._engines))}'
        return 'cuda:0'

    def _add_shift_stats(self, results, data, conditions, device):
        """comparisons.py:406-433: mean, lower median and population sd of every variable per
        condition, on the values as given."""
        cond = _to_numpy(conditions)
        labels = make_labels(np.zeros_like(cond), cond)
        res = self._stage(data, labels, device, tests=0, input_standardised=True)
        for column, row in shift_columns(self._motor_dwell_offset > 0):
            results[column] = torch.from_numpy(res['stats'][row].copy())

    def _nonparametric_test(self, test, test_data, conditions):
        """comparisons.py:218-252 on filtered, standardised data (vars beyond intensity and dwell
        are ignored, as there)."""
        names = {'mann_whitney': 'MW', 'MW': 'MW', 'kolmogorov_smirnov': 'KS', 'KS': 'KS',
                 't_test': 'TT', 'TT': 'TT'}
        if test not in names:
            raise NanocomporeError("Invalid statistical method name (MW, KS, TT)")
        key = names[test]
        cond = _to_numpy(conditions)
        labels = make_labels(np.zeros_like(cond), cond)
        data = torch.as_tensor(test_data)[:, :, :2]
        res = self._stage(data, labels, self._stage_device(), tests=tests_mask([key]), input_standardised=True)
        a, b = _NONPARAMETRIC[key]
        return {f'{test}_intensity_pvalue': list(res['pvalues'][L.PV[a]]),
                f'{test}_dwell_pvalue': list(res['pvalues'][L.PV[b]])}

    def _gmm_test_split(self, data, samples, conditions, device):
        """comparisons.py:338-392 on filtered, standardised data of 2 or 3 variables: both mixture
        fits, BIC gate, contingency, counts, LOR, chi-squared."""
        labels = make_labels(_to_numpy(samples), _to_numpy(conditions))
        res = self._stage(data, labels, device, tests=tests_mask(['GMM'], self._with_logit),
                          input_standardised=True)
        n = res['status'].size
        return self._test_columns('GMM', res['pvalues'], res['stats'], res['counts'],
                                  np.zeros(n, dtype=np.int32), None)

    def _split_by_ndim(self, test_data):
        """comparisons.py:395-403: positions where at least 90 % of the reads with an intensity also
        carry a motor dwell are fitted in 3 variables, the others in 2."""
        valid = ~torch.isnan(test_data)
        with_motor = (valid[:, :, INTENSITY] & valid[:, :, MOTOR]).sum(1)
        use_motor = with_motor >= 0.9 * valid[:, :, INTENSITY].sum(1)
        return test_data[use_motor], test_data[~use_motor][:, :, :MOTOR], (~use_motor).int()

    def _gmm_test(self, test_data, samples, conditions, device):
        """comparisons.py:285-335: with a motor-dwell column the positions are split by
        ``_split_by_ndim``, each part goes through ``_gmm_test_split`` and the columns are merged back
        in position order.  (``compare_transcript`` does not come through here: the kernels make the
        same split per position inside one call.)"""
        if self._motor_dwell_offset == 0:
            return self._gmm_test_split(test_data, samples, conditions, device)
        dim3, dim2, split = self._split_by_ndim(test_data)
        parts = {}
        if dim3.shape[0] > 0:
            parts[0] = self._gmm_test_split(dim3, samples, conditions, device)
        if dim2.shape[0] > 0:
            parts[1] = self._gmm_test_split(dim2, samples, conditions, device)
        split = _to_numpy(split)
        merged = {}
        for column in {c for part in parts.values() for c in part}:
            cursor = {0: 0, 1: 0}
            values = []
            for which in split:
                values.append(parts[int(which)][column][cursor[int(which)]])
                cursor[int(which)] += 1
            merged[column] = values
        return merged

    def _get_mod_cluster(self, contingency):
        """comparisons.py:564-605: the cluster in which the depleted condition has the smaller share
        of its reads is the modified one (ties: cluster 0)."""
        contingency = torch.as_tensor(contingency)
        totals = contingency.sum(2).to(torch.uint32)
        share = contingency / totals.unsqueeze(2)
        return share[:, self._depleted_condition_id(), :].argmin(1)

    def _get_cluster_counts(self, contingency, samples, predictions):
        """comparisons.py:477-517: reads of every sample in the modified / the other cluster; a NaN
        prediction (read not in the fit) counts for neither."""
        predictions = torch.as_tensor(predictions)
        samples = torch.as_tensor(samples)
        mod = self._get_mod_cluster(contingency)[:, None]
        counts = {}
        for label, sid in self._config.get_sample_ids().items():
            pred = predictions[:, samples == sid]
            in_fit = (~torch.isnan(pred.to(torch.float64))).sum(1)
            n_mod = (pred == mod).sum(1)
            counts[f'{label}_mod'] = n_mod.cpu().numpy().astype(np.uint32)
            counts[f'{label}_unmod'] = (in_fit - n_mod).cpu().numpy().astype(np.uint32)
        return counts

    def _get_soft_cluster_counts(self, contingency, samples, cluster_probs):
        """comparisons.py:520-561: the posterior mass of every sample in the modified / the other
        cluster (NaN posteriors add nothing)."""
        cluster_probs = torch.as_tensor(cluster_probs)
        samples = torch.as_tensor(samples)
        B, N, _ = cluster_probs.shape
        mod = self._get_mod_cluster(contingency)[:, None, None].expand(B, N, 1)
        p_mod = torch.gather(cluster_probs, 2, mod).squeeze(2)
        p_other = torch.gather(cluster_probs, 2, 1 - mod).squeeze(2)
        counts = {}
        for label, sid in self._config.get_sample_ids().items():
            mask = samples == sid
            counts[f'{label}_mod'] = p_mod[:, mask].nansum(1).cpu().numpy()
            counts[f'{label}_unmod'] = p_other[:, mask].nansum(1).cpu().numpy()
        return counts

    def compare_transcripts(self, items, device):
        """
        Batch form of ``compare_transcript`` (an extension; the reference has no batch call):
        ``items`` is a list of ``(transcript, data, samples, conditions, positions)`` as
        ``compare_transcript`` takes them.  All transcripts are packed into ONE ragged batch by
        libncb's host packer and go through ONE ``ncb_compare`` call; the result is the list of
        ``(transcript, DataFrame | None)`` the per-transcript calls would give.
        """
        from nanocompore_b200.engine import pack_csr
        with_motor = self._motor_dwell_offset > 0
        live = [(i, it) for i, it in enumerate(items) if len(it[4]) > 0]
        out = [(it[0], None) for it in items]
        if not live:
            return out
        if with_motor:          # the packer writes 2 variables: 3-variable batches go one by one
            for i, (tr, data, samples, conditions, positions) in live:
                out[i] = self.compare_transcript(tr, data, samples, conditions, positions, device)
            return out
        engine = self._engine(device)
        batch, ptx, _ = pack_csr([(_to_numpy(d).astype(np.float32, copy=False), _to_numpy(s), _to_numpy(c))
                                  for _, (_, d, s, c, _) in live], pin=True)
        res = engine.compare_csr(batch).numpy()
        for k, (i, (tr, data, samples, conditions, positions)) in enumerate(live):
            sel = ptx == k
            out[i] = (tr, self._assemble(tr, positions, res['status'][sel], res['pvalues'][:, sel],
                                         res['stats'][:, sel], res['counts'][:, :, sel], with_motor))
        return out

    def compare_packed(self, batch, device, min_coverage=0, dwell_is_raw=True):
        """
        Entry of the reader pipeline (``readers.TranscriptPipeline``): one ragged CSR batch as
        ``readers.pack_rows`` builds it -- RAW dwell times, every position of every transcript --
        through one ``ncb_compare`` call.  The log10 of run.py:573 and the coverage filter of
        run.py:511-515 run inside the statistics kernel.  Returns the result columns as numpy
        arrays, one slot per position of the batch (``status`` carries the filter's verdict).
        """
        with_motor = self._motor_dwell_offset > 0
        if batch.signal.shape[1] != (3 if with_motor else 2):
            raise NanocomporeError(
                f"batch has {batch.signal.shape[1]} variables, motor_dwell_offset = {self._motor_dwell_offset}")
        engine = self._engine(device)
        engine.configure(min_coverage=min_coverage, dwell_is_raw=dwell_is_raw)
        try:
            return engine.compare_csr(batch).numpy()
        finally:
            engine.configure(min_coverage=0, dwell_is_raw=False)

    def assemble(self, transcript, positions, status, pv, st, counts):
        """Result frame of one transcript from its slice of the result columns of a batch."""
        if len(positions) == 0:
            return None
        return self._assemble(transcript, positions, status, pv, st, counts, self._motor_dwell_offset > 0)

    def assemble_batch(self, transcripts, starts, keeps, res, kmer_len=None):
        """Result frames of all transcripts of one batch: ``_assemble`` for every transcript, built as ONE
        frame over the batch's rows and cut into per-transcript slices (25 columns inserted into a fresh
        DataFrame cost ~1 ms per transcript, twice the kernels of the transcript; the batch frame costs
        ~0.1 ms per transcript).  ``starts`` [n + 1]: first batch position of every transcript; ``keeps``:
        per transcript the boolean mask of its tested positions; ``res``: the result columns of the batch
        (``Results.numpy()``).  ``kmer_len``: adds the ``kmer`` column of run.py:461-463 for transcripts
        with a sequence.  Returns ``[DataFrame | None]``; values, dtypes, column order, row labels, log
        messages and errors are those of the per-transcript path (tests/test_assemble_host.py)."""
        with_motor = self._motor_dwell_offset > 0
        methods = list(self._config.get_comparison_methods())
        seqs = [getattr(t, 'seq', None) for t in transcripts]
        with_kmer = kmer_len is not None and any(seqs)
        if 'auto' in methods or (with_kmer and not all(seqs)):
            # the incremental merge of the ``auto`` method (and a batch where only some transcripts
            # carry a sequence) goes transcript by transcript
            out = []
            for k, t in enumerate(transcripts):
                a, b = int(starts[k]), int(starts[k + 1])
                keep = np.asarray(keeps[k], dtype=bool)
                positions = np.nonzero(keep)[0]
                frame = None
                if positions.size:
                    frame = self.assemble(t, positions.astype(np.int16), res['status'][a:b][keep],
                                          res['pvalues'][:, a:b][:, keep], res['stats'][:, a:b][:, keep],
                                          res['counts'][:, :, a:b][:, :, keep])
                if frame is not None and kmer_len is not None and seqs[k]:
                    from nanocompore_b200.readers import encode_kmers
                    frame['kmer'] = encode_kmers(seqs[k], frame.pos.to_numpy(), kmer_len)
                out.append(frame)
            return out

        n_tx = len(transcripts)
        pos_list = [np.nonzero(np.asarray(k, dtype=bool))[0] for k in keeps]
        n_sel = np.array([p.size for p in pos_list], dtype=np.int64)
        if n_tx == 0 or n_sel.sum() == 0:
            return [None] * n_tx
        rows = np.concatenate([p + int(s) for p, s in zip(pos_list, starts[:n_tx])])
        pos = np.concatenate(pos_list).astype(np.int16)
        labels = np.concatenate([np.arange(p.size) for p in pos_list])   # row labels of the unfiltered frames
        seg = np.repeat(np.arange(n_tx), n_sel)
        status = res['status'][rows]
        too_many = (status & L.NCB_POS_TOO_MANY_READS) != 0
        if too_many.any():
            name = transcripts[int(seg[np.nonzero(too_many)[0][0]])].name
            raise NanocomporeError(
                f"A position of transcript {name} has more than {L.NCB_MAX_READS} reads; "
                "lower downsample_high_coverage")
        bad_stds = (status & L.NCB_POS_BAD_STD) != 0                     # comparisons.py:108-120
        if bad_stds.any():
            for k in np.unique(seg[bad_stds]):
                self._worker.log(
                    "warning",
                    "The standard deviation cannot be calculated for some positions on "
                    f"transcript {transcripts[int(k)].name}, but are required for scaling the data. "
                    f"The positions {pos[(seg == k) & bad_stds].tolist()} will be skipped.")
            good = ~bad_stds
            rows, pos, labels, seg, status = rows[good], pos[good], labels[good], seg[good], status[good]
        n_rows = np.bincount(seg, minlength=n_tx)
        ends = np.cumsum(n_rows)
        ids = [t.id for t in transcripts]
        columns = {'transcript_id': np.repeat(np.array(ids), n_rows) if rows.size else np.full(0, ids[0]),
                   'pos': pos}
        pv, st, counts = res['pvalues'][:, rows], res['stats'][:, rows], res['counts'][:, :, rows]
        for column, row in shift_columns(with_motor):
            columns[column] = st[row]
        skipped = (status & L.NCB_POS_MW_EXACT_SKIPPED) != 0
        if skipped.any() and any(m.upper() == 'MW' for m in methods):
            for k in np.unique(seg[skipped]):
                self._worker.log("error", "Error in nonparametric test: exact Mann-Whitney "
                                 f"distribution too large on transcript {transcripts[int(k)].name}")
        quiet = status & ~np.int32(L.NCB_POS_MW_EXACT_SKIPPED)
        for test in self._get_test_masks(None, rows.size):
            columns.update(self._test_columns(test, pv, st, counts, quiet, None))
        context = self._config.get_sequence_context()
        if context > 0:
            tests = [col for col in columns if 'pvalue' in col]
            live = n_rows > 0
            offsets = np.concatenate([[0], ends[live]])
            combined = self._combine_context(pos, [columns[test] for test in tests], offsets)
            for test, values in zip(tests, combined):
                columns[f"{test}_context_{context}"] = values
        if with_kmer:
            from nanocompore_b200.readers import encode_kmers
            columns['kmer'] = np.concatenate([encode_kmers(seqs[k], pos[ends[k] - n_rows[k]:ends[k]], kmer_len)
                                              for k in range(n_tx)])
        table = pd.DataFrame(columns, index=labels, copy=False)
        return [None if n_sel[k] == 0 else table.iloc[int(ends[k] - n_rows[k]):int(ends[k])]
                for k in range(n_tx)]

    def _assemble(self, transcript, positions, status, pv, st, counts, with_motor):
        """Result frame of one transcript from its slice of the result columns
        (comparisons.py:83-84, 108-148, 204-215).  Without the ``auto`` method every test covers
        every row, and the frame is built once from a dict of columns (inserting ~25 columns one
        by one costs more than the kernels of the transcript); with ``auto`` the reference's
        incremental merge is followed step by step."""
        pos_host = positions.cpu() if isinstance(positions, torch.Tensor) else torch.as_tensor(positions)
        if (status & L.NCB_POS_TOO_MANY_READS).any():
            raise NanocomporeError(
                f"A position of transcript {transcript.name} has more than {L.NCB_MAX_READS} reads; "
                "lower downsample_high_coverage")
        # comparisons.py:108-120
        bad_stds = (status & L.NCB_POS_BAD_STD) != 0
        if bad_stds.any():
            self._worker.log(
                "warning",
                "The standard deviation cannot be calculated for some positions on "
                f"transcript {transcript.name}, but are required for scaling the data. "
                f"The positions {pos_host[torch.from_numpy(bad_stds)].tolist()} will be skipped.")
        keep = ~bad_stds
        n_positions = int(keep.sum())
        methods = list(self._config.get_comparison_methods())
        context = self._config.get_sequence_context()

        if 'auto' not in methods:
            pos = pos_host.numpy()[keep]
            columns = {'transcript_id': np.full(n_positions, transcript.id), 'pos': pos}
            for column, row in shift_columns(with_motor):
                columns[column] = st[row][keep]
            status, pv, st, counts = status[keep], pv[:, keep], st[:, keep], counts[:, :, keep]
            for test in self._get_test_masks(None, n_positions):
                columns.update(self._test_columns(test, pv, st, counts, status, transcript))
            if context > 0:
                tests = [col for col in columns if 'pvalue' in col]
                combined = self._combine_context(pos, [columns[test] for test in tests])
                for test, values in zip(tests, combined):
                    columns[f"{test}_context_{context}"] = values
            # the index keeps the row labels of the unfiltered frame, like the reference's drop()
            return pd.DataFrame(columns, index=np.arange(keep.size)[keep])

        results = pd.DataFrame({'transcript_id': transcript.id, 'pos': pos_host})
        for column, row in shift_columns(with_motor):
            results[column] = st[row]
        if bad_stds.any():
            results.drop(np.arange(results.shape[0])[bad_stds], inplace=True, axis=0)
        status, pv, st, counts = status[keep], pv[:, keep], st[:, keep], counts[:, :, keep]
        auto_test_mask = np.where((status & L.NCB_POS_AUTO_GMM) != 0, 'GMM', 'KS')
        for test, mask in self._get_test_masks(auto_test_mask, n_positions).items():
            test_results = self._test_columns(test, pv, st, counts, status, transcript)
            self._merge_results(results, test_results, test, mask, auto_test_mask, n_positions)
        if context > 0:
            tests = [col for col in results.columns if 'pvalue' in col]
            combined = self._combine_context(results.pos.to_numpy(),
                                             [results[test].to_numpy(dtype=float) for test in tests])
            for test, values in zip(tests, combined):
                results[f"{test}_context_{context}"] = values
        return results

    def _get_test_masks(self, auto_test_mask, n_positions):
        # comparisons.py:150-162; iteration order is the config order (the reference iterates
        # a set, so its column order between tests is not defined)
        methods = list(dict.fromkeys(self._config.get_comparison_methods()))
        base_tests = [m for m in methods if m != 'auto']
        masks = {test: np.ones(n_positions, dtype=bool) for test in base_tests}
        if auto_test_mask is not None:
            for test in sorted(set(auto_test_mask) - set(base_tests)):
                masks[test] = auto_test_mask == test
        return masks

    def _test_columns(self, test, pv, st, counts, status, transcript):
        """Result columns of one test over all kept positions (what ``_run_test`` returns)."""
        key = test.upper()
        if key in _NONPARAMETRIC:
            a, b = _NONPARAMETRIC[key]
            if key == 'MW' and (status & L.NCB_POS_MW_EXACT_SKIPPED).any():
                self._worker.log("error", "Error in nonparametric test: exact Mann-Whitney "
                                 f"distribution too large on transcript {transcript.name}")
            return {f'{test}_intensity_pvalue': pv[L.PV[a]].copy(),
                    f'{test}_dwell_pvalue': pv[L.PV[b]].copy()}
        if key == 'GMM':
            out = {'GMM_chi2_pvalue': pv[L.PV['GMM_CHI2']].copy(),
                   'GMM_LOR': st[L.ST['GMM_LOR']].copy()}
            if self._with_logit:
                out['GMM_logit_pvalue'] = pv[L.PV['GMM_LOGIT']].copy()
                out['GMM_logit_coef'] = pv[L.PV['GMM_LOGIT_COEF']].copy()
            mode = self._config.get_cluster_counts()
            if mode in (HARD_ASSIGNMENT, SOFT_ASSIGNMENT):
                for label, sid in self._config.get_sample_ids().items():
                    mod, unmod = counts[sid, 0], counts[sid, 1]
                    if mode == SOFT_ASSIGNMENT:
                        mod, unmod = mod.view(np.float32), unmod.view(np.float32)
                    else:
                        mod, unmod = mod.astype(np.uint32), unmod.astype(np.uint32)
                    out[f'{label}_mod'] = mod.copy()
                    out[f'{label}_unmod'] = unmod.copy()
            return out
        raise NotImplementedError(f"Test {test} is not implemented!")

    def _merge_results(self, results, test_results, test, mask, auto_test_mask, n_positions):
        # comparisons.py:204-215
        has_auto = 'auto' in self._config.get_comparison_methods()
        if has_auto and 'auto_pvalue' not in results:
            results['auto_pvalue'] = np.full(n_positions, np.nan)
            results['auto_test'] = auto_test_mask
        for column, values in test_results.items():
            if mask.all():
                results[column] = values
            else:
                results.loc[mask, column] = values[mask]
            if has_auto and column == _AUTO_PVALUE.get(test):
                rows = auto_test_mask == test
                results.loc[rows, 'auto_pvalue'] = results.loc[rows, column]

    def _combine_context_pvalues(self, results, test):
        return self._combine_context(results.pos.to_numpy(), [results[test].to_numpy(dtype=float)])[0]

    def _combine_context(self, pos, columns, tx_offsets=None):
        """comparisons.py:436-474 for every p-value column of the transcript in one
        ``ncb_context_combine`` call: dense track, correlation of the track with
        its shifts, Hou's combination per row -- all in the context kernel.  Where the reference
        raises (cross_corr_matrix, comparisons.py:632-637) this raises the same error."""
        context = self._config.get_sequence_context()
        if self._config.get_sequence_context_weights() == "harmonic":
            weights = harmomic_series(context)
        else:
            weights = [1] * (2 * context + 1)
        if self._context_engine is None:
            # called on its own (the reference's tests do, tests/test_comparisons.py:157-206): the
            # combination runs on the device of the comparator's engine, cuda:0 if none exists yet
            self._engine(self._stage_device())
        if context > L.NCB_CTX_MAX:
            raise NanocomporeError(f"sequence_context {context} is larger than NCB_CTX_MAX = {L.NCB_CTX_MAX}")
        pos = np.asarray(pos).astype(np.int64)
        pvalues = np.stack([np.asarray(c, dtype=np.float64) for c in columns])
        combined, status = self._context_engine.context_combine(
            [0, pos.size] if tx_offsets is None else tx_offsets, pos, pvalues, context, weights)
        if (status == L.NCB_CTX_TOO_SHORT).any():
            raise RuntimeError("Not enough p-values for a context of order %s" % context)
        if (status == L.NCB_CTX_INVALID_PVALUE).any():
            raise RuntimeError("At least one p-value is invalid")
        return list(combined)

    def retry(self, fn, delay=5, backoff=5, max_attempts=3, exception=Exception):
        # comparisons.py:608-619; libncb sizes its scratch per batch, so OOM is not expected
        import time
        for attempt in range(1, max_attempts):
            try:
                return fn()
            except exception as e:
                self._worker.log("warning", f"Got error {e} on attempt {attempt}/{max_attempts}")
                time.sleep(delay)
                delay += backoff
        return fn()


def _to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


# ---------------------------------------------------------------------------------------
# Module-level helpers the reference exports (tests/test_comparisons.py:9-12, plotting.py:22).
# Small host-side tensor utilities: kept for API compatibility; compare_transcript does not
# use them (the kernels compute the same quantities per position).
# ---------------------------------------------------------------------------------------

def nanstd(X, dim):
    """Population standard deviation ignoring NaNs (comparisons.py:622-624)."""
    count = (~torch.isnan(X)).sum(dim)
    centred = X - X.nanmean(dim).unsqueeze(1)
    return torch.sqrt(torch.nansum(centred * centred, dim) / count)


def calculate_lors(contingencies):
    """log odds ratio of each 2x2 table, rounded to 3 decimals (comparisons.py:708-711)."""
    odds = contingencies[:, :, 0] / contingencies[:, :, 1]
    return torch.round(torch.log(odds[:, 0] / odds[:, 1]), decimals=3)


def get_contingency_matrices(conditions, predictions):
    """(positions, 2, 2) counts of condition x predicted cluster; NaN predictions never match
    (comparisons.py:714-729)."""
    cond = conditions.unsqueeze(0)
    tables = torch.zeros((predictions.shape[0], 2, 2))
    for c in (0, 1):
        for k in (0, 1):
            tables[:, c, k] = ((cond == c) & (predictions == k)).sum(1)
    return tables


def get_soft_contingency_matrices(conditions, cluster_probs):
    """(positions, 2, 2) sums of posteriors per condition (comparisons.py:732-744)."""
    tables = torch.zeros((cluster_probs.shape[0], 2, 2))
    for c in (0, 1):
        tables[:, c, :] = cluster_probs[:, conditions == c, :].nansum(1)
    return tables


def harmomic_series(sequence_context):
    """Symmetric harmonic weights 1/(|i|+1) (comparisons.py:654-656; the reference's spelling)."""
    return [1 / (abs(i) + 1) for i in range(-sequence_context, sequence_context + 1)]


def cross_corr_matrix(pvalues, context=2):
    """Correlation of the p-value track with its own shifts by -context..context
    (comparisons.py:627-651)."""
    pvalues = np.asarray(pvalues, dtype=float)
    if len(pvalues) < 3 * context + 3:
        raise RuntimeError("Not enough p-values for a context of order %s" % context)
    pvalues = np.nan_to_num(pvalues, nan=1)
    if (pvalues == 0).any() or np.isinf(pvalues).any() or (pvalues > 1).any():
        raise RuntimeError("At least one p-value is invalid")
    width = 2 * context + 1
    if (pvalues == 1).all():
        return np.ones((width, width))
    size = pvalues.size
    shifted = [np.roll(pvalues, s)[context:size - context] for s in range(-context, context + 1)]
    matrix = np.zeros((width, width))
    for x in range(width):
        for y in range(x + 1):
            matrix[x, y] = matrix[y, x] = np.corrcoef(shifted[x], shifted[y])[0, 1]
    return matrix


def combine_pvalues_hou(pvalues, weights, cor_mat):
    """Hou's approximation for a weighted Fisher combination of dependent p-values
    (comparisons.py:659-705; https://doi.org/10.1016/j.spl.2004.11.028)."""
    from scipy.stats import chi2
    if len(pvalues) != len(weights):
        raise ValueError("Can't combine pvalues if pvalues and weights are not the same length.")
    if cor_mat.shape[0] != cor_mat.shape[1] or cor_mat.shape[0] != len(pvalues):
        raise ValueError("The correlation matrix needs to be square, with each"
                         "dimension equal to the length of the pvalued vector.")
    p = np.asarray(pvalues, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float64)
    if (p == 1).all():
        return 1
    if (p == 0).any() or np.isinf(p).any() or (p > 1).any():
        raise ValueError("At least one p-value is invalid")
    k = len(p)
    # covariance of -2 ln p terms, Kost & McDermott eq. 8
    cov_sum = np.float64(0)
    for i in range(k):
        for j in range(i + 1, k):
            r = cor_mat[i][j]
            cov_sum += w[i] * w[j] * ((3.263 * r) + (0.710 * r ** 2) + (0.027 * r ** 3))
    sw_sum = np.float64(0)
    w_sum = np.float64(0)
    tau = np.float64(0)
    for i in range(k):
        sw_sum += w[i] ** 2
        w_sum += w[i]
        tau += w[i] * (-2 * np.log(p[i]))
    scale = (2 * sw_sum + cov_sum) / (2 * w_sum)
    dof = (4 * w_sum ** 2) / (2 * sw_sum + cov_sum)
    combined = chi2.sf(tau / scale, dof)
    if combined == 0:
        combined = np.finfo(np.float64).tiny
    return combined
