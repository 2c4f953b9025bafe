"""
Host side of the hot path: owns a libncb handle on one GPU, stages batches of positions
(dense reference tensors or ragged CSR batches), launches the kernels on the current torch
stream and returns the result columns.

PyTorch is used for device memory, pinned host memory and streams only; every number in the
results is produced by the CUDA kernels in nanocompore_b200/csrc (there is no CPU fallback:
without a GPU or without libncb.so this module raises).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from nanocompore_b200 import _lib as L

TEST_BITS = {'KS': L.NCB_TEST_KS, 'MW': L.NCB_TEST_MW, 'TT': L.NCB_TEST_TT, 'GMM': L.NCB_TEST_GMM,
             'LOGIT': L.NCB_TEST_LOGIT, 'AUTO': L.NCB_TEST_AUTO}
COUNT_MODES = {None: L.NCB_COUNTS_NONE, 'none': L.NCB_COUNTS_NONE, 'hard': L.NCB_COUNTS_HARD,
               'soft': L.NCB_COUNTS_SOFT}


def tests_mask(methods, with_logit=False):
    mask = 0
    for m in methods:
        key = m.upper()
        if key not in TEST_BITS:
            raise NotImplementedError(f"Test {m} is not implemented!")
        mask |= TEST_BITS[key]
    if with_logit:
        mask |= L.NCB_TEST_LOGIT
    return mask


def make_labels(samples, conditions):
    """label byte per read: bit 0 = condition (0 -> 0, anything else -> 1, as
    comparisons.py:228-234 splits on ``conditions == 0``), bits 1..7 = sample id."""
    samples = np.asarray(samples).astype(np.int64)
    conditions = np.asarray(conditions).astype(np.int64)
    if samples.size and (samples.min() < 0 or samples.max() >= L.NCB_MAX_SAMPLES):
        raise ValueError(f"sample ids must be in [0, {L.NCB_MAX_SAMPLES})")
    return ((samples << 1) | (conditions != 0)).astype(np.uint8)


@dataclass
class CsrBatch:
    """Ragged batch: the valid reads of each position, position after position.

    pos_offsets int64 [P+1]; signal [E, V] {intensity, dwell[, motor dwell]}, V = 2 or 3 -- float32
    (NCB_LAYOUT_CSR) or int16 codes (NCB_LAYOUT_CSR_I16: value = float(code), -32768 = NaN, the encoding
    of Uncalled4's tags, 5 instead of 9 bytes per entry on the host-to-device link); label uint8 [E].
    With ``run_ends`` (uint16 [P, n_runs]) and ``run_labels`` (numpy uint8 [n_runs]) an int16 batch travels
    as NCB_LAYOUT_CSR_I16_RUNS: the reads of a position come sample after sample, so the label bytes are
    rebuilt on the device from the run ends (4 bytes per entry on the link; ``label`` stays on the host).
    Tensors are either all pinned-host or all on the device."""
    pos_offsets: torch.Tensor
    signal: torch.Tensor
    label: torch.Tensor
    max_entries: int = 0      # most entries a position has, if the packer knows (0 = unknown): the
                              # size classes above it are not launched (ncb_batch.n_reads)
    run_ends: torch.Tensor = None
    run_labels: np.ndarray = None

    @property
    def n_positions(self):
        return self.pos_offsets.numel() - 1

    @property
    def n_entries(self):
        return self.signal.shape[0]

    def nbytes(self):
        """Bytes ``ncb_compare`` copies to the device for this batch."""
        last = self.run_ends if self.run_ends is not None else self.label
        return sum(t.numel() * t.element_size() for t in (self.pos_offsets, self.signal, last))

    @property
    def wire(self):
        if self.run_ends is not None:
            return 'i16r'
        return 'i16' if self.signal.dtype == torch.int16 else 'f32'

    def with_label_runs(self, threads=None):
        """This int16 batch with the run ends of its labels (NCB_LAYOUT_CSR_I16_RUNS), or None when the
        labels of a position do not come as runs in one common order (such a batch keeps its label bytes)."""
        import os
        if self.signal.dtype != torch.int16 or self.signal.is_cuda:
            raise ValueError("label runs are built for int16 batches on the host")
        lab = self.label.numpy()
        # run order: ascending sample id (the reference concatenates the samples in config order,
        # run.py:820-826, which is the order of their ids), else the order of the labels' first entries
        present, first = np.unique(lab, return_index=True)
        if present.size < 1 or present.size > L.NCB_MAX_RUNS:
            return None
        P = self.n_positions
        ends = torch.empty((P, present.size), dtype=torch.uint16, pin_memory=self.signal.is_pinned())
        by_first = present[np.argsort(first, kind='stable')]
        orders = [present] if np.array_equal(present, by_first) else [present, by_first]
        for order in orders:
            run_labels = np.ascontiguousarray(order, dtype=np.uint8)
            rc = L.load().ncb_pack_label_runs(self.pos_offsets.data_ptr(), self.label.data_ptr(), P,
                                              run_labels.ctypes.data, run_labels.size, ends.data_ptr(),
                                              threads or min(16, os.cpu_count() or 1))
            if rc == L.NCB_OK:
                break
        else:
            return None
        return CsrBatch(self.pos_offsets, self.signal, self.label, self.max_entries, ends, run_labels)

    def signal_f32(self):
        """The entries as float32 numbers whatever the wire format (int16 code -32768 = NaN)."""
        if self.signal.dtype != torch.int16:
            return self.signal
        out = self.signal.to(torch.float32)
        out[self.signal == -32768] = float('nan')
        return out

    def to(self, device, non_blocking=False):
        return CsrBatch(self.pos_offsets.to(device, non_blocking=non_blocking),
                        self.signal.to(device, non_blocking=non_blocking),
                        self.label.to(device, non_blocking=non_blocking), self.max_entries,
                        None if self.run_ends is None else self.run_ends.to(device, non_blocking=non_blocking),
                        self.run_labels)

    def pin(self):
        return CsrBatch(self.pos_offsets.pin_memory(), self.signal.pin_memory(), self.label.pin_memory(),
                        self.max_entries, None if self.run_ends is None else self.run_ends.pin_memory(),
                        self.run_labels)


def pack_csr(transcripts, pin=False, threads=None, wire='f32'):
    """Packs dense reference tensors into one CSR batch with libncb's host packer
    (ncb_pack_count / ncb_pack_fill, multi-threaded C++).

    ``wire``: 'f32' -> float32 entries; 'i16' -> int16 entries (every number of the tensors must be an
    integer in [-32767, 32767], as in Uncalled4 input: ValueError otherwise); 'i16r' -> int16 entries and
    label runs instead of label bytes (``CsrBatch.with_label_runs``; label bytes when the reads are not
    grouped by sample); 'auto' -> int16 when the data allows it, else float32.

    ``transcripts``: iterable of (data (P,R,V>=2) float32 with NaN, samples (R,), conditions (R,)).
    A read is kept at a position when any of its two variables is a number.  Read order inside
    a position is the read (column) order of the dense tensor.  Returns (CsrBatch,
    pos_transcript int32 [P_total], pos_index int32 [P_total])."""
    import os
    lib = L.load()
    threads = threads or min(16, os.cpu_count() or 1)
    items = []
    for data, samples, conditions in transcripts:
        data = np.ascontiguousarray(data, dtype=np.float32)
        assert data.ndim == 3 and data.shape[2] >= 2
        lab = make_labels(samples, conditions)
        if lab.shape != (data.shape[1],):
            raise ValueError(f"{data.shape[1]} reads in the tensor, {lab.size} sample / condition ids")
        items.append((data, lab))
    counts = []
    for data, lab in items:
        c = np.empty(data.shape[0], dtype=np.int64)
        rc = lib.ncb_pack_count(data.ctypes.data, data.shape[0], data.shape[1], data.shape[2],
                                c.ctypes.data, threads)
        if rc != L.NCB_OK:
            raise L.NcbError(rc, "ncb_pack_count")
        counts.append(c)
    all_counts = np.concatenate(counts) if counts else np.zeros(0, np.int64)
    offsets = np.zeros(all_counts.size + 1, dtype=np.int64)
    np.cumsum(all_counts, out=offsets[1:])
    E = int(offsets[-1])
    kw = dict(pin_memory=True) if pin else {}
    t_off = torch.from_numpy(offsets)
    if pin:
        t_off = t_off.pin_memory()
        offsets = t_off.numpy()
    if wire not in ('f32', 'i16', 'i16r', 'auto'):
        raise ValueError(f"wire must be 'f32', 'i16', 'i16r' or 'auto', not {wire!r}")
    want_runs, wire = wire == 'i16r', 'i16' if wire == 'i16r' else wire
    t_lab = torch.empty((E,), dtype=torch.uint8, **kw)
    ptx = [np.full(d.shape[0], ti, dtype=np.int32) for ti, (d, _) in enumerate(items)]
    pidx = [np.arange(d.shape[0], dtype=np.int32) for d, _ in items]

    def fill(i16):
        t_sig = torch.empty((E, 2), dtype=torch.int16 if i16 else torch.float32, **kw)
        fn = lib.ncb_pack_fill_i16 if i16 else lib.ncb_pack_fill
        p0 = 0
        for data, lab in items:
            P = data.shape[0]
            rc = fn(data.ctypes.data, lab.ctypes.data, P, data.shape[1], data.shape[2],
                    offsets[p0:].ctypes.data, t_sig.data_ptr(), t_lab.data_ptr(), threads)
            if rc != L.NCB_OK:
                if i16 and rc == L.NCB_ERR_INVALID:
                    return None      # a value is not an int16 code
                raise L.NcbError(rc, fn.__name__)
            p0 += P
        return t_sig

    t_sig = fill(True) if wire in ('i16', 'auto') else None
    if t_sig is None:
        if wire == 'i16':
            raise ValueError("pack_csr(wire='i16'): the data holds values that are not int16 codes")
        t_sig = fill(False)
    batch = CsrBatch(t_off, t_sig, t_lab, int(all_counts.max()) if all_counts.size else 0)
    if want_runs and E > 0:
        batch = batch.with_label_runs(threads) or batch
    return (batch,
            np.concatenate(ptx) if ptx else np.zeros(0, np.int32),
            np.concatenate(pidx) if pidx else np.zeros(0, np.int32))


def pack_csr_numpy(transcripts):
    """Reference packing in numpy (used by the tests to check the native packer)."""
    offs, sigs, labs = [np.zeros(1, dtype=np.int64)], [], []
    total = 0
    for data, samples, conditions in transcripts:
        data = np.asarray(data, dtype=np.float32)
        lab = make_labels(samples, conditions)
        valid = ~np.isnan(data[:, :, :2]).all(2)
        counts = valid.sum(1).astype(np.int64)
        sigs.append(data[:, :, :2][valid])
        labs.append(np.broadcast_to(lab[None, :], valid.shape)[valid])
        offs.append(total + np.cumsum(counts))
        total += int(counts.sum())
    sig = np.concatenate(sigs) if sigs else np.zeros((0, 2), np.float32)
    lab = np.concatenate(labs) if labs else np.zeros((0,), np.uint8)
    return np.concatenate(offs), sig.reshape(-1, 2), lab


class Results:
    """Result columns of one batch (struct of arrays, one slot per position)."""

    def __init__(self, n_positions, n_samples, device, diagnostics=False, pinned=True):
        P, S = n_positions, n_samples
        self.n_positions, self.n_samples = P, S
        on_dev = device is not None
        kw = dict(device=device) if on_dev else dict(pin_memory=pinned)
        self.memory = L.NCB_MEM_DEVICE if on_dev else L.NCB_MEM_HOST
        self.status = torch.empty(P, dtype=torch.int32, **kw)
        if on_dev:
            # the kernels write every row of device results
            self.pvalues = torch.empty((L.NCB_PV_ROWS, P), dtype=torch.float64, **kw)
            self.stats = torch.empty((L.NCB_ST_ROWS, P), dtype=torch.float32, **kw)
            self.counts = torch.empty((max(S, 1), 2, P), dtype=torch.int32, **kw)
        else:
            # host results receive only the rows of the configured tests (ncb.h, ncb_results): the
            # others read as "not tested"
            self.pvalues = torch.full((L.NCB_PV_ROWS, P), float('nan'), dtype=torch.float64, **kw)
            self.stats = torch.full((L.NCB_ST_ROWS, P), float('nan'), dtype=torch.float32, **kw)
            self.counts = torch.zeros((max(S, 1), 2, P), dtype=torch.int32, **kw)
        self.diag_i64 = torch.empty((L.NCB_DI_ROWS, P), dtype=torch.int64, **kw) if diagnostics else None
        self.diag_f64 = torch.empty((L.NCB_DF_ROWS, P), dtype=torch.float64, **kw) if diagnostics else None

    def carve(self, n_positions):
        """Result buffers of exactly ``n_positions`` <= P positions carved from this object's memory
        (libncb writes rows of length P back to back, so a smaller batch needs arrays of its own
        shape): pinned buffers allocated once for the largest batch serve batches of any size.  Rows
        the configured tests do not write hold whatever the memory held before."""
        n = int(n_positions)
        if n == self.n_positions:
            return self
        if n > self.n_positions:
            raise ValueError(f"carve({n}) from result buffers of {self.n_positions} positions")
        v = Results.__new__(Results)
        v.n_positions, v.n_samples, v.memory = n, self.n_samples, self.memory
        v.status = self.status[:n]
        v.pvalues = self.pvalues.view(-1)[:L.NCB_PV_ROWS * n].view(L.NCB_PV_ROWS, n)
        v.stats = self.stats.view(-1)[:L.NCB_ST_ROWS * n].view(L.NCB_ST_ROWS, n)
        s = self.counts.shape[0]
        v.counts = self.counts.view(-1)[:s * 2 * n].view(s, 2, n)
        v.diag_i64 = None if self.diag_i64 is None else self.diag_i64.view(-1)[:L.NCB_DI_ROWS * n].view(L.NCB_DI_ROWS, n)
        v.diag_f64 = None if self.diag_f64 is None else self.diag_f64.view(-1)[:L.NCB_DF_ROWS * n].view(L.NCB_DF_ROWS, n)
        return v

    def struct(self):
        r = L.NcbResults()
        r.memory = self.memory
        r.status = self.status.data_ptr()
        r.pvalues = self.pvalues.data_ptr()
        r.stats = self.stats.data_ptr()
        r.counts = self.counts.data_ptr() if self.n_samples > 0 else None
        r.diag_i64 = self.diag_i64.data_ptr() if self.diag_i64 is not None else None
        r.diag_f64 = self.diag_f64.data_ptr() if self.diag_f64 is not None else None
        return r

    def nbytes(self):
        ts = [self.status, self.pvalues, self.stats, self.counts, self.diag_i64, self.diag_f64]
        return sum(t.numel() * t.element_size() for t in ts if t is not None)

    # numpy views (host results, or device results copied back)
    def numpy(self):
        def get(t):
            return None if t is None else t.cpu().numpy()
        out = dict(status=get(self.status), pvalues=get(self.pvalues), stats=get(self.stats),
                   counts=get(self.counts), diag_i64=get(self.diag_i64), diag_f64=get(self.diag_f64))
        return out


class Engine:
    """One libncb handle on one CUDA device."""

    def __init__(self, device=0, *, tests=('GMM', 'KS'), n_samples=4, depleted_condition=0,
                 min_coverage=0, dwell_is_raw=False, cluster_counts='hard', with_logit=False,
                 em_max_iter=100, em_tol=1e-3, reg_covar=1e-6, em_fp64=True, input_standardised=False,
                 lib_path=None):
        if not torch.cuda.is_available():
            raise RuntimeError("nanocompore_b200 needs a CUDA device: there is no CPU fallback")
        self.lib = L.load(lib_path)
        if isinstance(device, str):
            device = torch.device(device)
        if isinstance(device, torch.device):
            device = device.index if device.index is not None else torch.cuda.current_device()
        self.device_index = int(device)
        self.device = torch.device('cuda', self.device_index)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)   # make sure the primary context exists
        self.opts = L.default_opts()
        self._handle = C.c_void_p()
        self.configure(tests=tests, n_samples=n_samples, depleted_condition=depleted_condition,
                       min_coverage=min_coverage, dwell_is_raw=dwell_is_raw,
                       cluster_counts=cluster_counts, with_logit=with_logit,
                       em_max_iter=em_max_iter, em_tol=em_tol, reg_covar=reg_covar, em_fp64=em_fp64,
                       input_standardised=input_standardised, _apply=False)
        rc = self.lib.ncb_create(C.byref(self._handle), self.device_index, C.byref(self.opts))
        if rc != L.NCB_OK:
            self._handle = C.c_void_p()
            raise L.NcbError(rc, "ncb_create failed (no usable GPU / library built for another arch)")

    # ------------------------------------------------------------------ options
    def configure(self, *, tests=None, n_samples=None, depleted_condition=None, min_coverage=None,
                  dwell_is_raw=None, cluster_counts='__keep__', with_logit=None, em_max_iter=None,
                  em_tol=None, reg_covar=None, em_fp64=None, input_standardised=None, _apply=True):
        o = self.opts
        if tests is not None:
            logit = bool(with_logit) if with_logit is not None else bool(o.tests & L.NCB_TEST_LOGIT)
            o.tests = tests if isinstance(tests, int) else tests_mask(tests, logit)
        elif with_logit is not None:
            o.tests = (o.tests & ~L.NCB_TEST_LOGIT) | (L.NCB_TEST_LOGIT if with_logit else 0)
        if n_samples is not None:
            o.n_samples = int(n_samples)
        if depleted_condition is not None:
            o.depleted_condition = int(depleted_condition)
        if min_coverage is not None:
            o.min_coverage = int(min_coverage)
        if dwell_is_raw is not None:
            o.dwell_is_raw = int(bool(dwell_is_raw))
        if cluster_counts != '__keep__':
            o.cluster_counts = COUNT_MODES[cluster_counts]
        if em_max_iter is not None:
            o.em_max_iter = int(em_max_iter)
        if em_tol is not None:
            o.em_tol = float(em_tol)
        if reg_covar is not None:
            o.reg_covar = float(reg_covar)
        if em_fp64 is not None:
            o.em_fp64 = int(bool(em_fp64))
        if input_standardised is not None:
            o.input_standardised = int(bool(input_standardised))
        if _apply:
            self._check(self.lib.ncb_set_opts(self._handle, C.byref(o)))

    def _check(self, rc):
        if rc == L.NCB_OK:
            return
        msg = self.lib.ncb_last_error(self._handle).decode(errors='replace')
        if rc == L.NCB_ERR_NOMEM:
            # keeps the caller's re-queue logic working (run.py:467-480)
            raise torch.OutOfMemoryError(f"libncb: {msg}")
        raise L.NcbError(rc, msg)

    def close(self):
        if self._handle:
            self.lib.ncb_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ hot path
    def _stream(self, stream=None):
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        return C.c_void_p(s.cuda_stream)

    def compare_csr(self, batch: CsrBatch, results: Results = None, *, diagnostics=False,
                    stream=None, wait=True):
        """Runs the hot path on a CSR batch (host-pinned or device tensors).  Results are
        allocated where the batch lives unless ``results`` is given."""
        on_dev = batch.signal.is_cuda
        P = batch.n_positions
        if results is None:
            results = Results(P, self.opts.n_samples, self.device if on_dev else None, diagnostics)
        b = L.NcbBatch()
        b.layout = L.NCB_LAYOUT_CSR
        b.memory = L.NCB_MEM_DEVICE if on_dev else L.NCB_MEM_HOST
        b.n_positions = P
        b.n_entries = batch.n_entries
        assert batch.pos_offsets.dtype == torch.int64 and batch.label.dtype == torch.uint8
        assert batch.signal.dtype in (torch.float32, torch.int16)
        if batch.signal.dtype == torch.int16:
            b.layout = L.NCB_LAYOUT_CSR_I16
        assert batch.signal.is_contiguous() and batch.label.is_contiguous()
        b.pos_offsets = batch.pos_offsets.data_ptr()
        b.signal = batch.signal.data_ptr()
        b.label = batch.label.data_ptr()
        if batch.run_ends is not None:
            assert batch.run_ends.dtype == torch.uint16 and batch.run_ends.is_contiguous()
            assert batch.run_ends.is_cuda == on_dev and batch.run_ends.shape == (P, batch.run_labels.size)
            b.layout = L.NCB_LAYOUT_CSR_I16_RUNS
            b.label = batch.run_ends.data_ptr()
            b.run_labels = batch.run_labels.ctypes.data
            b.n_runs = batch.run_labels.size
        b.n_vars = batch.signal.shape[1]     # 2, or 3 with the motor-dwell column
        b.n_reads = min(int(batch.max_entries), 0x7fffffff)
        r = results.struct()
        s = self._stream(stream)
        self._check(self.lib.ncb_compare(self._handle, C.byref(b), C.byref(r), s))
        if wait:
            self._check(self.lib.ncb_wait(self._handle, s))
        return results

    def compare_struct(self, batch_struct, results: Results, *, stream=None, wait=True):
        """Runs the hot path on an ``ncb_batch`` that is already filled in (the device-resident batch
        of ``ncb_decoder_launch``); ``results`` may live on the host or on the device."""
        r = results.struct()
        s = self._stream(stream)
        self._check(self.lib.ncb_compare(self._handle, C.byref(batch_struct), C.byref(r), s))
        if wait:
            self._check(self.lib.ncb_wait(self._handle, s))
        return results

    def compare_dense(self, data: torch.Tensor, labels: torch.Tensor, results: Results = None, *,
                      diagnostics=False, stream=None, wait=True):
        """Runs the hot path on the reference's dense tensor ``data`` (P, R, V) float32 with
        NaN for missing reads, ``labels`` uint8 (R,) from ``make_labels``."""
        assert data.dtype == torch.float32 and data.dim() == 3 and data.is_contiguous()
        assert labels.dtype == torch.uint8 and labels.is_contiguous()
        on_dev = data.is_cuda
        assert labels.is_cuda == on_dev
        P, R, V = data.shape
        if labels.numel() != R:
            raise ValueError(f"{R} reads in the tensor, {labels.numel()} labels")
        if results is None:
            results = Results(P, self.opts.n_samples, self.device if on_dev else None, diagnostics,
                              pinned=False)
        b = L.NcbBatch()
        b.layout = L.NCB_LAYOUT_DENSE
        b.memory = L.NCB_MEM_DEVICE if on_dev else L.NCB_MEM_HOST
        b.n_positions = P
        b.n_entries = P * R
        b.pos_offsets = None
        b.signal = data.data_ptr()
        b.label = labels.data_ptr()
        b.n_reads = R
        b.n_vars = V
        r = results.struct()
        s = self._stream(stream)
        self._check(self.lib.ncb_compare(self._handle, C.byref(b), C.byref(r), s))
        if wait:
            self._check(self.lib.ncb_wait(self._handle, s))
        return results

    def wait(self, stream=None):
        self._check(self.lib.ncb_wait(self._handle, self._stream(stream)))

    def results_host_bytes(self, n_positions, n_vars=2, with_counts=True):
        """Bytes one ``compare_*`` call with host results copies device -> host."""
        return int(self.lib.ncb_results_host_bytes(self._handle, n_positions, n_vars, int(with_counts)))

    # ------------------------------------------------------------------ unit entry points
    def ks_exact_pvalues(self, n0, n1, h):
        n0 = np.ascontiguousarray(n0, dtype=np.int32)
        n1 = np.ascontiguousarray(n1, dtype=np.int32)
        h = np.ascontiguousarray(h, dtype=np.int64)
        out = np.empty(n0.size, dtype=np.float64)
        self._check(self.lib.ncb_ks_exact_pvalues(
            self._handle, n0.size, n0.ctypes.data, n1.ctypes.data, h.ctypes.data, out.ctypes.data,
            L.NCB_MEM_HOST, self._stream()))
        return out

    def bh_fdr(self, pvalues, scale=1.0, above=0, total=None):
        """Benjamini-Hochberg q-values, NaNs skipped (postprocessing.py:255-285).  Accepts a
        numpy array (host) or a CUDA float64 tensor (stays on the device).  ``above`` / ``total``:
        the values are one value range of a column of ``total`` numbers, ``above`` of them larger
        than all of these (``ncb_bh_fdr_segment``; the distributed correction of sharding.py)."""
        m_total = -1 if total is None else int(total)
        if isinstance(pvalues, torch.Tensor) and pvalues.is_cuda:
            p = pvalues.contiguous().to(torch.float64)
            q = torch.empty_like(p)
            self._check(self.lib.ncb_bh_fdr_segment(self._handle, p.numel(), p.data_ptr(), q.data_ptr(),
                                                    float(scale), int(above), m_total, L.NCB_MEM_DEVICE,
                                                    self._stream()))
            return q
        p = np.ascontiguousarray(pvalues, dtype=np.float64)
        q = np.empty_like(p)
        self._check(self.lib.ncb_bh_fdr_segment(self._handle, p.size, p.ctypes.data, q.ctypes.data,
                                                float(scale), int(above), m_total, L.NCB_MEM_HOST,
                                                self._stream()))
        return q

    def context_combine(self, tx_offsets, pos, pvalues, context, weights):
        """Hou's sequence-context combination (comparisons.py:436-474, 627-705) of ``pvalues``
        [n_cols, n] over the transcripts delimited by ``tx_offsets`` (rows of transcript t:
        tx_offsets[t]..tx_offsets[t+1]; ``pos`` = reference position of every row).  Returns
        (combined [n_cols, n], status [n_cols, n_tx]) -- ``status`` holds NCB_CTX_* where the
        reference raises.  ``pvalues`` may be a numpy array (host) or a CUDA float64 tensor."""
        tx_offsets = np.ascontiguousarray(tx_offsets, dtype=np.int64)
        pos = np.ascontiguousarray(pos, dtype=np.int64)
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        context = int(context)
        if weights.size != 2 * context + 1:
            raise ValueError("context_combine: 2 * context + 1 weights are needed")
        n_tx = tx_offsets.size - 1
        on_device = isinstance(pvalues, torch.Tensor) and pvalues.is_cuda
        if on_device:
            p = pvalues.to(torch.float64).contiguous()
            p = p.reshape(1, -1) if p.dim() == 1 else p
            out = torch.empty_like(p)
            p_ptr, out_ptr, n_cols, n = p.data_ptr(), out.data_ptr(), p.shape[0], p.shape[1]
        else:
            p = np.ascontiguousarray(pvalues, dtype=np.float64)
            p = p.reshape(1, -1) if p.ndim == 1 else p
            out = np.empty_like(p)
            p_ptr, out_ptr, n_cols, n = p.ctypes.data, out.ctypes.data, p.shape[0], p.shape[1]
        if n != pos.size or (n_tx >= 0 and tx_offsets[-1] != n):
            raise ValueError("context_combine: tx_offsets / pos / pvalues disagree on the number of rows")
        status = np.zeros((n_cols, max(n_tx, 0)), dtype=np.int32)
        self._check(self.lib.ncb_context_combine(
            self._handle, n_tx, tx_offsets.ctypes.data, pos.ctypes.data, n_cols, p_ptr, context,
            weights.ctypes.data, out_ptr, status.ctypes.data,
            L.NCB_MEM_DEVICE if on_device else L.NCB_MEM_HOST, self._stream()))
        return out, status

    # ------------------------------------------------------------------ introspection
    def kernel_launches(self):
        return int(self.lib.ncb_kernel_launches(self._handle))

    def set_timing(self, enabled=True):
        self._check(self.lib.ncb_set_timing(self._handle, int(enabled)))

    def last_timing(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._check(self.lib.ncb_last_timing(self._handle, C.byref(a), C.byref(b), C.byref(c)))
        return dict(ms_position=a.value, ms_ks=b.value, ms_total=c.value)

    def probe_fp64_peak(self):
        """Measured non-tensor fp64 rate of the device in TFLOP/s (DFMA microbenchmark)."""
        t = C.c_double()
        self._check(self.lib.ncb_probe_fp64_peak(self._handle, C.byref(t), self._stream()))
        return t.value

    def timing_totals(self):
        """Sums of the CUDA-event timings of every batch since ``set_timing(True)``."""
        n, a, b, c, d = C.c_int64(), C.c_double(), C.c_double(), C.c_double(), C.c_double()
        self._check(self.lib.ncb_timing_totals(self._handle, C.byref(n), C.byref(a), C.byref(b),
                                               C.byref(c), C.byref(d)))
        i, e, f = C.c_double(), C.c_double(), C.c_double()
        self._check(self.lib.ncb_timing_mixture(self._handle, C.byref(i), C.byref(e), C.byref(f)))
        return dict(batches=n.value, ms_classify=a.value, ms_stats=b.value, ms_gmm=c.value,
                    ms_position=b.value + c.value, ms_ks=d.value,
                    ms_gmm_init=i.value, ms_gmm_em=e.value, ms_gmm_final=f.value)
