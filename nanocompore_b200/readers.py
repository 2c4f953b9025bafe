"""
Transcript readers: signal files -> pinned CSR batches (north_star item 1).

Mirror of the two ``_read_data`` implementations of the reference and of what follows them in
``Worker.run`` up to the comparison (/root/reference/nanocompore):

  * ``Uncalled4Reader.read``   = ``Uncalled4Worker._read_data`` (run.py:657-718) over
    ``Uncalled4.get_data`` (uncalled4.py:57-201): BAM records of the transcript, ``ur/uc/ul`` tags,
    reads ordered by read id, invalid-kmer filter (common.py:298-329)
  * ``DbReader.read``          = ``GenericWorker._read_data`` (run.py:741-836) over
    ``PreprocessingDB.get_signal_data`` (database.py:125-141, 425-452): float32 blobs per read
    from the SQLite ``signal_data`` table, samples in config order
  * ``select_reads``           = ``Worker._downsample`` (run.py:518-529) as a read selection
  * ``pack_rows``              = the hand-over: rows of many transcripts -> ONE ragged CSR batch in
    pinned memory (libncb ``ncb_pack_rows_*``), which ``Engine.compare_csr`` streams to HBM
  * ``TranscriptPipeline``     = the body of ``Worker.run`` between "task received" and "results
    frame" (run.py:420-463) for a list of transcripts at a time

What differs from the reference: a read is never expanded into the dense, mostly-NaN
``(positions, reads, vars)`` tensor.  Every read is a pair of planar rows (intensity, dwell) --
the SQLite blobs as they are, or rows decoded from the BAM tags by libncb's native BGZF/BAM
reader -- and the packer gathers the valid entries position by position straight into the CSR
arrays.  ``TranscriptRows.dense()`` rebuilds the reference's tensor for callers (and tests) that
want it.  BAM decoding and packing are native code in libncb.so; without it this module raises.
"""
import ctypes as C
import os
import sqlite3
from concurrent.futures import ThreadPoolExecutor
from contextlib import closing

import numpy as np
import torch

from nanocompore_b200 import _lib as L
from nanocompore_b200.engine import CsrBatch, make_labels

# common.py:21-32: kmer length and (one-based) most influential position of the pore models
KITS = {'RNA002': (5, 4), 'RNA004': (9, 5)}

# database.py:125-132
GET_SIGNAL_DATA_FOR_TRANSCRIPT_QUERY = """
SELECT intensity, dwell
FROM signal_data sd
JOIN reads r ON sd.read_id = r.id
JOIN transcripts t ON sd.transcript_id = t.id
WHERE t.name = ?
  AND r.invalid_kmers <= ?
"""


def kit_geometry(kit):
    """(len, center) of a kit given as a name or as the reference's ``Kit`` enum member."""
    if hasattr(kit, 'len') and hasattr(kit, 'center'):
        return int(kit.len), int(kit.center)
    name = getattr(kit, 'name', kit)
    if name not in KITS:
        raise ValueError(f"unknown kit {kit!r}")
    return KITS[name]


_BASE_CODE = {'A': 0, 'a': 0, 'G': 1, 'g': 1, 'C': 2, 'c': 2, 'T': 3, 't': 3}


def encode_kmer(kmer):
    """common.encode_kmer (common.py:224-238): two bits per base, A G C T = 0 1 2 3."""
    code = 0
    for base in kmer:
        if base not in _BASE_CODE:
            raise RuntimeError(f"Bad nucleotide in kmer: {kmer}")
        code = (code << 2) | _BASE_CODE[base]
    return code


_CODE_TABLE = np.full(256, -1, dtype=np.int64)
for _b, _c in _BASE_CODE.items():
    _CODE_TABLE[ord(_b)] = _c


def encode_kmers(seq, positions, klen):
    """``encode_kmer(seq[p:p + klen])`` for every p of ``positions`` at once (the ``kmer`` column of
    run.py:461-463 is a Python call per result row in the reference).  A k-mer cut short by the end of
    the sequence is encoded over the bases it has, like the slice; a bad nucleotide raises the scalar
    function's error."""
    pos = np.asarray(positions, dtype=np.int64)
    codes = _CODE_TABLE[np.frombuffer(seq.encode('latin-1'), dtype=np.uint8)]
    out = np.zeros(pos.size, dtype=np.int64)
    full = (pos >= 0) & (pos + klen <= codes.size)
    p = pos[full]
    acc = np.zeros(p.size, dtype=np.int64)
    bad = np.zeros(p.size, dtype=bool)
    for k in range(klen):
        c = codes[p + k]
        bad |= c < 0
        acc = (acc << 2) | np.maximum(c, 0)
    if bad.any():
        q = int(p[np.nonzero(bad)[0][0]])
        encode_kmer(seq[q:q + klen])                # raises with the reference's message
    out[full] = acc
    for i in np.nonzero(~full)[0]:
        out[i] = encode_kmer(seq[pos[i]:pos[i] + klen])
    return out


def _threads():
    return min(16, os.cpu_count() or 1)


class TranscriptRows:
    """The reads of one transcript as planar rows, in the read order the comparison must see.

    ``intensity`` / ``dwell`` are float32 arrays (R, n_positions); NaN = not covered; the dwell is
    raw (seconds or samples: the log10 of run.py:573 is taken inside the statistics kernel).
    ``sample_ids`` / ``condition_ids`` are the integer ids of config.py:443-459."""

    def __init__(self, n_positions, intensity, dwell, sample_ids, condition_ids, read_names=None):
        self.n_positions = int(n_positions)
        n_rows = len(sample_ids)     # (R, 0) is a legal shape: no rows / no positions are empty sets
        self.intensity = np.ascontiguousarray(intensity, dtype=np.float32).reshape(n_rows, self.n_positions)
        self.dwell = np.ascontiguousarray(dwell, dtype=np.float32).reshape(n_rows, self.n_positions)
        assert self.intensity.shape == self.dwell.shape
        self.sample_ids = np.asarray(sample_ids, dtype=np.int64)
        self.condition_ids = np.asarray(condition_ids, dtype=np.int64)
        self.read_names = read_names
        assert self.sample_ids.shape == (self.n_reads,) and self.condition_ids.shape == (self.n_reads,)

    @property
    def n_reads(self):
        return self.intensity.shape[0]

    def take(self, index):
        """Rows ``index`` (a boolean mask or an index array), order given by the index."""
        names = None if self.read_names is None else np.asarray(self.read_names)[index]
        return TranscriptRows(self.n_positions, self.intensity[index], self.dwell[index],
                              self.sample_ids[index], self.condition_ids[index], names)

    def labels(self):
        return make_labels(self.sample_ids, self.condition_ids)

    def row_pointers(self):
        stride = self.n_positions * 4
        idx = np.arange(self.n_reads, dtype=np.uint64) * np.uint64(stride)
        return (np.uint64(self.intensity.ctypes.data) + idx, np.uint64(self.dwell.ctypes.data) + idx)

    def invalid_ratios(self):
        """common.get_reads_invalid_ratio (common.py:298-329) per read."""
        out = np.empty(self.n_reads, dtype=np.float64)
        if self.n_reads:
            ip, _ = self.row_pointers()
            rc = L.load().ncb_rows_invalid_ratio(ip.ctypes.data, self.n_reads, self.n_positions,
                                                 out.ctypes.data, _threads())
            if rc != L.NCB_OK:
                raise L.NcbError(rc, "ncb_rows_invalid_ratio")
        return out

    def dense(self):
        """The reference's ``(positions, reads, 2)`` float32 tensor (run.py:833-838)."""
        return np.ascontiguousarray(np.stack([self.intensity.T, self.dwell.T], axis=2))


class NativeBam:
    """libncb's BGZF/BAM reader (the slice of ``pysam.AlignmentFile`` the path needs)."""

    def __init__(self, path, threads=None):
        self._lib = L.load()
        self._h = C.c_void_p()
        self.path = str(path)
        rc = self._lib.ncb_bam_open(C.byref(self._h), self.path.encode(), threads or _threads())
        if rc != L.NCB_OK:
            msg = self._lib.ncb_bam_last_error(self._h).decode(errors='replace') if self._h else \
                f"cannot open {self.path}"
            self.close()
            raise L.NcbError(rc, f"{self.path}: {msg}")
        n = self._lib.ncb_bam_n_references(self._h)
        self.references = [self._lib.ncb_bam_reference_name(self._h, i).decode(errors='replace') for i in range(n)]
        self.lengths = [int(self._lib.ncb_bam_reference_length(self._h, i)) for i in range(n)]

    def count(self, reference):
        """Records placed on the reference (pysam ``count(reference=...)``, uncalled4.py:76-77)."""
        i = self._lib.ncb_bam_reference_index(self._h, reference.encode())
        return 0 if i < 0 else int(self._lib.ncb_bam_reference_records(self._h, i))

    def mapped_references(self):
        """reference name -> records, for references with data (common.py:284-295)."""
        return {name: self.count(name) for name in self.references if self.count(name) > 0}

    def read_uncalled4(self, reference, ref_len, kit):
        """Primary reads of ``reference``: (intensity (R, ref_len), dwell (R, ref_len), names)."""
        klen, kcenter = kit_geometry(kit)
        i = self._lib.ncb_bam_reference_index(self._h, reference.encode())
        cap = 0 if i < 0 else int(self._lib.ncb_bam_reference_records(self._h, i))
        stride = 256       # a BAM read name has at most 254 characters + NUL: never truncated
        inten = np.empty((cap, ref_len), dtype=np.float32)
        dwell = np.empty((cap, ref_len), dtype=np.float32)
        names = np.zeros((max(cap, 1), stride), dtype=np.uint8)
        n = 0
        if cap:
            n = self._lib.ncb_bam_read_uncalled4(self._h, i, ref_len, klen, kcenter, inten.ctypes.data,
                                                 dwell.ctypes.data, names.ctypes.data, stride, cap)
            if n < 0:
                raise L.NcbError(int(n), self._lib.ncb_bam_last_error(self._h).decode(errors='replace'))
        read_names = [bytes(names[k]).split(b'\0', 1)[0].decode(errors='replace') for k in range(n)]
        return inten[:n], dwell[:n], read_names

    def close(self):
        if self._h:
            self._lib.ncb_bam_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Uncalled4Reader:
    """``Uncalled4Worker.setup`` + ``_read_data`` (run.py:657-718)."""

    # several transcripts may be read at once: the native handles are read-only after the index is built
    concurrent_reads = True
    int16_source = True       # tag values are int16: the pipeline packs NCB_LAYOUT_CSR_I16 entries

    def __init__(self, config, threads=None):
        self._conf = config
        self._bams = {sample: NativeBam(bam, threads)
                      for sample, _, bam in config.get_sample_condition_bam_data()}

    def read(self, transcript_name, ref_len):
        conf = self._conf
        sample_ids_of = conf.get_sample_ids()
        cond_of = conf.sample_to_condition()
        cond_ids = conf.get_condition_ids()
        kit = conf.get_kit()
        with ThreadPoolExecutor(max_workers=max(len(self._bams), 1)) as ex:
            parts = list(ex.map(lambda kv: (kv[0],) + kv[1].read_uncalled4(transcript_name, ref_len, kit),
                                self._bams.items()))
        names = np.array([n for p in parts for n in p[3]], dtype=str)
        samples = np.array([sample_ids_of[p[0]] for p in parts for _ in p[3]], dtype=np.int64)
        conditions = np.array([cond_ids[cond_of[p[0]]] for p in parts for _ in p[3]], dtype=np.int64)
        # run.py:698-700: reads ordered by read id (the reference's argsort is not stable; equal
        # ids -- one read in two files -- keep the config order of their samples here)
        order = np.argsort(names, kind='stable')
        # the rows of every sample go straight to their place in that order: one copy of the signal
        # (concatenating the samples and then reordering made two)
        dest = np.empty_like(order)
        dest[order] = np.arange(order.size)
        inten = np.empty((order.size, ref_len), np.float32)
        dwell = np.empty((order.size, ref_len), np.float32)
        first = 0
        for p in parts:
            where = dest[first:first + len(p[3])]
            inten[where] = p[1]
            dwell[where] = p[2]
            first += len(p[3])
        rows = TranscriptRows(ref_len, inten, dwell, samples[order], conditions[order], names[order])
        # run.py:711-718
        valid = rows.invalid_ratios() < conf.get_max_invalid_kmers_freq()
        return rows if valid.all() else rows.take(valid)

    def references(self):
        """Transcripts with data in any sample."""
        out = {}
        for bam in self._bams.values():
            for name, n in bam.mapped_references().items():
                out[name] = out.get(name, 0) + n
        return out

    def close(self):
        for bam in self._bams.values():
            bam.close()


class DbReader:
    """``GenericWorker.setup`` + ``_read_data`` (run.py:726-848) over the SQLite stores written by
    the reference's preprocessing (eventalign_collapse / Remora)."""

    concurrent_reads = False      # one sqlite3 connection per sample: transcripts are read one after the other

    def __init__(self, config):
        self._conf = config
        self._dbs = {sample: sqlite3.connect(db, check_same_thread=False)
                     for sample, _, db in config.get_sample_condition_db_data()}

    def _signal_data(self, sample, transcript_name):
        with closing(self._dbs[sample].cursor()) as cursor:   # database.py:449-452
            return cursor.execute(GET_SIGNAL_DATA_FOR_TRANSCRIPT_QUERY,
                                  (transcript_name, self._conf.get_max_invalid_kmers_freq())).fetchall()

    def read(self, transcript_name, ref_len=None):
        conf = self._conf
        labels = conf.get_sample_labels()
        with ThreadPoolExecutor(max_workers=max(len(labels), 1)) as ex:
            rows = dict(zip(labels, ex.map(lambda s: self._signal_data(s, transcript_name), labels)))
        counts = {s: len(rows[s]) for s in labels}
        if all(c == 0 for c in counts.values()):
            return TranscriptRows(ref_len or 0, np.empty((0, ref_len or 0), np.float32),
                                  np.empty((0, ref_len or 0), np.float32), [], [])
        # every blob must be a float32 row of the transcript's length: the reference's assignment
        # into its (positions, reads, 2) tensor raises otherwise (run.py:795-801)
        row_bytes = 4 * ref_len if ref_len is not None else len(next(i for s in labels for i, _ in rows[s]))
        for s in labels:
            for i, d in rows[s]:
                if len(i) != row_bytes or len(d) != row_bytes:
                    raise ValueError(f"a signal blob of {transcript_name} in sample {s} has {len(i)} / {len(d)} bytes, "
                                     f"expected {row_bytes}")
        # blobs are float32 rows of length ref_len (database.py:69-76); one copy each into a
        # contiguous (R, P) block, samples in config order (run.py:820-826)
        inten = np.frombuffer(b''.join(i for s in labels for i, _ in rows[s]), dtype=np.float32)
        dwell = np.frombuffer(b''.join(d for s in labels for _, d in rows[s]), dtype=np.float32)
        n_reads = sum(counts.values())
        n_positions = inten.size // n_reads
        if ref_len is not None and n_positions != ref_len:
            raise ValueError(f"signal blobs of {transcript_name} have {n_positions} positions, expected {ref_len}")
        sample_ids = np.concatenate([np.full(counts[s], sid) for s, sid in conf.get_sample_ids().items()])
        s2c, cids = conf.sample_to_condition(), conf.get_condition_ids()
        condition_ids = np.concatenate([np.full(counts[s], cids[s2c[s]]) for s in conf.get_sample_ids()])
        return TranscriptRows(n_positions, inten, dwell, sample_ids, condition_ids)

    def close(self):
        for db in self._dbs.values():
            db.close()


def coverage(rows):
    """Valid intensities per position and condition: (cov0, cov1), run.py:511-515."""
    valid = ~np.isnan(rows.intensity)
    c0 = rows.condition_ids == 0
    return valid[c0].sum(0), valid[rows.condition_ids == 1].sum(0)


def select_reads(rows, max_reads, min_coverage, motor_dwell_offset=0):
    """``Worker._downsample`` (run.py:518-529) as a selection of rows: when the transcript has
    more than ``max_reads`` reads, per condition the ``max_reads`` reads with the most positions
    whose every variable is a number are kept; positions are those that pass the coverage filter
    and have a valid intensity (the tensor the reference ranks on, run.py:587, 511-515).  The
    ranking is stable (equally complete reads keep their order), like ``prep.downsample``.
    Returns (rows, kept_positions_mask or None when nothing was dropped)."""
    if rows.n_reads <= max_reads:
        return rows, None
    cov0, cov1 = coverage(rows)
    keep = (cov0 >= min_coverage) & (cov1 >= min_coverage) & ((cov0 + cov1) > 0)
    complete = ~np.isnan(rows.intensity) & ~np.isnan(rows.dwell)
    if motor_dwell_offset > 0:
        # run.py:575-584: with the offset beyond the transcript's end the motor column is all NaN
        motor = np.zeros_like(complete)
        end = rows.n_positions - motor_dwell_offset
        if end > 0:
            motor[:, :end] = ~np.isnan(rows.dwell[:, motor_dwell_offset:])
        complete &= motor
    score = complete[:, keep].sum(1)
    order = np.argsort(-score, kind='stable')
    selected = np.zeros(rows.n_reads, dtype=bool)
    for cond in (0, 1):
        chosen = order[rows.condition_ids[order] == cond][:max_reads]
        selected[chosen] = True
    return rows.take(selected), keep


def pack_rows(row_sets, motor_dwell_offset=0, pin=False, threads=None, wire='f32'):
    """Rows of one or more transcripts -> one CSR batch (pos_offsets, signal [E, V], label) with
    V = 3 when ``motor_dwell_offset`` > 0.  Returns (CsrBatch, pos_transcript, pos_index) like
    ``engine.pack_csr``.  Transcripts are packed concurrently (the native calls release the GIL).
    ``wire`` as in ``engine.pack_csr``: 'i16' / 'auto' write int16 entries (NCB_LAYOUT_CSR_I16) when
    every value is an int16 code, which Uncalled4 rows are (uncalled4.py:118-192)."""
    lib = L.load()
    threads = threads or _threads()
    V = 3 if motor_dwell_offset > 0 else 2
    ptrs = [r.row_pointers() for r in row_sets]
    labs = [r.labels() for r in row_sets]

    def count(k):
        r = row_sets[k]
        c = np.zeros(r.n_positions, dtype=np.int64)
        rc = lib.ncb_pack_rows_count(ptrs[k][0].ctypes.data, ptrs[k][1].ctypes.data, r.n_reads,
                                     r.n_positions, motor_dwell_offset, c.ctypes.data, 1)
        if rc != L.NCB_OK:
            raise L.NcbError(rc, "ncb_pack_rows_count")
        return c

    with ThreadPoolExecutor(max_workers=threads) as ex:
        counts = list(ex.map(count, range(len(row_sets))))
        all_counts = np.concatenate(counts) if counts else np.zeros(0, np.int64)
        offsets = np.zeros(all_counts.size + 1, dtype=np.int64)
        np.cumsum(all_counts, out=offsets[1:])
        E = int(offsets[-1])
        kw = dict(pin_memory=True) if pin else {}
        t_off = torch.from_numpy(offsets)
        if pin:
            t_off = t_off.pin_memory()
            offsets = t_off.numpy()
        if wire not in ('f32', 'i16', 'auto'):
            raise ValueError(f"wire must be 'f32', 'i16' or 'auto', not {wire!r}")
        t_lab = torch.empty((E,), dtype=torch.uint8, **kw)
        starts = np.concatenate([[0], np.cumsum([r.n_positions for r in row_sets])]).astype(np.int64)

        def fill_all(i16):
            t_sig = torch.empty((E, V), dtype=torch.int16 if i16 else torch.float32, **kw)
            fn = lib.ncb_pack_rows_fill_i16 if i16 else lib.ncb_pack_rows_fill

            def fill(k):
                r = row_sets[k]
                return fn(ptrs[k][0].ctypes.data, ptrs[k][1].ctypes.data,
                          labs[k].ctypes.data, r.n_reads, r.n_positions,
                          motor_dwell_offset, offsets[starts[k]:].ctypes.data,
                          t_sig.data_ptr(), t_lab.data_ptr(), 1)

            for rc in ex.map(fill, range(len(row_sets))):
                if rc != L.NCB_OK:
                    if i16 and rc == L.NCB_ERR_INVALID:
                        return None          # a value is not an int16 code
                    raise L.NcbError(rc, fn.__name__)
            return t_sig

        t_sig = fill_all(True) if wire in ('i16', 'auto') else None
        if t_sig is None:
            if wire == 'i16':
                raise ValueError("pack_rows(wire='i16'): the rows hold values that are not int16 codes")
            t_sig = fill_all(False)
    ptx = np.concatenate([np.full(r.n_positions, k, dtype=np.int32) for k, r in enumerate(row_sets)]) \
        if row_sets else np.zeros(0, np.int32)
    pidx = np.concatenate([np.arange(r.n_positions, dtype=np.int32) for r in row_sets]) \
        if row_sets else np.zeros(0, np.int32)
    return CsrBatch(t_off, t_sig, t_lab, int(all_counts.max()) if all_counts.size else 0), ptx, pidx


class TranscriptPipeline:
    """Read -> select -> pack -> compare for a list of transcripts at a time: the body of
    ``Worker.run`` between "task received" and "results frame" (run.py:420-463), with all
    transcripts of a call in ONE libncb batch.

    ``reader``: ``Uncalled4Reader`` or ``DbReader``; ``comparator``: this package's
    ``TranscriptComparator`` (it owns the engine and assembles the frames)."""

    def __init__(self, config, reader, comparator, device='cuda:0', log=None):
        self._conf = config
        self._reader = reader
        self._comparator = comparator
        self._device = device
        worker = getattr(comparator, '_worker', None)
        self._log = log or (worker.log if worker is not None else (lambda level, msg: None))
        self.skipped = []          # names of the transcripts whose files could not be read

    def _read_or_skip(self, transcript):
        """run.py:481-484: an error on one transcript is logged, the transcript is skipped and the
        worker carries on with the others."""
        try:
            return self.read(transcript)
        except MemoryError:
            raise
        except Exception as e:
            self._log("error", f"Encountered an error for {transcript.name}: {e}")
            self.skipped.append(transcript.name)
            return None

    def read(self, transcript):
        rows = self._reader.read(transcript.name, transcript.length)
        offset = self._conf.get_motor_dwell_offset()
        return select_reads(rows, self._conf.get_downsample_high_coverage(),
                            self._conf.get_min_coverage(), offset)

    def prepare(self, transcripts):
        """Host half of a batch: read, select, pack into one pinned CSR batch (no GPU work; the
        native calls release the GIL, so this overlaps the kernels of the batch before)."""
        offset = self._conf.get_motor_dwell_offset()
        live, row_sets, masks = [], [], []
        # a transcript keeps one thread per sample busy: with more cores than samples several
        # transcripts are decoded at once (the native calls release the GIL)
        n_samples = max(len(self._conf.get_sample_ids()), 1)
        workers = min(len(transcripts), max(_threads() // n_samples, 1))
        if workers > 1 and getattr(self._reader, 'concurrent_reads', False):
            with ThreadPoolExecutor(max_workers=workers) as ex:
                results = list(ex.map(self._read_or_skip, transcripts))
        else:
            results = [self._read_or_skip(t) for t in transcripts]
        for i, result in enumerate(results):
            if result is None:
                continue
            rows, keep = result
            if rows.n_reads == 0:
                continue
            live.append(i)
            row_sets.append(rows)
            masks.append(keep)
        # Uncalled4 rows are int16 tag values: they travel as 5-byte entries; anything else as float32
        wire = 'auto' if getattr(self._reader, 'int16_source', False) else 'f32'
        packed = pack_rows(row_sets, offset, pin=True, wire=wire) if live else None
        return transcripts, live, row_sets, masks, packed

    def finish(self, prepared):
        """Device half: one ``ncb_compare`` for the batch and the result frames."""
        transcripts, live, row_sets, masks, packed = prepared
        conf = self._conf
        min_cov = conf.get_min_coverage()
        out = [(t, None) for t in transcripts]
        if not live:
            return out
        batch, ptx, pidx = packed
        # a position is tested when it has a valid intensity (run.py:587) and enough coverage in
        # both conditions (run.py:511-515): the statistics kernel applies both (min_coverage >= 1);
        # after a downsample the reference keeps the positions chosen BEFORE it, so the host mask
        # decides there and the kernel's filter is off
        # min_coverage 0 keeps every position with a valid intensity in EITHER condition (run.py:587):
        # the kernel's per-condition filter cannot express that, so the host mask decides there too
        pre_filtered = any(m is not None for m in masks) or min_cov == 0
        res = self._comparator.compare_packed(batch, self._device,
                                              min_coverage=0 if pre_filtered else max(min_cov, 1))
        klen, _ = kit_geometry(conf.get_kit())
        # the positions of a transcript are one contiguous run of the batch
        starts = np.concatenate([[0], np.cumsum([r.n_positions for r in row_sets])]).astype(np.int64)
        keeps = []
        for k, i in enumerate(live):
            if pre_filtered:
                if masks[k] is not None:
                    keep = masks[k]
                else:
                    c0, c1 = coverage(row_sets[k])
                    keep = (c0 >= min_cov) & (c1 >= min_cov) & ((c0 + c1) > 0)
            else:
                keep = (res['status'][int(starts[k]):int(starts[k + 1])] & L.NCB_POS_LOW_COVERAGE) == 0
            keeps.append(keep)
        # one frame over the batch's rows, cut per transcript (comparisons.assemble_batch)
        frames = self._comparator.assemble_batch([transcripts[i] for i in live], starts, keeps, res, kmer_len=klen)
        for i, frame in zip(live, frames):
            out[i] = (transcripts[i], frame)
        return out

    def run(self, transcripts):
        """Returns ``[(transcript, DataFrame | None)]``; ``None`` = no reads or no tested position
        (run.py:428-456).  The frame carries the ``kmer`` column of run.py:461-463 when the
        transcript has a sequence."""
        return self.finish(self.prepare(transcripts))

    def run_stream(self, transcripts, batch_size=32):
        """Generator over ``run`` results for a long list of transcripts, ``batch_size`` at a time:
        a host thread reads and packs batch k + 1 while the GPU works on batch k and its frames are
        assembled (the task queue of run.py:104-131 and 380-420, as a two-stage pipeline)."""
        batches = [transcripts[i:i + batch_size] for i in range(0, len(transcripts), batch_size)]
        if not batches:
            return
        with ThreadPoolExecutor(max_workers=1) as ex:
            pending = ex.submit(self.prepare, batches[0])
            for nxt in batches[1:] + [None]:
                prepared = pending.result()
                pending = ex.submit(self.prepare, nxt) if nxt is not None else None
                yield from self.finish(prepared)
