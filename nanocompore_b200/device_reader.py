"""
Uncalled4 BAM files -> comparison results with the decoding on the GPU.

``readers.TranscriptPipeline`` decodes on the host: BGZF members inflated with zlib, tags turned into
planar float rows, rows packed into a pinned CSR batch that ``ncb_compare`` uploads -- about 6 ms of host
work per 2 000-position transcript at 400 reads next to 0.5 ms of kernels.  Here the host half of a batch
is a lookup in the BAM handles' indexes plus a copy of the COMPRESSED members into pinned memory
(``ncb_decoder_stage``); inflate, record walk, tag decoding (uncalled4.py:118-201), invalid-kmer filter
(run.py:711-718), read order (run.py:698-700) and CSR assembly are kernels (csrc/decode.cu) whose
output, a device-resident NCB_LAYOUT_CSR_I16 batch, goes straight into ``ncb_compare``.

The results are those of the host path to the bit (tests/test_decode_host.py on the CPU emulation,
tests/test_device_reader_gpu.py on the device).  What the kernels do not reproduce exactly they hand
back: a transcript with more records than ``downsample_high_coverage`` (the downsample of run.py:518-529
ranks reads on the host), ``min_coverage`` 0, and any transcript the decoder flags (malformed records,
values that are not int16 codes, overlapping ur segments, non-ASCII read names) go through
``TranscriptPipeline``'s host path, which also owns the error messages and the skip-and-log behaviour
of run.py:481-484.
"""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from nanocompore_b200 import _lib as L
from nanocompore_b200.engine import Results, make_labels
from nanocompore_b200.readers import TranscriptPipeline, Uncalled4Reader, _threads, encode_kmers, kit_geometry


class DeviceDecoder:
    """ncb_decoder_* over the BAM handles of an ``Uncalled4Reader`` (one decoder per GPU and process)."""

    def __init__(self, config, reader, device_index=0, n_slots=2):
        if not isinstance(reader, Uncalled4Reader):
            raise TypeError("the device-side decoder reads Uncalled4 BAM files")
        self._lib = L.load()
        self._conf = config
        self._reader = reader
        self.n_slots = n_slots
        files = config.get_sample_condition_bam_data()
        sids, cids, s2c = config.get_sample_ids(), config.get_condition_ids(), config.sample_to_condition()
        self._samples = [s for s, _, _ in files]
        self._bams = [reader._bams[s] for s in self._samples]
        self._handles = (C.c_void_p * len(files))(*[b._h for b in self._bams])
        self._labels = make_labels(np.array([sids[s] for s in self._samples]),
                                   np.array([cids[s2c[s]] for s in self._samples]))
        self._h = C.c_void_p()
        rc = self._lib.ncb_decoder_create(C.byref(self._h), device_index, n_slots)
        if rc != L.NCB_OK:
            raise L.NcbError(rc, "ncb_decoder_create: no usable GPU (there is no CPU fallback)")
        self._staged = [None] * n_slots

    def _check(self, rc):
        if rc == L.NCB_OK:
            return
        msg = self._lib.ncb_decoder_last_error(self._h).decode(errors='replace')
        if rc == L.NCB_ERR_NOMEM:
            raise torch.OutOfMemoryError(f"libncb: {msg}")
        raise L.NcbError(rc, msg)

    def records(self, transcript_name):
        """Records of the transcript over all files (an upper bound on its reads)."""
        return sum(b.count(transcript_name) for b in self._bams)

    def stage(self, slot, transcripts):
        """Host half: ``transcripts`` = [(name, length)].  No GPU work; releases the GIL."""
        refs = np.array([[self._lib.ncb_bam_reference_index(b._h, name.encode()) for b in self._bams]
                         for name, _ in transcripts], dtype=np.int32)
        lens = np.array([length for _, length in transcripts], dtype=np.int64)
        self._check(self._lib.ncb_decoder_stage(self._h, slot, len(self._bams), self._handles,
                                                self._labels.ctypes.data, len(transcripts), refs.ctypes.data,
                                                lens.ctypes.data, _threads()))
        self._staged[slot] = (len(transcripts), lens)

    def launch(self, slot, stream=None, motor_offset=0):
        """Enqueues upload + kernels; returns the ``ncb_batch`` of the slot (device pointers)."""
        klen, kcenter = kit_geometry(self._conf.get_kit())
        b = L.NcbBatch()
        s = C.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)
        self._check(self._lib.ncb_decoder_launch(self._h, slot, klen, kcenter, motor_offset,
                                                 float(self._conf.get_max_invalid_kmers_freq()), C.byref(b), s))
        return b

    def collect(self, slot):
        """Waits for the slot's kernels: (flags int32 [n_tx], kept reads int32 [n_tx], entries)."""
        n_tx, _ = self._staged[slot]
        flags = np.zeros(n_tx, dtype=np.int32)
        reads = np.zeros(n_tx, dtype=np.int32)
        entries = C.c_int64()
        self._check(self._lib.ncb_decoder_collect(self._h, slot, flags.ctypes.data, reads.ctypes.data, C.byref(entries)))
        return flags, reads, entries.value

    def staged_bytes(self, slot):
        text, members, records = C.c_int64(), C.c_int64(), C.c_int64()
        comp = self._lib.ncb_decoder_staged_bytes(self._h, slot, C.byref(text), C.byref(members), C.byref(records))
        return dict(compressed=int(comp), inflated=text.value, members=members.value, records=records.value)

    def kernel_launches(self, slot):
        return int(self._lib.ncb_decoder_kernel_launches(self._h, slot))

    def download(self, slot, n_vars=2):
        """The slot's CSR batch as numpy arrays (tests): (pos_offsets, signal [E, V], label)."""
        n_tx, lens = self._staged[slot]
        P = int(lens.sum())
        _, _, entries = self.collect(slot)
        off = np.zeros(P + 1, dtype=np.int64)
        sig = np.zeros((max(entries, 1), n_vars), dtype=np.int16)
        lab = np.zeros(max(entries, 1), dtype=np.uint8)
        self._check(self._lib.ncb_decoder_download(self._h, slot, off.ctypes.data, sig.ctypes.data, lab.ctypes.data,
                                                   max(entries, 1), n_vars))
        return off, sig[:entries], lab[:entries]

    def close(self):
        if self._h:
            self._lib.ncb_decoder_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceTranscriptPipeline(TranscriptPipeline):
    """``TranscriptPipeline`` with the Uncalled4 decoding on the GPU (module docstring).  Same interface:
    ``run`` / ``run_stream`` return ``[(transcript, DataFrame | None)]`` in the order given."""

    def __init__(self, config, reader, comparator, device='cuda:0', log=None, n_slots=3):
        super().__init__(config, reader, comparator, device, log)
        dev = torch.device(device)
        self._index = dev.index if dev.index is not None else 0
        # three slots: batch k + 1 is staged while batch k runs on the GPU and the frames of batch
        # k - 1 are read out of its result buffers (run_stream)
        self._decoder = DeviceDecoder(config, reader, self._index, max(int(n_slots), 3))
        self._next_slot = 0
        self._results = {}           # pinned result buffers, by slot
        self._stream = torch.cuda.Stream(device=dev)

    def close(self):
        self._decoder.close()

    def prepare(self, transcripts):
        conf = self._conf
        slot = self._next_slot
        self._next_slot = (slot + 1) % self._decoder.n_slots
        on_device, on_host = [], []
        if conf.get_min_coverage() < 1:
            on_host = list(range(len(transcripts)))
        else:
            cap = conf.get_downsample_high_coverage()
            for i, t in enumerate(transcripts):
                n = self._decoder.records(t.name)
                if n == 0:
                    continue                       # no reads anywhere: the frame is None
                (on_device if n <= cap else on_host).append(i)
        if on_device:
            self._decoder.stage(slot, [(transcripts[i].name, transcripts[i].length) for i in on_device])
        host_part = super().prepare([transcripts[i] for i in on_host]) if on_host else None
        return transcripts, slot, on_device, on_host, host_part

    def launch(self, prepared):
        """Enqueues the device half of a prepared batch -- upload of the compressed members, decode
        kernels, ``ncb_compare``, results to pinned memory -- on the pipeline's stream and returns
        without waiting."""
        transcripts, slot, on_device, on_host, host_part = prepared
        conf = self._conf
        view, done, starts = None, None, np.zeros(1, dtype=np.int64)
        if on_device:
            lens = np.array([transcripts[i].length for i in on_device], dtype=np.int64)
            starts = np.concatenate([[0], np.cumsum(lens)])
            P = int(starts[-1])
            if P > 0:
                engine = self._comparator._engine(self._device)
                buf = self._results.get(slot)
                if buf is None or buf.n_positions < P:
                    buf = Results(max(P, 1 << 16), engine.opts.n_samples, None)
                    self._results[slot] = buf
                view = buf.carve(P)
                engine.configure(min_coverage=max(conf.get_min_coverage(), 1), dwell_is_raw=True)
                try:
                    batch = self._decoder.launch(slot, self._stream, conf.get_motor_dwell_offset())
                    engine.compare_struct(batch, view, stream=self._stream, wait=False)
                finally:
                    engine.configure(min_coverage=0, dwell_is_raw=False)
                done = torch.cuda.Event()
                done.record(self._stream)
        return prepared, view, done, starts

    def collect(self, launched):
        """Waits for a launched batch and assembles its frames."""
        (transcripts, slot, on_device, on_host, host_part), view, done, starts = launched
        conf = self._conf
        out = [(t, None) for t in transcripts]
        redo = []
        if on_device and view is not None:
            done.synchronize()
            flags, _, _ = self._decoder.collect(slot)
            res = view.numpy()
            klen, _ = kit_geometry(conf.get_kit())
            ok = flags == L.NCB_DECODE_OK
            redo = [i for k, i in enumerate(on_device) if not ok[k]]
            # transcripts the decoder flagged keep no row here (the host path redoes them below)
            keeps = [((res['status'][int(starts[k]):int(starts[k + 1])] & L.NCB_POS_LOW_COVERAGE) == 0)
                     if ok[k] else np.zeros(int(starts[k + 1] - starts[k]), dtype=bool)
                     for k in range(len(on_device))]
            frames = self._comparator.assemble_batch([transcripts[i] for i in on_device], starts, keeps, res,
                                                     kmer_len=klen)
            for k, i in enumerate(on_device):
                if ok[k]:
                    out[i] = (transcripts[i], frames[k])
        if host_part is not None:
            for i, item in zip(on_host, super().finish(host_part)):
                out[i] = item
        # transcripts the decoder handed back: the host reader decides (result or logged error)
        for i in redo:
            out[i] = super().finish(super().prepare([transcripts[i]]))[0]
        return out

    def finish(self, prepared):
        return self.collect(self.launch(prepared))

    def run_stream(self, transcripts, batch_size=64):
        """``TranscriptPipeline.run_stream`` as a three-stage pipeline: a host thread stages the compressed
        members of batch k + 1, the GPU decodes and compares batch k, and this thread assembles the frames
        of batch k - 1 -- the GPU is never idle behind pandas.  Larger batches than the host path's: one
        BGZF member is one warp of the inflate kernel, and a batch of 64 transcripts x 400 reads is
        ~2 400 members for 148 SMs."""
        from concurrent.futures import ThreadPoolExecutor as Pool
        batches = [transcripts[i:i + batch_size] for i in range(0, len(transcripts), batch_size)]
        if not batches:
            return
        with Pool(max_workers=1) as ex:
            pending = ex.submit(self.prepare, batches[0])
            in_flight = None
            for nxt in batches[1:] + [None]:
                prepared = pending.result()
                pending = ex.submit(self.prepare, nxt) if nxt is not None else None
                launched = self.launch(prepared)
                if in_flight is not None:
                    yield from self.collect(in_flight)
                in_flight = launched
            yield from self.collect(in_flight)
