"""
The device-side Uncalled4 BAM decoder (nanocompore_b200/csrc/decode.cu) without a GPU: its staging code and
the per-unit functions its kernels call (decode_stage.h, inflate_core.cuh, decode_core.cuh), run by the host
emulation tests/csrc/decode_host.cpp, against the host path of the product -- ``Uncalled4Reader.read`` +
``pack_rows(wire='i16')`` -- on synthetic BAMs (random ur segments, -32768 currents, zero and negative
lengths, secondary alignments, records straddling BGZF members, unsorted files, both kits) and on the
reference's fixture BAMs.  The CSR arrays must be identical: offsets, int16 entries, labels.
"""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import signal_files as SF  # noqa: E402
from nanocompore_b200 import readers as R  # noqa: E402
from test_readers import REF_FIXTURES, bam_config, needs_reference, synthetic_bams  # noqa: E402

CSRC = os.path.join(ROOT, 'nanocompore_b200', 'csrc')
pytestmark = pytest.mark.skipif(shutil.which('g++') is None, reason="g++ is needed to build the host emulation")


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp('decode') / 'decode_host.so')
    subprocess.run(['g++', '-x', 'c++', '-std=c++17', '-O2', '-shared', '-fPIC', '-I', os.path.join(ROOT, 'include'),
                    '-I', CSRC, os.path.join(CSRC, 'reader.cu'), os.path.join(ROOT, 'tests', 'csrc', 'decode_host.cpp'),
                    '-lz', '-lpthread', '-o', so], check=True)
    lib = C.CDLL(so)
    lib.ncb_bam_open.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
    lib.ncb_bam_close.argtypes = [C.c_void_p]
    lib.ncb_bam_reference_index.argtypes = [C.c_void_p, C.c_char_p]
    lib.ncb_test_decode.restype = C.c_longlong
    lib.ncb_test_decode.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                                    C.c_void_p, C.c_void_p]
    return lib


def emulate(lib, cfg, transcripts, motor_offset=0):
    """transcripts: [(name, length)].  Returns (pos_offsets, signal [E, V], label, flags, kept)."""
    files = cfg.get_sample_condition_bam_data()
    sids, cids, s2c = cfg.get_sample_ids(), cfg.get_condition_ids(), cfg.sample_to_condition()
    handles = (C.c_void_p * len(files))()
    for i, (_, _, path) in enumerate(files):
        h = C.c_void_p()
        assert lib.ncb_bam_open(C.byref(h), path.encode(), 2) == 0
        handles[i] = h
    labels = np.array([(sids[s] << 1) | cids[s2c[s]] for s, _, _ in files], dtype=np.uint8)
    refs = np.array([[lib.ncb_bam_reference_index(handles[i], name.encode()) for i in range(len(files))]
                     for name, _ in transcripts], dtype=np.int32)
    lens = np.array([length for _, length in transcripts], dtype=np.int64)
    klen, kcenter = R.kit_geometry(cfg.get_kit())
    P = int(lens.sum())
    V = 3 if motor_offset > 0 else 2
    cap = 4 * 64 * P + 16
    off = np.zeros(P + 1, dtype=np.int64)
    sig = np.zeros((cap, V), dtype=np.int16)
    lab = np.zeros(cap, dtype=np.uint8)
    flags = np.zeros(len(transcripts), dtype=np.int32)
    kept = np.zeros(len(transcripts), dtype=np.int64)
    E = lib.ncb_test_decode(len(files), handles, labels.ctypes.data, len(transcripts), refs.ctypes.data, lens.ctypes.data,
                            klen, kcenter, motor_offset, cfg.get_max_invalid_kmers_freq(), off.ctypes.data, sig.ctypes.data,
                            lab.ctypes.data, cap, flags.ctypes.data, kept.ctypes.data)
    for h in handles:
        lib.ncb_bam_close(h)
    assert E >= 0, E
    return off, sig[:E], lab[:E], flags, kept


def host_path(cfg, transcripts, motor_offset=0):
    reader = R.Uncalled4Reader(cfg, threads=2)
    rows = [reader.read(name, length) for name, length in transcripts]
    batch, _, _ = R.pack_rows(rows, motor_offset, wire='i16')
    reader.close()
    return batch.pos_offsets.numpy(), batch.signal.numpy(), batch.label.numpy(), [r.n_reads for r in rows]


@pytest.mark.parametrize('kit,unsorted,motor', [('RNA002', False, 0), ('RNA004', False, 0), ('RNA004', True, 0),
                                                ('RNA002', False, 7)])
def test_emulated_device_decoder_equals_the_host_path(emu, tmp_path, kit, unsorted, motor):
    refs, paths = synthetic_bams(tmp_path, kit, 21 + unsorted + motor, n_reads=60, unsorted=unsorted)
    cfg = bam_config(paths, kit=kit, max_invalid_kmers_freq=0.08)
    off, sig, lab, flags, kept = emulate(emu, cfg, refs, motor)
    want_off, want_sig, want_lab, want_reads = host_path(cfg, refs, motor)
    assert (flags == 0).all()
    assert list(kept) == want_reads and kept[2] == 0 and 0 < kept[0] < 4 * 63
    assert np.array_equal(off, want_off)
    assert np.array_equal(sig, want_sig)
    assert np.array_equal(lab, want_lab)


@needs_reference
def test_reference_fixture_bams(emu):
    paths = [os.path.join(REF_FIXTURES, f) for f in ('kd1.bam', 'kd2.bam', 'wt1.bam', 'wt2.bam')]
    if not all(os.path.exists(p) for p in paths):
        pytest.skip("fixture BAMs under other names")
    cfg = bam_config(paths)
    reader = R.Uncalled4Reader(cfg, threads=2)
    names = sorted(reader.references())
    lengths = {n: l for bam in reader._bams.values() for n, l in zip(bam.references, bam.lengths)}
    reader.close()
    transcripts = [(n, lengths[n]) for n in names]
    off, sig, lab, flags, kept = emulate(emu, cfg, transcripts)
    want_off, want_sig, want_lab, want_reads = host_path(cfg, transcripts)
    assert (flags == 0).all() and list(kept) == want_reads
    assert np.array_equal(off, want_off) and np.array_equal(sig, want_sig) and np.array_equal(lab, want_lab)


def test_what_the_device_path_hands_back_to_the_host_reader(emu, tmp_path):
    """Values that are no int16 codes and malformed tags flag the transcript, the others stay exact."""
    rng = np.random.default_rng(5)
    refs = [('ok', 150), ('wide', 150), ('broken', 150)]
    paths = []
    for sample in ('kd1', 'kd2', 'wt1', 'wt2'):
        reads = []
        for ri in range(3):
            for k in range(8):
                r = SF.make_uncalled4_read(rng, 150, 'RNA002', f'{sample}-{ri}-{k}', ref=ri, n_segments=1)
                if ri == 1 and k == 3:
                    r['ul_type'] = 'i'
                    r['ul'][5] = 70000                     # a dwell that is no int16 code
                if ri == 2 and k == 2:
                    r['ur'] = r['ur'][:1]                  # odd ur tag
                reads.append(r)
        p = str(tmp_path / f'{sample}.bam')
        SF.write_bam(p, refs, reads, block_bytes=3000)
        paths.append(p)
    cfg = bam_config(paths, max_invalid_kmers_freq=0.5)
    off, sig, lab, flags, kept = emulate(emu, cfg, refs)
    assert flags[0] == 0 and flags[1] == 2 and flags[2] == 1
    want_off, want_sig, want_lab, _ = host_path(cfg, refs[:1])
    assert np.array_equal(off[:151], want_off) and np.array_equal(sig[:off[150]], want_sig)
    assert off[150] == off[-1]                              # flagged transcripts contribute no entries
