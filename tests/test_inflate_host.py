"""
The DEFLATE decoder of the device-side BAM reader (nanocompore_b200/csrc/inflate_core.cuh) compiled for
the host: the same functions the CUDA warp runs (bit reader, dynamic header, canonical codes, lookup
tables, command queue, copy semantics), driven by a one-lane "warp".

  * against zlib on streams of every block kind (stored, fixed, dynamic), through ctypes;
  * the mutation harness tests/csrc/asan_inflate_harness.cpp under AddressSanitizer / UBSan: on damaged
    streams the decoder accepts exactly what zlib's inflate accepts (the host reader's rule, reader.cu)
    and never reads or writes outside its buffers.
"""
import ctypes as C
import os
import shutil
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'nanocompore_b200', 'csrc')
pytestmark = pytest.mark.skipif(shutil.which('g++') is None, reason="g++ is needed to build the host harness")


@pytest.fixture(scope='module')
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp('inflate') / 'inflate_host.so')
    subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-I', CSRC, os.path.join(ROOT, 'tests', 'csrc', 'inflate_host.cpp'),
                    '-o', so], check=True)
    lib = C.CDLL(so)
    lib.ncb_test_inflate.argtypes = [C.c_char_p, C.c_longlong, C.c_void_p, C.c_longlong]
    return lib


def _inflate(lib, payload, usize):
    out = np.zeros(max(usize, 1), dtype=np.uint8)
    rc = lib.ncb_test_inflate(payload, len(payload), out.ctypes.data, usize)
    return rc, bytes(out[:usize])


def _deflate(data, level, strategy):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return c.compress(data) + c.flush()


def test_tables_fit_a_warps_share_of_shared_memory(lib):
    # 8 warps per CTA, 6 CTAs per SM: 48 x sizeof(Tables) must stay below the 227 KB of an SM
    assert lib.ncb_test_inflate_tables_bytes() * 48 < 227 * 1024


@pytest.mark.parametrize('n', [1, 2, 100, 4000, 65000, 65536])
def test_every_block_kind_against_zlib(lib, n):
    rng = np.random.default_rng(n)
    cases = [rng.integers(0, 256, n, dtype=np.uint8).tobytes(),                       # stored blocks
             bytes(n),                                                                 # overlapping copies
             rng.integers(-3000, 3000, n // 2 + 1).astype('<i2').tobytes()[:n],        # signal-like
             (b'abcabcabd' * (n // 9 + 1))[:n], rng.integers(0, 4, n, dtype=np.uint8).tobytes()]
    for data in cases:
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                rc, got = _inflate(lib, _deflate(data, level, strategy), len(data))
                assert rc == 0 and got == data, (n, level, strategy, rc)


def test_wrong_isize_and_truncation_are_errors(lib):
    data = np.random.default_rng(3).integers(-3000, 3000, 3000).astype('<i2').tobytes()
    comp = _deflate(data, 6, zlib.Z_DEFAULT_STRATEGY)
    assert _inflate(lib, comp, len(data))[0] == 0
    assert _inflate(lib, comp, len(data) - 1)[0] != 0
    assert _inflate(lib, comp, len(data) + 1)[0] != 0
    assert _inflate(lib, comp[:len(comp) // 2], len(data))[0] != 0
    assert _inflate(lib, b'', len(data))[0] != 0
    assert _inflate(lib, b'\x07', 10)[0] != 0          # block type 3


def test_multi_block_streams(lib):
    """Several deflate blocks in one member (Z_FULL_FLUSH between them: stored empty blocks included)."""
    rng = np.random.default_rng(5)
    parts = [rng.integers(-2000, 2000, 1500).astype('<i2').tobytes() for _ in range(5)]
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = b''.join(c.compress(p) + c.flush(zlib.Z_FULL_FLUSH) for p in parts) + c.flush()
    rc, got = _inflate(lib, comp, sum(map(len, parts)))
    assert rc == 0 and got == b''.join(parts)


def test_mutation_harness_agrees_with_zlib_under_sanitizers(tmp_path):
    exe = str(tmp_path / 'inflate_harness')
    build = subprocess.run(['g++', '-std=c++17', '-O1', '-g', '-fsanitize=address,undefined',
                            '-fno-sanitize-recover=undefined', '-I', CSRC,
                            os.path.join(ROOT, 'tests', 'csrc', 'asan_inflate_harness.cpp'), '-lz', '-o', exe],
                           capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("sanitizer build not available here: " + build.stderr[-300:])
    run = subprocess.run([exe, '500'], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    assert 'all agree' in run.stdout
