"""
Mutation test of the native BAM reader (libncb: ncb_bam_open / ncb_bam_read_uncalled4): a small valid
Uncalled4 BAM is damaged -- bytes, bits and whole length fields of its uncompressed stream, now and
then a byte of the compressed file -- and read back.  Every outcome must be a result or an NcbError
with a message; a crash (wrong bounds check, exception across the C ABI) ends the process, which is
what tests/test_readers.py::test_damaged_bams_never_crash_the_reader looks for.

    python tests/fuzz_bam.py [seed] [files] [--keep DIR]

``--keep DIR`` writes the damaged files to DIR and does not read them (input of the sanitizer harness
tests/csrc/asan_bam_harness.cpp).
"""
import os
import struct
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import signal_files as SF  # noqa: E402
from nanocompore_b200 import _lib as L  # noqa: E402
from nanocompore_b200 import readers as R  # noqa: E402

REFS = [('tx', 120), ('ty', 80)]
NASTY = [-1, -2 ** 31, 2 ** 31 - 1, 0, 65536, 1 << 20]


def valid_stream(rng):
    """The uncompressed byte stream of a BAM with 9 Uncalled4 reads on two references."""
    reads = [SF.make_uncalled4_read(rng, 120, 'RNA002', f'r{k}', ref=0) for k in range(6)] + \
            [SF.make_uncalled4_read(rng, 80, 'RNA002', f's{k}', ref=1) for k in range(3)]
    text = b'@HD\tVN:1.6\tSO:coordinate\n'
    out = b'BAM\1' + struct.pack('<i', len(text)) + text + struct.pack('<i', len(REFS))
    for name, length in REFS:
        out += struct.pack('<i', len(name) + 1) + name.encode() + b'\0' + struct.pack('<i', length)
    for r in reads:
        tags = SF._aux_array('ur', 'i', r['ur']) + SF._aux_array('uc', 's', r['uc']) + \
            SF._aux_array('ul', 's', r['ul'])
        out += SF._record(r['ref'], 0, r['name'], 0, tags)
    return out


def damaged(rng, stream):
    b = bytearray(stream)
    mode = int(rng.integers(0, 3))
    for _ in range(int(rng.integers(1, 4))):
        i = int(rng.integers(4, len(b)))
        if mode == 0:
            b[i] = int(rng.integers(0, 256))                                  # a byte
        elif mode == 1:
            b[i:i + 4] = struct.pack('<i', int(rng.choice(NASTY)))            # a length-like field
        else:
            b[i] ^= 1 << int(rng.integers(0, 8))                              # a bit
    return bytes(b)


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n_files = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    keep = sys.argv[sys.argv.index('--keep') + 1] if '--keep' in sys.argv else None
    rng = np.random.default_rng(seed)
    tmp = keep or tempfile.mkdtemp()
    os.makedirs(tmp, exist_ok=True)
    stream = valid_stream(rng)
    survived = errors = 0
    for it in range(n_files):
        payload = damaged(rng, stream)
        path = os.path.join(tmp, f'f{it}.bam')
        with open(path, 'wb') as f:
            for i in range(0, len(payload), 700):
                f.write(SF._bgzf_block(payload[i:i + 700]))
            f.write(SF._bgzf_block(b''))
        if rng.random() < 0.2:              # the compressed file itself
            raw = bytearray(open(path, 'rb').read())
            raw[int(rng.integers(0, len(raw)))] = int(rng.integers(0, 256))
            open(path, 'wb').write(raw)
        if keep:
            continue
        try:
            bam = R.NativeBam(path, threads=2)
            for name, length in REFS:
                bam.read_uncalled4(name, length, 'RNA002')
            bam.close()
            survived += 1
        except L.NcbError:
            errors += 1
        os.remove(path)
    print('survived', survived, 'errors', errors)


if __name__ == '__main__':
    main()
