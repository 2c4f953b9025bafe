"""
Writers of small synthetic signal files for the reader tests: Uncalled4-style BAMs (BGZF, the
``ur/uc/ul`` tags of uncalled4.py:118-147) and the SQLite stores of the reference's preprocessing
(database.py:40-76).  Test infrastructure; stdlib + numpy only.
"""
import sqlite3
import struct
import zlib

import numpy as np

KITS = {'RNA002': (5, 4), 'RNA004': (9, 5)}


def _bgzf_block(payload):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(payload) + comp.flush()
    bsize = 12 + 6 + len(body) + 8 - 1
    assert bsize < 65536
    head = struct.pack('<BBBBIBBH', 31, 139, 8, 4, 0, 0, 255, 6) + b'BC' + struct.pack('<HH', 2, bsize)
    return head + body + struct.pack('<II', zlib.crc32(payload) & 0xffffffff, len(payload))


def _aux_array(tag, sub, values):
    fmt = {'c': 'b', 'C': 'B', 's': 'h', 'S': 'H', 'i': 'i', 'I': 'I'}[sub]
    return tag.encode() + b'B' + sub.encode() + struct.pack('<I', len(values)) + \
        struct.pack('<%d%s' % (len(values), fmt), *[int(v) for v in values])


def _record(ref_id, pos, name, flag, tags):
    l_seq = 4
    core = struct.pack('<iiBBHHHiiii', ref_id, pos, len(name) + 1, 60, 4680, 1, flag, l_seq, -1, -1, 0)
    body = core + name.encode() + b'\0' + struct.pack('<I', (l_seq << 4) | 0) + b'\x12\x48' + b'\xff' * l_seq
    # a few unrelated aux fields around the signal tags, of every scalar/string/array kind
    body += b'NMC' + struct.pack('<B', 3) + b'RGZ' + b'grp\0' + b'xfB' + b'f' + struct.pack('<I', 2) + struct.pack('<2f', 1.5, 2.5)
    body += tags
    body += b'asi' + struct.pack('<i', -7)
    return struct.pack('<i', len(body)) + body


def write_bam(path, references, reads, block_bytes=4000):
    """``references``: [(name, length)]; ``reads``: [dict(ref=index, name=, flag=, ur=, uc=, ul=,
    pos=0, ur_type='i', ul_type='s')] in file order.  Small BGZF blocks (records straddle block
    boundaries) and the standard empty EOF block."""
    text = b'@HD\tVN:1.6\tSO:coordinate\n'
    out = b'BAM\1' + struct.pack('<i', len(text)) + text + struct.pack('<i', len(references))
    for name, length in references:
        out += struct.pack('<i', len(name) + 1) + name.encode() + b'\0' + struct.pack('<i', length)
    for r in reads:
        tags = _aux_array('ur', r.get('ur_type', 'i'), r['ur']) + _aux_array('uc', 's', r['uc']) + \
            _aux_array('ul', r.get('ul_type', 's'), r['ul'])
        out += _record(r['ref'], r.get('pos', 0), r['name'], r.get('flag', 0), tags)
    with open(path, 'wb') as f:
        for i in range(0, len(out), block_bytes):
            f.write(_bgzf_block(out[i:i + block_bytes]))
        f.write(_bgzf_block(b''))


def make_uncalled4_read(rng, ref_len, kit, name, ref=0, flag=0, n_segments=None):
    """A read with random ``ur`` segments and matching ``uc`` / ``ul`` arrays: -32768 currents
    (Uncalled4's NaN), zero lengths (skips) and negative padding lengths included."""
    klen, kcenter = KITS[kit]
    shift = klen - kcenter
    n_segments = n_segments or int(rng.integers(1, 4))
    # segment boundaries in tensor coordinates, then mapped back to ur values
    lo = int(rng.integers(0, max(1, ref_len // 4)))
    hi = int(rng.integers(ref_len - ref_len // 4, ref_len + 1))
    cuts = np.sort(rng.choice(np.arange(lo + 2, hi - 2), size=2 * (n_segments - 1), replace=False)) \
        if n_segments > 1 else np.array([], dtype=np.int64)
    bounds = [lo] + [int(c) for c in cuts] + [hi]
    ur, sizes = [], []
    for k in range(n_segments):
        a, b = bounds[2 * k], bounds[2 * k + 1]
        sizes.append(b - a)
        # uncalled4.py:168-181: first start maps to itself, last end is shifted by klen - 1,
        # the inner boundaries by klen - kcenter
        ur.append(a if k == 0 else a + shift)
        ur.append(b + klen - 1 if k == n_segments - 1 else b + shift)
    total = sum(sizes)
    uc = rng.integers(-12000, 12000, size=total)
    uc[rng.random(total) < 0.03] = -32768
    ul = rng.integers(1, 120, size=total)
    ul[rng.random(total) < 0.1] = 0
    ul[0] = max(int(ul[0]), 1)
    ul_tag = []
    for v in ul:       # negative paddings anywhere: the parser drops them
        if rng.random() < 0.02:
            ul_tag.append(-1)
        ul_tag.append(int(v))
    # both arrays are stored 3' -> 5'
    return dict(ref=ref, name=name, flag=flag, ur=ur, uc=uc[::-1].tolist(), ul=ul_tag[::-1])


def write_db(path, transcripts, reads, signal):
    """``transcripts``: [(name, id)]; ``reads``: [(read name, id, invalid_kmers)]; ``signal``:
    [(transcript id, read id, intensity float32[P], dwell float32[P])]."""
    conn = sqlite3.connect(path)
    conn.executescript("""
        CREATE TABLE transcripts (name TEXT NOT NULL, id INTEGER NOT NULL, PRIMARY KEY(name));
        CREATE TABLE reads (read TEXT NOT NULL, id INTEGER NOT NULL, invalid_kmers REAL);
        CREATE TABLE signal_data (transcript_id INTEGER NOT NULL, read_id INTEGER NOT NULL,
                                  intensity BLOB, dwell BLOB);
        CREATE INDEX reads_id_index ON reads (id);
        CREATE INDEX signal_data_index ON signal_data (transcript_id, read_id);
    """)
    conn.executemany("INSERT INTO transcripts VALUES (?, ?)", transcripts)
    conn.executemany("INSERT INTO reads VALUES (?, ?, ?)", reads)
    conn.executemany("INSERT INTO signal_data VALUES (?, ?, ?, ?)",
                     [(t, r, np.asarray(i, np.float32).tobytes(), np.asarray(d, np.float32).tobytes())
                      for t, r, i, d in signal])
    conn.commit()
    conn.close()
