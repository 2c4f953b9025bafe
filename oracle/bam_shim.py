"""
Minimal stdlib BAM reader with the slice of the pysam API the reference's Uncalled4 parser
uses (``AlignmentFile.fetch / count``, ``read.get_tag / query_name / is_secondary /
is_supplementary``).  pysam is not installed in the build container; this shim lets
/root/reference/nanocompore/uncalled4.py run UNMODIFIED over the reference's fixture BAMs
when the golden vectors are generated (tests/golden/make_golden.py), and oracle/readers.py.  Test infrastructure.
"""
import gzip
import struct

_AUX_FMT = {'A': ('c', 1), 'c': ('b', 1), 'C': ('B', 1), 's': ('h', 2), 'S': ('H', 2),
            'i': ('i', 4), 'I': ('I', 4), 'f': ('f', 4)}


class Read:
    def __init__(self, query_name, flag, ref_id, pos, tags):
        self.query_name = query_name
        self.flag = flag
        self.reference_id = ref_id
        self.reference_start = pos
        self._tags = tags

    @property
    def is_secondary(self):
        return bool(self.flag & 256)

    @property
    def is_supplementary(self):
        return bool(self.flag & 2048)

    def get_tag(self, name):
        return self._tags[name]


def _parse_aux(buf, off, end):
    tags = {}
    while off < end:
        tag = buf[off:off + 2].decode()
        typ = chr(buf[off + 2])
        off += 3
        if typ in _AUX_FMT:
            fmt, size = _AUX_FMT[typ]
            val = struct.unpack_from('<' + fmt, buf, off)[0]
            off += size
        elif typ in 'ZH':
            z = buf.index(b'\0', off)
            val = buf[off:z].decode()
            off = z + 1
        elif typ == 'B':
            sub = chr(buf[off])
            n = struct.unpack_from('<I', buf, off + 1)[0]
            fmt, size = _AUX_FMT[sub]
            val = list(struct.unpack_from('<%d%s' % (n, fmt), buf, off + 5))
            off += 5 + n * size
        else:
            raise ValueError(f"unknown aux type {typ}")
        tags[tag] = val
    return tags


class AlignmentFile:
    def __init__(self, path, mode='rb'):
        with gzip.open(path, 'rb') as f:
            buf = f.read()
        assert buf[:4] == b'BAM\1'
        l_text = struct.unpack_from('<i', buf, 4)[0]
        off = 8 + l_text
        n_ref = struct.unpack_from('<i', buf, off)[0]
        off += 4
        self.references = []
        self.lengths = []
        for _ in range(n_ref):
            l_name = struct.unpack_from('<i', buf, off)[0]
            self.references.append(buf[off + 4:off + 4 + l_name - 1].decode())
            self.lengths.append(struct.unpack_from('<i', buf, off + 4 + l_name)[0])
            off += 8 + l_name
        self._reads = []
        while off < len(buf):
            block = struct.unpack_from('<i', buf, off)[0]
            start = off + 4
            (ref_id, pos, l_read_name, mapq, bin_, n_cigar, flag, l_seq, next_ref, next_pos,
             tlen) = struct.unpack_from('<iiBBHHHiiii', buf, start)
            p = start + 32
            name = buf[p:p + l_read_name - 1].decode()
            p += l_read_name + 4 * n_cigar + (l_seq + 1) // 2 + l_seq
            tags = _parse_aux(buf, p, start + block)
            self._reads.append(Read(name, flag, ref_id, pos, tags))
            off = start + block

    def fetch(self, reference=None):
        rid = self.references.index(reference)
        return [r for r in self._reads if r.reference_id == rid]

    def count(self, reference=None):
        return len(self.fetch(reference))

    def get_index_statistics(self):
        """pysam's per-contig index statistics (common.py:284-295 reads ``contig`` and ``mapped``):
        records placed on the contig whose unmapped flag is clear."""
        from collections import namedtuple
        Stat = namedtuple('IndexStats', 'contig mapped unmapped total')
        out = []
        for rid, name in enumerate(self.references):
            mapped = sum(1 for r in self._reads if r.reference_id == rid and not (r.flag & 4))
            unmapped = sum(1 for r in self._reads if r.reference_id == rid and (r.flag & 4))
            out.append(Stat(name, mapped, unmapped, mapped + unmapped))
        return out

    def close(self):
        pass
