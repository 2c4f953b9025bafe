"""
Transcript readers: libncb's native BGZF/BAM + Uncalled4 decoder, the SQLite
reader and the rows -> CSR packer against (i) the tensors the UNMODIFIED reference parser produced
from its fixture BAMs (tests/golden/fixture_cfg1.npz), (ii) the known answers of the reference's
tests/test_uncalled4.py and tests/test_run.py:241-308, (iii) the numpy restatement
oracle/readers.py on synthetic files.  No GPU needed: readers and packer are host code.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import signal_files as SF  # noqa: E402
from nanocompore_b200.config import PathConfig  # noqa: E402
from nanocompore_b200.engine import pack_csr_numpy  # noqa: E402
from nanocompore_b200 import readers as R  # noqa: E402
from oracle import readers as OR  # noqa: E402

REF_FIXTURES = '/root/reference/tests/fixtures'
needs_reference = pytest.mark.skipif(not os.path.isdir(REF_FIXTURES),
                                     reason="the reference's fixture files are only in the build container")
GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'fixture_cfg1.npz')


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and \
        np.array_equal(np.nan_to_num(a), np.nan_to_num(b))


def bam_config(paths, kit='RNA002', **extra):
    return PathConfig(dict(data={'kd': {'kd1': {'bam': paths[0]}, 'kd2': {'bam': paths[1]}},
                                 'wt': {'wt1': {'bam': paths[2]}, 'wt2': {'bam': paths[3]}}},
                           depleted_condition='kd', kit=kit, **extra))


def ids_of(cfg):
    s2c, cids = cfg.sample_to_condition(), cfg.get_condition_ids()
    return cfg.get_sample_ids(), {s: cids[s2c[s]] for s in cfg.get_sample_ids()}


# ---------------------------------------------------------------------------------------------
# the reference's own fixtures and known answers
# ---------------------------------------------------------------------------------------------

@needs_reference
def test_fixture_bams_match_reference_parser():
    """Native reader and oracle restatement both reproduce, bit for bit, the tensors the unmodified
    reference parser gave for its 4 fixture BAMs (config 1), read order and ids included."""
    paths = [os.path.join(REF_FIXTURES, f'{s}.bam') for s in ('kd1', 'kd2', 'wt1', 'wt2')]
    cfg = bam_config(paths, min_coverage=1)
    reader = R.Uncalled4Reader(cfg)
    g = np.load(GOLDEN)
    sids, cids = ids_of(cfg)
    for ti in range(int(g['n_transcripts'])):
        name, raw = str(g[f'tx{ti}/name']), g[f'tx{ti}/raw']
        rows = reader.read(name, raw.shape[0])
        assert same(rows.dense(), raw)
        assert np.array_equal(rows.sample_ids, g[f'tx{ti}/sample_ids'])
        assert np.array_equal(rows.condition_ids, g[f'tx{ti}/condition_ids'])
        data, s, c = OR.read_data_uncalled4(name, raw.shape[0], dict(zip(('kd1', 'kd2', 'wt1', 'wt2'), paths)),
                                            sids, cids, 'RNA002')
        assert same(data, raw) and np.array_equal(s, rows.sample_ids) and np.array_equal(c, rows.condition_ids)
    reader.close()


@needs_reference
def test_reference_known_answers_uncalled4():
    """tests/test_uncalled4.py:12-100 (RNA004, two signal segments) and :103-134 (empty BAM)."""
    bam = R.NativeBam(os.path.join(REF_FIXTURES, 'uncalled4_sample.bam'))
    ref = bam.references[0]
    inten, dwell, names = bam.read_uncalled4(ref, 1781, 'RNA004')
    assert inten.shape == (1, 1781)
    t = inten[0]
    assert t[1770] == 7992 and np.isnan(t[1771])
    assert t[604] == -5072 and np.isnan(t[600:604]).all()
    assert np.isnan(t[554:558]).all() and t[553] == -10820
    assert t[6] == 3802 and np.isnan(t[5])
    empty = R.NativeBam(os.path.join(REF_FIXTURES, 'empty.bam'))
    inten, dwell, names = empty.read_uncalled4(ref, 1781, 'RNA004')
    assert inten.shape == (0, 1781) and names == []


@needs_reference
def test_reference_known_answers_read_data():
    """tests/test_run.py:241-266 (Uncalled4) and :269-308 (eventalign SQLite)."""
    ref_id = 'ENST00000674681.1|ENSG00000075624.17|OTTHUMG00000023268|-|ACTB-219|ACTB|2554|protein_coding|'
    paths = [os.path.join(REF_FIXTURES, f'{s}.bam') for s in ('kd1', 'kd2', 'wt1', 'wt2')]
    rows = R.Uncalled4Reader(bam_config(paths)).read(ref_id, 2554)
    assert rows.dense().shape == (2554, 12, 2)
    assert not np.isnan(rows.dense()[301:303]).any()
    assert np.array_equal(rows.sample_ids, [3, 3, 0, 1, 3, 2, 0, 1, 2, 1, 0, 2])
    assert np.array_equal(rows.condition_ids, [1, 1, 0, 0, 1, 1, 0, 0, 1, 0, 0, 1])

    dbs = {s: os.path.join(REF_FIXTURES, f'{s}_eventalign.db') for s in ('kd1', 'kd2', 'wt1', 'wt2')}
    cfg = PathConfig(dict(data={'kd': {'kd1': {'db': dbs['kd1']}, 'kd2': {'db': dbs['kd2']}},
                                'wt': {'wt1': {'db': dbs['wt1']}, 'wt2': {'db': dbs['wt2']}}},
                          depleted_condition='kd'))
    rows = R.DbReader(cfg).read(ref_id, 2554)
    assert rows.dense().shape == (2554, 4, 2)
    assert np.array_equal(rows.sample_ids, [0, 0, 1, 2])
    assert np.array_equal(rows.condition_ids, [0, 0, 0, 1])
    sids, cids = ids_of(cfg)
    data, s, c = OR.read_data_db(ref_id, dbs, sids, cids)
    assert same(data, rows.dense()) and np.array_equal(s, rows.sample_ids) and np.array_equal(c, rows.condition_ids)


# ---------------------------------------------------------------------------------------------
# synthetic files: native reader vs the restatement
# ---------------------------------------------------------------------------------------------

def synthetic_bams(tmp_path, kit, seed, n_reads=40, unsorted=False):
    rng = np.random.default_rng(seed)
    refs = [('txA|gene|1', 700), ('txB|gene|2', 333), ('txC|gene|3', 90)]
    paths = []
    for si, sample in enumerate(('kd1', 'kd2', 'wt1', 'wt2')):
        reads = []
        for ri, (name, length) in enumerate(refs[:2]):     # txC has no reads at all
            for k in range(n_reads + si):
                flag = 0 if rng.random() > 0.15 else int(rng.choice([256, 2048, 16]))
                reads.append(SF.make_uncalled4_read(rng, length, kit, f'read-{rng.integers(1 << 30):08x}-{sample}-{k}',
                                                    ref=ri, flag=flag))
        if unsorted:
            rng.shuffle(reads)
        # integer array subtypes vary between writers
        for r in reads[::3]:
            r['ur_type'] = 'I'
        path = str(tmp_path / f'{sample}_{kit}_{seed}.bam')
        SF.write_bam(path, refs, reads, block_bytes=1500 + 977 * si)
        paths.append(path)
    return refs, paths


@pytest.mark.parametrize('kit,unsorted', [('RNA002', False), ('RNA004', False), ('RNA004', True)])
def test_native_bam_reader_matches_restatement(tmp_path, kit, unsorted):
    refs, paths = synthetic_bams(tmp_path, kit, 7 + unsorted, unsorted=unsorted)
    cfg = bam_config(paths, kit=kit, max_invalid_kmers_freq=0.08)
    reader = R.Uncalled4Reader(cfg, threads=3)
    sids, cids = ids_of(cfg)
    for name, length in refs:
        rows = reader.read(name, length)
        data, s, c = OR.read_data_uncalled4(name, length, dict(zip(('kd1', 'kd2', 'wt1', 'wt2'), paths)),
                                            sids, cids, kit, 0.08)
        assert rows.n_positions == length
        assert same(rows.dense(), data.reshape(length, -1, 2))
        assert np.array_equal(rows.sample_ids, s) and np.array_equal(rows.condition_ids, c)
    assert set(reader.references()) == {'txA|gene|1', 'txB|gene|2'}
    # some reads must have been dropped by the invalid-kmer filter and by their flags
    assert 0 < reader.read(refs[0][0], refs[0][1]).n_reads < 4 * 43
    reader.close()


def test_bam_errors(tmp_path):
    from nanocompore_b200 import _lib as L
    p = tmp_path / 'not_a.bam'
    p.write_bytes(b'hello world, this is not gzip' * 10)
    with pytest.raises(L.NcbError):
        R.NativeBam(str(p))
    with pytest.raises(L.NcbError):
        R.NativeBam(str(tmp_path / 'missing.bam'))
    # a read whose signal does not fill its ur segments: the reference's numpy assignment raises
    rng = np.random.default_rng(3)
    read = SF.make_uncalled4_read(rng, 200, 'RNA002', 'short')
    read['uc'] = read['uc'][:-5]
    path = str(tmp_path / 'bad.bam')
    SF.write_bam(path, [('tx', 200)], [read])
    with pytest.raises(L.NcbError, match='does not fit'):
        R.NativeBam(path).read_uncalled4('tx', 200, 'RNA002')
    # a reference absent from the header has no reads
    assert R.NativeBam(path).read_uncalled4('other', 10, 'RNA002')[0].shape == (0, 10)


def test_bam_edge_files(tmp_path):
    """Header-only files, truncation in the middle of a block, the same read id in two samples."""
    from nanocompore_b200 import _lib as L
    rng = np.random.default_rng(12)
    refs = [('tx', 120)]
    empty = str(tmp_path / 'header_only.bam')
    SF.write_bam(empty, refs, [])
    bam = R.NativeBam(empty)
    assert bam.references == ['tx'] and bam.count('tx') == 0 and bam.mapped_references() == {}
    assert bam.read_uncalled4('tx', 120, 'RNA002')[0].shape == (0, 120)

    reads = [SF.make_uncalled4_read(rng, 120, 'RNA002', f'r{k}') for k in range(30)]
    full = str(tmp_path / 'full.bam')
    SF.write_bam(full, refs, reads, block_bytes=900)
    raw = open(full, 'rb').read()
    cut = str(tmp_path / 'cut.bam')
    open(cut, 'wb').write(raw[:len(raw) // 2])
    with pytest.raises(L.NcbError):
        R.NativeBam(cut)

    # one read id present in two samples: both copies are kept, in the config order of the samples
    paths = []
    for s in ('kd1', 'kd2', 'wt1', 'wt2'):
        p = str(tmp_path / f'{s}_dup.bam')
        SF.write_bam(p, refs, [SF.make_uncalled4_read(rng, 120, 'RNA002', 'same-id', n_segments=1),
                               SF.make_uncalled4_read(rng, 120, 'RNA002', f'only-{s}', n_segments=1)])
        paths.append(p)
    rows = R.Uncalled4Reader(bam_config(paths, max_invalid_kmers_freq=1.1)).read('tx', 120)
    assert list(rows.read_names) == ['only-kd1', 'only-kd2', 'only-wt1', 'only-wt2'] + ['same-id'] * 4
    assert list(rows.sample_ids) == [0, 1, 2, 3, 0, 1, 2, 3]


def synthetic_dbs(tmp_path, seed, P=150):
    rng = np.random.default_rng(seed)
    paths = {}
    for si, sample in enumerate(('kd1', 'kd2', 'wt1', 'wt2')):
        n = [9, 0, 14, 6][si]
        reads = [(f'r{sample}{k}', k, float(rng.choice([0.0, 0.05, 0.1, 0.3]))) for k in range(n)]
        signal = []
        for tid, tname in ((1, 'tx1'), (2, 'tx2')):
            for k in range(n):
                inten = rng.normal(100, 5, P).astype(np.float32)
                dwell = rng.lognormal(-4.6, 0.5, P).astype(np.float32)
                gap = rng.random(P) < 0.3
                inten[gap] = np.nan
                dwell[gap & (rng.random(P) < 0.9)] = np.nan
                signal.append((tid, k, inten, dwell))
        path = str(tmp_path / f'{sample}_{seed}.db')
        SF.write_db(path, [('tx1', 1), ('tx2', 2)], reads, signal)
        paths[sample] = path
    return paths


def test_db_reader_matches_restatement(tmp_path):
    dbs = synthetic_dbs(tmp_path, 11)
    cfg = PathConfig(dict(data={'kd': {'kd1': {'db': dbs['kd1']}, 'kd2': {'db': dbs['kd2']}},
                                'wt': {'wt1': {'db': dbs['wt1']}, 'wt2': {'db': dbs['wt2']}}},
                          depleted_condition='kd', max_invalid_kmers_freq=0.1))
    reader = R.DbReader(cfg)
    sids, cids = ids_of(cfg)
    for name in ('tx1', 'tx2'):
        rows = reader.read(name, 150)
        data, s, c = OR.read_data_db(name, dbs, sids, cids, 0.1)
        assert rows.n_reads > 0 and same(rows.dense(), data)
        assert np.array_equal(rows.sample_ids, s) and np.array_equal(rows.condition_ids, c)
    assert reader.read('absent', 150).n_reads == 0
    reader.close()


# ---------------------------------------------------------------------------------------------
# rows -> CSR
# ---------------------------------------------------------------------------------------------

def random_rows(rng, R_, P):
    inten = rng.normal(0, 1, (R_, P)).astype(np.float32)
    dwell = rng.random((R_, P)).astype(np.float32)
    inten[rng.random((R_, P)) < 0.4] = np.nan
    dwell[rng.random((R_, P)) < 0.4] = np.nan
    return R.TranscriptRows(P, inten, dwell, rng.integers(0, 4, R_), rng.integers(0, 2, R_))


@pytest.mark.parametrize('offset', [0, 1, 7])
def test_pack_rows_matches_dense_packing(offset):
    rng = np.random.default_rng(5 + offset)
    sets = [random_rows(rng, 37, 1300), random_rows(rng, 5, 3), random_rows(rng, 0, 10), random_rows(rng, 12, 700)]
    batch, ptx, pidx = R.pack_rows(sets, offset, threads=3)
    V = 3 if offset else 2
    assert batch.signal.shape[1] == V
    want_sig, want_lab, want_cnt = [], [], []
    for r in sets:
        d = r.dense()
        if offset:      # the third variable of Worker._prepare_data (run.py:575-584)
            d = np.pad(d, ((0, 0), (0, 0), (0, 1)), constant_values=np.nan)
            if r.n_positions > offset:
                d[:r.n_positions - offset, :, 2] = d[offset:, :, 1]
        valid = ~np.isnan(d).all(2)
        want_sig.append(d[valid])
        want_lab.append(np.broadcast_to(r.labels()[None, :], valid.shape)[valid])
        want_cnt.append(valid.sum(1))
    cnt = np.concatenate(want_cnt)
    assert np.array_equal(batch.pos_offsets.numpy(), np.concatenate([[0], np.cumsum(cnt)]))
    assert same(batch.signal.numpy(), np.concatenate(want_sig).reshape(-1, V))
    assert np.array_equal(batch.label.numpy(), np.concatenate(want_lab))
    assert np.array_equal(ptx, np.repeat(np.arange(4), [1300, 3, 10, 700]))
    if not offset:      # and it is the batch the dense packer builds
        o, s, lab = pack_csr_numpy([(r.dense(), r.sample_ids, r.condition_ids) for r in sets])
        assert np.array_equal(o, batch.pos_offsets.numpy()) and same(s, batch.signal.numpy())


def test_invalid_ratio_matches_restatement():
    rng = np.random.default_rng(2)
    rows = random_rows(rng, 25, 400)
    rows.intensity[3, :] = np.nan
    got = rows.invalid_ratios()
    ok = ~np.isnan(rows.intensity).all(1)
    want = OR.reads_invalid_ratio(rows.intensity[ok].T)
    assert np.array_equal(got[ok], want) and np.isnan(got[3])


def test_select_reads_matches_downsample():
    """readers.select_reads = Worker._downsample (run.py:518-529) after _prepare_data and the
    coverage filter, checked against the restated prep functions on the dense tensor."""
    from oracle import prep as oprep
    rng = np.random.default_rng(9)
    rows = random_rows(rng, 60, 120)
    rows.condition_ids[:] = np.arange(60) % 2
    kept, mask = R.select_reads(rows, 20, 5)
    data, s, c, pos = oprep.prepare_data(rows.dense(), rows.sample_ids, rows.condition_ids)
    data, pos = oprep.filter_low_cov_positions(data, pos, c, 5)
    d2, s2, c2 = oprep.downsample(data, s, c, 20)
    assert np.array_equal(np.nonzero(mask)[0], pos)
    assert kept.n_reads == d2.shape[1] == 40
    assert same(kept.dense()[mask][:, :, 0], d2[:, :, 0])
    assert np.array_equal(kept.sample_ids, s2) and np.array_equal(kept.condition_ids, c2)
    same_rows, none = R.select_reads(rows, 60, 5)
    assert same_rows is rows and none is None


def test_encode_kmer():
    assert R.encode_kmer('AAAAA') == 0 and R.encode_kmer('ACGTT') == 0b0010011111
    with pytest.raises(RuntimeError):
        R.encode_kmer('ACGNN')


def test_encode_kmers_is_the_scalar_function_per_row():
    rng = np.random.default_rng(0)
    seq = ''.join(rng.choice(list('ACGTacgt'), 300))
    pos = np.array([0, 1, 5, 100, 294, 295, 296, 299, 298])       # the last k-mers are cut short by the end
    for klen in (5, 9):
        assert list(R.encode_kmers(seq, pos, klen)) == [R.encode_kmer(seq[p:p + klen]) for p in pos]
    assert R.encode_kmers(seq, np.array([], dtype=np.int64), 5).size == 0
    bad = seq[:50] + 'N' + seq[50:]
    assert list(R.encode_kmers(bad, np.array([10, 60]), 5)) == [R.encode_kmer(bad[10:15]), R.encode_kmer(bad[60:65])]
    with pytest.raises(RuntimeError, match="Bad nucleotide in kmer"):
        R.encode_kmers(bad, np.array([10, 48]), 5)


def test_corrupt_length_fields_are_errors_not_crashes(tmp_path):
    """A length field of a corrupt file must never size a buffer: negative / absurd l_text, n_ref,
    l_name, record block_size and BGZF ISIZE all end in an NcbError with a message (nothing is thrown
    across the C ABI)."""
    import struct
    from nanocompore_b200 import _lib as L
    rng = np.random.default_rng(21)

    def bam_bytes(l_text=None, n_ref=None, l_name=None, block_size=None):
        text = b'@HD\tVN:1.6\n'
        out = b'BAM\1' + struct.pack('<i', len(text) if l_text is None else l_text) + text
        out += struct.pack('<i', 1 if n_ref is None else n_ref)
        name = b'tx\0'
        out += struct.pack('<i', len(name) if l_name is None else l_name) + name + struct.pack('<i', 120)
        read = SF.make_uncalled4_read(rng, 120, 'RNA002', 'r0')
        tags = SF._aux_array('ur', 'i', read['ur']) + SF._aux_array('uc', 's', read['uc']) + \
            SF._aux_array('ul', 's', read['ul'])
        rec = SF._record(0, 0, 'r0', 0, tags)
        if block_size is not None:
            rec = struct.pack('<i', block_size) + rec[4:]
        return out + rec

    def write(path, payload, isize=None):
        block = SF._bgzf_block(payload)
        if isize is not None:
            block = block[:-4] + struct.pack('<I', isize)
        with open(path, 'wb') as f:
            f.write(block + SF._bgzf_block(b''))

    good = str(tmp_path / 'good.bam')
    write(good, bam_bytes())
    assert R.NativeBam(good).read_uncalled4('tx', 120, 'RNA002')[0].shape == (1, 120)
    cases = dict(l_text=dict(l_text=-5), l_text_huge=dict(l_text=2 ** 31 - 1), n_ref=dict(n_ref=-1),
                 n_ref_huge=dict(n_ref=2 ** 31 - 1), l_name=dict(l_name=-3), l_name_huge=dict(l_name=2 ** 30),
                 block_size=dict(block_size=-8), block_size_huge=dict(block_size=2 ** 31 - 1))
    for tag, kw in cases.items():
        p = str(tmp_path / f'{tag}.bam')
        write(p, bam_bytes(**kw))
        with pytest.raises(L.NcbError):
            R.NativeBam(p)
    for isize in (0xffffffff, 70000):
        p = str(tmp_path / f'isize_{isize}.bam')
        write(p, bam_bytes(), isize=isize)
        with pytest.raises(L.NcbError, match='ISIZE'):
            R.NativeBam(p)


def test_pipeline_skips_a_transcript_whose_files_do_not_parse(tmp_path, monkeypatch):
    """run.py:481-484: an error on one transcript is logged and the worker carries on.  The host half
    of the pipeline (read -> select -> pack; no GPU) keeps the other transcripts of the batch."""
    rng = np.random.default_rng(33)
    refs = [('txA', 150), ('txB', 150), ('txC', 150)]
    paths = []
    for si, sample in enumerate(('kd1', 'kd2', 'wt1', 'wt2')):
        reads = []
        for ri, (name, length) in enumerate(refs):
            for k in range(8):
                r = SF.make_uncalled4_read(rng, length, 'RNA002', f'{sample}-{name}-{k}', ref=ri)
                if sample == 'wt1' and name == 'txB' and k == 3:
                    r['uc'] = r['uc'][:-7]            # signal shorter than its ur segments
                reads.append(r)
        p = str(tmp_path / f'{sample}.bam')
        SF.write_bam(p, refs, reads)
        paths.append(p)
    cfg = bam_config(paths, max_invalid_kmers_freq=1.1, min_coverage=1)

    class T:
        def __init__(self, name, length):
            self.name, self.length = name, length

    messages = []
    pack_rows = R.pack_rows
    monkeypatch.setattr(R, 'pack_rows',
                        lambda rows, offset=0, pin=False, threads=None, wire='f32': pack_rows(rows, offset, wire=wire))
    pipe = R.TranscriptPipeline(cfg, R.Uncalled4Reader(cfg), comparator=None,
                                log=lambda level, msg: messages.append((level, msg)))
    transcripts, live, row_sets, masks, packed = pipe.prepare([T(n, l) for n, l in refs])
    assert live == [0, 2] and pipe.skipped == ['txB']
    assert len(messages) == 1 and messages[0][0] == 'error' and 'txB' in messages[0][1] and 'does not fit' in messages[0][1]
    batch, ptx, pidx = packed
    assert batch.n_positions == 300 and set(ptx.tolist()) == {0, 1}
    assert [r.n_reads for r in row_sets] == [32, 32]


def test_db_reader_rejects_blobs_of_the_wrong_length(tmp_path):
    """Two blobs whose lengths are off by +1 / -1 positions add up to the right total: only a per-blob
    check keeps the rows aligned (the reference's tensor assignment raises, run.py:795-801)."""
    rng = np.random.default_rng(8)
    P = 40
    paths = {}
    for sample in ('kd1', 'kd2', 'wt1', 'wt2'):
        reads = [(f'{sample}{k}', k, 0.0) for k in range(2)]
        lens = (P + 1, P - 1) if sample == 'wt2' else (P, P)
        signal = [(1, k, rng.normal(100, 5, n).astype(np.float32), rng.random(n).astype(np.float32))
                  for k, n in enumerate(lens)]
        paths[sample] = str(tmp_path / f'{sample}.db')
        SF.write_db(paths[sample], [('tx1', 1)], reads, signal)
    cfg = PathConfig(dict(data={'kd': {'kd1': {'db': paths['kd1']}, 'kd2': {'db': paths['kd2']}},
                                'wt': {'wt1': {'db': paths['wt1']}, 'wt2': {'db': paths['wt2']}}},
                          depleted_condition='kd', max_invalid_kmers_freq=0.1))
    reader = R.DbReader(cfg)
    with pytest.raises(ValueError, match='wt2'):
        reader.read('tx1', P)
    with pytest.raises(ValueError, match='wt2'):
        reader.read('tx1')
    reader.close()


def test_long_read_names_are_not_truncated(tmp_path):
    """Reads are ordered by read id (run.py:698-700): names that differ only after the 200th character
    must still sort by their full text."""
    rng = np.random.default_rng(4)
    refs = [('tx', 100)]
    stem = 'x' * 230
    paths = []
    for sample in ('kd1', 'kd2', 'wt1', 'wt2'):
        reads = [SF.make_uncalled4_read(rng, 100, 'RNA002', f'{stem}-{sample}-{k}', n_segments=1) for k in (2, 0, 1)]
        p = str(tmp_path / f'{sample}_long.bam')
        SF.write_bam(p, refs, reads)
        paths.append(p)
    rows = R.Uncalled4Reader(bam_config(paths, max_invalid_kmers_freq=1.1)).read('tx', 100)
    names = list(rows.read_names)
    assert len(names) == 12 and all(len(n) > 230 for n in names) and names == sorted(names)
    assert names[0] == f'{stem}-kd1-0' and names[-1] == f'{stem}-wt2-2'


def test_damaged_bams_never_crash_the_reader():
    """tests/fuzz_bam.py in a child process (a crash must fail this test, not end the test run):
    250 damaged copies of a valid BAM, each read back -- a result or an NcbError, never a crash."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'fuzz_bam.py'), '11', '250'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-1500:])
    words = r.stdout.split()
    assert words[0] == 'survived' and int(words[1]) + int(words[3]) == 250 and int(words[3]) > 50
