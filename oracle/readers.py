"""
Transcript readers -- numpy restatement of the reference's two ``_read_data`` implementations
(/root/reference/nanocompore):

  * ``uncalled4_get_data``      uncalled4.py:57-201 (``Uncalled4.get_data``,
                                ``_copy_signal_to_tensor``, ``_resolve_skips``)
  * ``read_data_uncalled4``     run.py:657-718 (``Uncalled4Worker._read_data``)
  * ``reads_invalid_ratio``     common.py:298-329
  * ``read_data_db``            run.py:741-836 (``GenericWorker._read_data``) over
                                database.py:125-141, 425-452

Test infrastructure only (see oracle/__init__.py).  BAM files are read with the stdlib parser
``oracle/bam_shim.py`` (pysam is not installed here).

Pinned: ``tests/golden/fixture_cfg1.npz`` holds the tensors the UNMODIFIED reference parser
produced from its fixture BAMs; ``tests/test_readers.py`` checks this restatement against them
and against the known answers of the reference's tests/test_uncalled4.py:12-100 and
tests/test_run.py:241-308.
"""
import sqlite3
from contextlib import closing

import numpy as np

from oracle import bam_shim

INTENSITY_POS, DWELL_POS = 0, 1
KITS = {'RNA002': (5, 4), 'RNA004': (9, 5)}      # common.py:21-32: (len, center)

QUERY = """
SELECT intensity, dwell
FROM signal_data sd
JOIN reads r ON sd.read_id = r.id
JOIN transcripts t ON sd.transcript_id = t.id
WHERE t.name = ?
  AND r.invalid_kmers <= ?
"""


def resolve_skips(ul):
    """uncalled4.py:195-200: a zero length repeats the previous kmer's length."""
    resolved = np.empty(len(ul))
    resolved[0] = ul[0]
    for i in range(1, len(ul)):
        resolved[i] = ul[i] if ul[i] != 0 else resolved[i - 1]
    return resolved


def copy_signal_to_tensor(read, tensor, read_index, kit):
    """uncalled4.py:118-192."""
    klen, kcenter = KITS[kit]
    ur = read.get_tag('ur')
    uc = np.array(read.get_tag('uc')[::-1])
    ul = np.array([d for d in read.get_tag('ul')[::-1] if d >= 0])
    ul = resolve_skips(ul)
    segments = [(ur[i], ur[i + 1]) for i in range(0, len(ur), 2)]
    signal_start = 0
    for ref_start, ref_end in segments:
        if ref_start == ur[0]:
            ref_start += klen - kcenter
        if ref_end == ur[-1]:
            ref_end -= kcenter - 1
        ref_start -= klen - kcenter
        ref_end -= klen - kcenter
        size = ref_end - ref_start
        signal_end = signal_start + size
        tensor[ref_start:ref_end, read_index, DWELL_POS] = ul[signal_start:signal_end]
        tensor[ref_start:ref_end, read_index, INTENSITY_POS] = uc[signal_start:signal_end]
        signal_start += size


def uncalled4_get_data(ref_id, ref_len, bams, kit):
    """uncalled4.py:57-115.  ``bams``: {sample label: bam_shim.AlignmentFile}, walked in dict
    order (the reference takes its samples ``as_completed``; the caller's sort by read id makes
    the order immaterial).  Returns (tensor (P, R, 2) float32, read ids, sample labels)."""
    read_ids, sample_labels = [], []
    read_count = sum(bam.count(reference=ref_id) if ref_id in bam.references else 0 for bam in bams.values())
    tensor = np.full((ref_len, read_count, 2), np.nan, dtype=np.float32)
    read_index = 0
    for sample, bam in bams.items():
        if ref_id not in bam.references:
            continue
        for read in bam.fetch(ref_id):
            if read.is_secondary or read.is_supplementary:
                continue
            copy_signal_to_tensor(read, tensor, read_index, kit)
            read_ids.append(read.query_name)
            sample_labels.append(sample)
            read_index += 1
    tensor = tensor[:, :read_index, :]
    tensor = np.where(tensor == -32768, np.nan, tensor)
    return tensor, np.array(read_ids), np.array(sample_labels)


def reads_invalid_ratio(intensity):
    """common.py:298-329.  intensity (positions, reads)."""
    nreads = intensity.shape[1]
    ratios = np.empty(nreads)
    for n in range(nreads):
        valid_positions = np.where(~np.isnan(intensity[:, n]))[0]
        length = valid_positions.max() - valid_positions.min() + 1
        ratios[n] = (length - len(valid_positions)) / length
    return ratios


def read_data_uncalled4(ref_id, ref_len, bam_paths, sample_ids, sample_to_condition_id, kit,
                        max_invalid_kmers_freq=0.1):
    """run.py:692-718.  ``bam_paths``: {sample label: path} in config order.  Returns
    (tensor (P, R, 2), sample ids, condition ids) with the reads ordered by read id (stable)."""
    bams = {s: bam_shim.AlignmentFile(p, 'rb') for s, p in bam_paths.items()}
    data, reads, samples = uncalled4_get_data(ref_id, ref_len, bams, kit)
    if len(reads) == 0:
        return data, np.array([], dtype=np.int64), np.array([], dtype=np.int64)
    order = np.argsort(reads, kind='stable')
    data, samples = data[:, order, :], samples[order]
    sids = np.array([sample_ids[s] for s in samples], dtype=np.int64)
    cids = np.array([sample_to_condition_id[s] for s in samples], dtype=np.int64)
    ok = reads_invalid_ratio(data[:, :, INTENSITY_POS]) < max_invalid_kmers_freq
    return data[:, ok], sids[ok], cids[ok]


def read_data_db(ref_id, db_paths, sample_ids, sample_to_condition_id, max_invalid_kmers_freq=0.1):
    """run.py:766-836.  ``db_paths``: {sample label: path} in config order."""
    intensities, dwells, counts = {}, {}, {}
    for sample, path in db_paths.items():
        with closing(sqlite3.connect(path)) as conn, closing(conn.cursor()) as cur:
            signal_data = cur.execute(QUERY, (ref_id, max_invalid_kmers_freq)).fetchall()
        counts[sample] = len(signal_data)
        if not signal_data:
            continue
        intensities[sample] = np.array([np.frombuffer(i, dtype=np.float32) for i, _ in signal_data])
        dwells[sample] = np.array([np.frombuffer(d, dtype=np.float32) for _, d in signal_data])
    if all(c == 0 for c in counts.values()):
        return np.empty((0, 0, 0)), np.array([]), np.array([])
    tensor = np.stack([np.concatenate([intensities[s] for s in db_paths if s in intensities]).T,
                       np.concatenate([dwells[s] for s in db_paths if s in dwells]).T], axis=2)
    sids = np.concatenate([np.full(counts[s], sample_ids[s]) for s in sample_ids])
    cids = np.concatenate([np.full(counts[s], sample_to_condition_id[s]) for s in sample_ids])
    return tensor, sids, cids
