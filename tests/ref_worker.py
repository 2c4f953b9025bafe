"""
Runs the reference's OWN worker loop -- ``Uncalled4Worker.run`` (run.py:364-491, 649-718), unmodified,
imported from /root/reference or from the staged copy baseline/_ref -- in this process over the
reference's fixture BAMs, optionally with ``nanocompore.run.TranscriptComparator`` replaced by another
class.  This is the drop-in proof of the boundary: everything around the comparator is the
reference's code -- task queue protocol, ``_read_data`` / ``Uncalled4.get_data``, ``_prepare_data``,
``_filter_low_cov_positions``, ``_downsample``, the k-mer column and ``ResultsDB.save_test_results`` into
a real SQLite results database.

Absent third-party packages are stood in for by test-side shims: pysam -> oracle/bam_shim.py (stdlib BAM
reader), pyfaidx.Fasta -> a dict of sequences, gmm_gpu -> oracle/gmm.py (only when the reference's own
comparator runs).  Test infrastructure.
"""
import os
import queue
import sqlite3
import threading

import numpy as np


def read_fasta(path):
    seqs, name = {}, None
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith('>'):
                name = line[1:].split()[0]
                seqs[name] = []
            elif name:
                seqs[name].append(line)
    return {k: ''.join(v) for k, v in seqs.items()}


class TaskQueue(queue.Queue):
    """The slice of multiprocessing.JoinableQueue the worker uses (qsize, get(block=False), put,
    task_done); queue.Queue raises the same queue.Empty."""


class Counter:
    value = 0


def run_worker(outpath, comparator_cls=None, device='cpu', **config_overrides):
    """Returns (rows of the results database as a dict of numpy columns, the worker's log messages)."""
    from loguru import logger
    from oracle import bam_shim
    from oracle import refharness as H
    logger.remove()                 # the worker logs every step at DEBUG
    ref = H.load_reference()
    import nanocompore.common as common
    import nanocompore.run as run
    from nanocompore.database import ResultsDB

    fixtures = os.path.join(H.reference_root(), 'tests', 'fixtures')
    cfg = {
        'data': {'kd': {'kd1': {'bam': os.path.join(fixtures, 'kd1.bam')},
                        'kd2': {'bam': os.path.join(fixtures, 'kd2.bam')}},
                 'wt': {'wt1': {'bam': os.path.join(fixtures, 'wt1.bam')},
                        'wt2': {'bam': os.path.join(fixtures, 'wt2.bam')}}},
        'fasta': os.path.join(fixtures, 'test_reference.fa'),
        'depleted_condition': 'kd', 'kit': 'RNA002', 'resquiggler': 'uncalled4',
        'outpath': str(outpath), 'min_coverage': 1,          # tests/test_run.py:33-35
    }
    cfg.update(config_overrides)
    conf = ref.config.Config(cfg)

    fasta = read_fasta(cfg['fasta'])
    saved = (run.pysam.AlignmentFile, common.pysam.AlignmentFile, run.Fasta, run.TranscriptComparator)
    run.pysam.AlignmentFile = bam_shim.AlignmentFile
    common.pysam.AlignmentFile = bam_shim.AlignmentFile
    run.Fasta = lambda path: fasta
    if comparator_cls is not None:
        run.TranscriptComparator = comparator_cls
    messages = []
    try:
        ResultsDB(conf, init_db=True)                                          # run.py:100
        transcripts = run.RunCmd(conf)._get_transcripts_for_processing()        # run.py:101
        tasks = TaskQueue()
        for t in sorted(transcripts, key=lambda t: t.ref_id):
            tasks.put((t, 0))
        worker = run.Uncalled4Worker(0, tasks, threading.Lock(), threading.Lock(), Counter(), device, conf)
        log = worker.log
        worker.log = lambda level, msg: (messages.append((level, msg)), log(level, msg))[1]
        worker.run()              # in this process: multiprocessing.Process.run is the body of the child
        db_path = ResultsDB(conf, init_db=False).db_path
    finally:
        run.pysam.AlignmentFile, common.pysam.AlignmentFile, run.Fasta, run.TranscriptComparator = saved
    with sqlite3.connect(db_path) as con:
        cur = con.execute("SELECT t.name, r.* FROM kmer_results r JOIN transcripts t ON r.transcript_id = t.id "
                          "ORDER BY t.name, r.pos")
        names = [d[0] for d in cur.description]
        rows = cur.fetchall()
    cols = {}
    for j, name in enumerate(names):
        if name in cols:
            continue
        v = [r[j] for r in rows]
        try:
            cols[name] = np.array([np.nan if x is None else x for x in v], dtype=np.float64)
        except (TypeError, ValueError):
            cols[name] = np.array(v)
    return cols, messages
