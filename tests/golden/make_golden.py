"""
Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference/nanocompore, imported through oracle/refharness.py) in the build container.

    python tests/golden/make_golden.py

Outputs (committed; /root/reference does not exist on the GPU box):
  fixture_cfg1.npz      BASELINE config 1: the reference's bundled fixture BAMs read by the
                        reference's own Uncalled4 parser (through oracle/bam_shim.py),
                        the raw per-transcript tensors, and the result columns of the
                        reference's compare_transcript with the default methods [GMM, KS]
  synthetic_ref.npz     seeded synthetic transcripts and the reference's result columns for
                        KS, MW, TT and GMM
  helpers_ref.npz       outputs of the reference's module-level helpers on fixed inputs
  reference_tests.json  result of running the reference's own tests/test_comparisons.py against
                        this repo's restated gmm_gpu

gmm_gpu (the reference's GMM package) is absent: every GMM column below was produced by the
reference's code calling oracle/gmm.py -- "reference code, restated gmm_gpu".
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle.refharness import REFERENCE_ROOT, RefWorker, load_reference  # noqa: E402

FIXTURES = os.path.join(REFERENCE_ROOT, 'tests', 'fixtures')


def frame_to_arrays(df, prefix):
    out = {}
    for col in df.columns:
        v = df[col].to_numpy()
        if v.dtype == object:
            v = v.astype(str)
        out[f'{prefix}/{col}'] = v
    out[f'{prefix}/__columns__'] = np.array(list(df.columns))
    return out


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


def fixture_config1(ref):
    import torch
    from oracle import bam_shim
    from nanocompore.uncalled4 import Uncalled4
    from nanocompore.common import get_reads_invalid_ratio, INTENSITY_POS
    from nanocompore.run import Worker
    from nanocompore.transcript import Transcript
    from oracle.refharness import reference_config

    cfg_dict = {
        'data': {'kd': {'kd1': {'bam': os.path.join(FIXTURES, 'kd1.bam')},
                        'kd2': {'bam': os.path.join(FIXTURES, 'kd2.bam')}},
                 'wt': {'wt1': {'bam': os.path.join(FIXTURES, 'wt1.bam')},
                        'wt2': {'bam': os.path.join(FIXTURES, 'wt2.bam')}}},
        'fasta': os.path.join(FIXTURES, 'test_reference.fa'),
        'depleted_condition': 'kd', 'kit': 'RNA002', 'resquiggler': 'uncalled4',
        'min_coverage': 1,
    }
    conf = ref.config.Config(cfg_dict)
    fasta = read_fasta(cfg_dict['fasta'])
    bams = {sample: bam_shim.AlignmentFile(bam, 'rb')
            for sample, _, bam in conf.get_sample_condition_bam_data()}

    class FakeWorker:           # the slice of Worker state the prep functions read
        _conf = conf
        _device = 'cpu'

        def log(self, level, msg):
            pass
    worker = FakeWorker()
    comparator = ref.comparisons.TranscriptComparator(config=conf, worker=worker)

    out = {}
    total_rows = 0
    for ti, name in enumerate(sorted(fasta)):
        seq = fasta[name]
        data, reads, samples = Uncalled4(name, len(seq), bams, conf.get_kit()).get_data()
        # Uncalled4Worker._read_data (run.py:692-718): read order, ids, invalid-kmer filter
        order = np.argsort(reads)
        data, samples = data[:, order, :], samples[order]
        sample_ids = np.vectorize(conf.get_sample_ids().get)(samples)
        s2c = np.vectorize(conf.sample_to_condition().get)
        condition_ids = np.vectorize(conf.get_condition_ids().get)(s2c(samples))
        ok = get_reads_invalid_ratio(data[:, :, INTENSITY_POS]) < conf.get_max_invalid_kmers_freq()
        data, sample_ids, condition_ids = data[:, ok], sample_ids[ok], condition_ids[ok]
        raw = data.copy()
        tensor, s_t, c_t, pos_t = Worker._prepare_data(worker, data, sample_ids, condition_ids)
        tensor, pos_t = Worker._filter_low_cov_positions(worker, tensor, pos_t, c_t, conf.get_min_coverage())
        transcript = Transcript(ti, name, seq)
        _, df = comparator.compare_transcript(transcript, tensor.clone(), s_t, c_t, pos_t, 'cpu')
        total_rows += len(df)
        out[f'tx{ti}/name'] = np.array(name)
        out[f'tx{ti}/raw'] = raw.astype(np.float32)
        out[f'tx{ti}/sample_ids'] = sample_ids.astype(np.int16)
        out[f'tx{ti}/condition_ids'] = condition_ids.astype(np.int16)
        out[f'tx{ti}/prepared_positions'] = pos_t.numpy()
        out |= frame_to_arrays(df, f'tx{ti}/result')
    out['n_transcripts'] = np.array(len(fasta))
    out['total_rows'] = np.array(total_rows)
    out['sample_labels'] = np.array(list(conf.get_sample_ids().keys()))
    out['sample_ids'] = np.array(list(conf.get_sample_ids().values()))
    out['depleted_condition_id'] = np.array(conf.get_condition_ids()[conf.get_depleted_condition()])
    np.savez_compressed(os.path.join(HERE, 'fixture_cfg1.npz'), **out)
    print('fixture_cfg1: rows', total_rows, '(tests/test_run.py:63-64 pins 3520 TSV lines = 3519 rows + header)')


def synthetic(ref):
    import torch
    from nanocompore.run import Worker
    from nanocompore.transcript import Transcript
    from nanocompore_b200.synthetic import make_transcript
    from oracle.refharness import reference_config

    out = {}
    cases = [dict(seed=1, P=60, n=40, quantised=False), dict(seed=2, P=60, n=40, quantised=True),
             dict(seed=3, P=40, n=90, quantised=False), dict(seed=4, P=80, n=8, quantised=True,
                                                             gap_rate=0.1),
             # motor_dwell_offset > 0: third variable, 2- and 3-variable mixtures
             dict(seed=5, P=70, n=40, quantised=False, motor=6, gap_rate=0.08),
             dict(seed=6, P=50, n=30, quantised=True, motor=3, gap_rate=0.02)]
    for ci, case in enumerate(cases):
        kw = {k: v for k, v in case.items() if k not in ('seed', 'P', 'n', 'motor')}
        tx = make_transcript(np.random.default_rng(case['seed']), case['P'], case['n'], **kw)
        conf = reference_config(ref, comparison_methods=['GMM', 'KS', 'MW', 'TT'], min_coverage=1,
                                motor_dwell_offset=case.get('motor', 0))

        class FakeWorker:
            _conf = conf
            _device = 'cpu'

            def log(self, level, msg):
                pass
        worker = FakeWorker()
        raw = tx.data.copy()
        tensor, s_t, c_t, pos_t = Worker._prepare_data(worker, tx.data.copy(), tx.samples, tx.conditions)
        tensor, pos_t = Worker._filter_low_cov_positions(worker, tensor, pos_t, c_t, 1)
        comparator = ref.comparisons.TranscriptComparator(config=conf, worker=worker)
        _, df = comparator.compare_transcript(Transcript(ci, f'syn{ci}', 'A' * case['P']),
                                              tensor.clone(), s_t, c_t, pos_t, 'cpu')
        out[f'case{ci}/motor_dwell_offset'] = np.array(case.get('motor', 0))
        out[f'case{ci}/raw'] = raw
        out[f'case{ci}/prepared'] = tensor.numpy()
        out[f'case{ci}/prepared_positions'] = pos_t.numpy()
        out[f'case{ci}/samples'] = tx.samples
        out[f'case{ci}/conditions'] = tx.conditions
        out |= frame_to_arrays(df, f'case{ci}/result')
    out['n_cases'] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, 'synthetic_ref.npz'), **out)
    print('synthetic_ref:', len(cases), 'cases')


def helpers(ref):
    import torch
    C = ref.comparisons
    rng = np.random.default_rng(7)
    out = {}
    x = rng.standard_normal((5, 40, 2)).astype(np.float32)
    x[rng.random(x.shape) < 0.2] = np.nan
    out['nanstd/in'] = x
    out['nanstd/out'] = C.nanstd(torch.from_numpy(x), 1).numpy()
    cont = rng.integers(1, 200, size=(64, 2, 2)).astype(np.float32)
    out['lor/in'] = cont
    out['lor/out'] = C.calculate_lors(torch.from_numpy(cont)).numpy()
    pv = rng.uniform(1e-6, 1, size=60)
    pv[rng.random(60) < 0.1] = np.nan
    for ctx in (1, 2, 3):
        out[f'xcorr/{ctx}'] = C.cross_corr_matrix(pv, ctx)
    out['xcorr/in'] = pv
    w = C.harmomic_series(2)
    cor = C.cross_corr_matrix(pv, 2)
    out['hou/out'] = np.array([C.combine_pvalues_hou(np.nan_to_num(pv[i:i + 5], nan=1), w, cor)
                               for i in range(0, 50)])
    np.savez_compressed(os.path.join(HERE, 'helpers_ref.npz'), **out)
    print('helpers_ref written')


def reference_own_tests():
    """Runs the reference's tests/test_comparisons.py with gmm_gpu = oracle/gmm.py."""
    site = os.path.join(ROOT, 'build', 'ref_site')
    os.makedirs(site, exist_ok=True)
    with open(os.path.join(site, 'sitecustomize.py'), 'w') as f:
        f.write("import sys\nsys.path.insert(0, %r)\nfrom oracle import refharness\nrefharness._install_stubs()\n" % ROOT)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([site, REFERENCE_ROOT]))
    r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-p', 'no:cacheprovider',
                        '--import-mode=importlib', 'tests/test_comparisons.py'],
                       cwd=REFERENCE_ROOT, env=env, capture_output=True, text=True)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
    with open(os.path.join(HERE, 'reference_tests.json'), 'w') as f:
        json.dump({'command': 'pytest tests/test_comparisons.py (reference, gmm_gpu = oracle/gmm.py)',
                   'returncode': r.returncode, 'summary': tail}, f, indent=1)
    print('reference tests:', tail)
    if r.returncode != 0:
        print(r.stdout[-3000:])


if __name__ == '__main__':
    ref = load_reference()
    import importlib
    for sub in ('run', 'uncalled4'):
        importlib.import_module(f'nanocompore.{sub}')
    reference_own_tests()
    helpers(ref)
    synthetic(ref)
    fixture_config1(ref)
