"""
Import the UNMODIFIED reference (/root/reference/nanocompore) in the build container.

Used only to validate the oracle and to generate the golden vectors under
tests/golden/ (tests/golden/make_golden.py).  /root/reference does not exist on the
GPU box, so nothing that runs there may call ``load_reference()``.

The reference imports packages that are absent here (pysam, pyfaidx, schema,
gmm_gpu, statsmodels, polars, gtfparse) and asks importlib.metadata for its own
version.  Those are stubbed in ``sys.modules``; ``gmm_gpu.gmm.GMM`` is replaced by
``oracle.gmm.GMM`` (the reference's real GMM package cannot be obtained -- see
oracle/gmm.py), so reports produced through this harness must say
"reference code, restated gmm_gpu".

Test infrastructure only (see oracle/__init__.py).
"""
import importlib
import importlib.metadata
import os
import sys
import types

REFERENCE_ROOT = '/root/reference'
# the byte-for-byte copy baseline/stage_reference.py makes for the reference arm of bench.py (git-ignored;
# it travels to the GPU box, where /root/reference does not exist)
STAGED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', '_ref')


def reference_root():
    """Where the unmodified reference package can be imported from: /root/reference in the build
    container, else the staged copy; None when neither is there."""
    for root in (REFERENCE_ROOT, STAGED_ROOT):
        if os.path.isfile(os.path.join(root, 'nanocompore', 'comparisons.py')):
            return root
    return None


def reference_available():
    return reference_root() is not None


class _Schema:
    def __init__(self, *a, **k):
        pass

    def validate(self, x):
        return x


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    mod('pysam', AlignmentFile=object)
    mod('pyfaidx', Fasta=object)
    mod('schema', Schema=_Schema, And=_Schema, Or=_Schema, Optional=_Schema,
        SchemaError=type('SchemaError', (Exception,), {}))
    from oracle.gmm import GMM
    mod('gmm_gpu')
    mod('gmm_gpu.gmm', GMM=GMM)
    mod('polars')
    mod('gtfparse', read_gtf=lambda *a, **k: None)
    mod('statsmodels')
    mod('statsmodels.stats')
    from oracle.fdr import fdr_bh
    mod('statsmodels.stats.multitest',
        multipletests=lambda p, method='fdr_bh', **k: (None, fdr_bh(p)))
    orig = importlib.metadata.version
    if not getattr(importlib.metadata.version, '_ncb_patched', False):
        def version(name):
            return '2.0.0' if name == 'nanocompore' else orig(name)
        version._ncb_patched = True
        importlib.metadata.version = version


def load_reference():
    """Returns the reference's ``nanocompore`` package (comparisons, run, config ...)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("neither /root/reference nor the staged copy baseline/_ref is present "
                           "(python baseline/stage_reference.py makes the copy in the build container)")
    _install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    pkg = importlib.import_module('nanocompore')
    for sub in ('common', 'config', 'transcript', 'comparisons'):
        importlib.import_module(f'nanocompore.{sub}')
    return pkg


class RefWorker:
    """Duck-typed worker: the comparator only calls ``worker.log`` (comparisons.py:86...)."""

    def __init__(self):
        self.messages = []

    def log(self, level, msg):
        self.messages.append((level, msg))


def reference_config(ref, **overrides):
    """A reference ``Config`` over a minimal 2x2 layout (schema validation is stubbed)."""
    cfg = {
        'data': {'kd': {'kd1': {'bam': ''}, 'kd2': {'bam': ''}},
                 'wt': {'wt1': {'bam': ''}, 'wt2': {'bam': ''}}},
        'fasta': '', 'depleted_condition': 'kd', 'kit': 'RNA002', 'resquiggler': 'uncalled4',
    }
    cfg.update(overrides)
    return ref.config.Config(cfg)
