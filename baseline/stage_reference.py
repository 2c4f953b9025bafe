"""
Stages the UNMODIFIED reference package for the reference arm of bench.py.

    python baseline/stage_reference.py          # /root/reference/nanocompore -> baseline/_ref/nanocompore

The base contract's way (``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse
--target baseline/_ref /root/reference``) fails in this image: the reference builds with hatchling
(pyproject.toml:66-68) and hatchling is neither installed nor in /opt/wheelhouse.  The package is
pure Python, so an install IS a copy of its directory: this script makes that copy, byte for byte,
and records the sha256 of every file in ``baseline/_ref/MANIFEST.json``.  ``baseline/_ref/`` is
git-ignored (no reference source enters this repository's history) but not gpurun-ignored, so it
travels to the GPU box, where /root/reference does not exist.

The arm imports the copy through ``oracle/refharness.py`` -- the same stubs for the absent
third-party packages (pysam, pyfaidx, schema, statsmodels, polars, gtfparse) and the same stand-in
for ``gmm_gpu.gmm.GMM`` (``oracle.gmm.GMM``: the package gmm-gpu==0.2.8 cannot be obtained), so its
lines say "reference code, restated gmm_gpu".
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/nanocompore'
DST = os.path.join(HERE, '_ref', 'nanocompore')
# the reference's own tests of the path (run against this repo's TranscriptComparator by
# tests/test_reference_tests_gpu.py) and the fixtures its Worker test reads
TESTS_SRC = '/root/reference/tests'
TESTS_DST = os.path.join(HERE, '_ref', 'tests')
TEST_FILES = ['common.py', 'test_comparisons.py', 'test_run.py', 'test_postprocessing.py']
FIXTURES = ['kd1.bam', 'kd2.bam', 'wt1.bam', 'wt2.bam', 'kd1.bam.bai', 'kd2.bam.bai', 'wt1.bam.bai', 'wt2.bam.bai',
            'test_reference.fa', 'test_reference.fa.fai']


def sha256(path):
    h = hashlib.sha256()
    with open(path, 'rb') as f:
        for chunk in iter(lambda: f.read(1 << 20), b''):
            h.update(chunk)
    return h.hexdigest()


def stage(src=SRC, dst=DST):
    if not os.path.isdir(src):
        raise RuntimeError(f"{src} is not present: the reference can only be staged in the build container")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    # the Python modules and the template config; the pore-model tables (models/) are not on the path
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__', 'models'))
    manifest = {}
    for root, _, files in os.walk(dst):
        for name in sorted(files):
            p = os.path.join(root, name)
            manifest[os.path.relpath(p, dst)] = sha256(p)
    for rel, digest in manifest.items():
        assert sha256(os.path.join(src, rel)) == digest, rel
    tests = {}
    if os.path.isdir(TESTS_SRC):
        if os.path.isdir(TESTS_DST):
            shutil.rmtree(TESTS_DST)
        os.makedirs(os.path.join(TESTS_DST, 'fixtures'))
        for name in TEST_FILES + [os.path.join('fixtures', f) for f in FIXTURES]:
            a, b = os.path.join(TESTS_SRC, name), os.path.join(TESTS_DST, name)
            if os.path.isfile(a):
                shutil.copyfile(a, b)
                tests[name] = sha256(b)
    with open(os.path.join(os.path.dirname(dst), 'MANIFEST.json'), 'w') as f:
        json.dump({'source': src, 'version': '2.0.0', 'files': manifest, 'tests': tests}, f, indent=1,
                  sort_keys=True)
    return dst


if __name__ == '__main__':
    print(stage(*(sys.argv[1:3])))
