"""
Build libncb.so (the C-ABI CUDA library) in-tree with plain nvcc for sm_100a.

    python -m nanocompore_b200.build            # incremental
    python -m nanocompore_b200.build --force

Output: nanocompore_b200/libncb.so (git-ignored; travels to the GPU box with the snapshot).
Objects go to build/ (git-ignored).  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(ROOT, 'include')
OBJDIR = os.path.join(ROOT, 'build', 'ncb')
LIB = os.path.join(HERE, 'libncb.so')

SOURCES = ['position_kernel.cu', 'gmm_split.cu', 'ks_pvalue.cu', 'fdr.cu', 'context.cu', 'api.cu', 'pack.cu', 'reader.cu', 'decode.cu', 'taskq.cu']
HEADERS = [os.path.join(CSRC, h) for h in ('ncb_internal.h', 'ncb_math.cuh', 'group_ops.cuh', 'gmm_common.cuh', 'bam_internal.h',
                                               'inflate_core.cuh', 'decode_core.cuh', 'decode_stage.h')] + \
          [os.path.join(INCLUDE, 'ncb.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-I' + INCLUDE, '-I' + CSRC]


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libncb cannot be built (there is no CPU fallback)")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), lib=None, objdir=None):
    """``defines`` / ``lib`` / ``objdir`` build an experimental variant next to the product
    library (tuning runs select it with NCB_LIB_PATH)."""
    global OBJDIR, LIB
    if lib:
        LIB = lib
    if objdir:
        OBJDIR = objdir
    os.makedirs(OBJDIR, exist_ok=True)
    exe = nvcc()
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _stale(o, [s] + HEADERS):
            cmd = [exe] + NVCC_FLAGS + ['-D' + d for d in defines] + \
                (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + ' '.join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        for out in ex.map(run, jobs):
            if verbose and out:
                print(out)
    if force or jobs or _stale(LIB, objs):
        run([exe, '-shared', '-o', LIB] + objs + ['-lcudart', '-lpthread', '-lz', '-lrt'])
    return LIB


if __name__ == '__main__':
    defs = [a[2:] for a in sys.argv[1:] if a.startswith('-D')]
    out = [a.split('=', 1)[1] for a in sys.argv[1:] if a.startswith('--out=')]
    if out:
        tag = os.path.splitext(os.path.basename(out[0]))[0]
        print(build(force=True, verbose='-v' in sys.argv, defines=defs, lib=os.path.abspath(out[0]),
                    objdir=os.path.join(ROOT, 'build', tag)))
    else:
        print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, defines=defs))
