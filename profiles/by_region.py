"""Aggregates profiles/by_line.py output of the mixture kernel into code regions.

    python profiles/by_region.py <src.csv> <disasm> <mangled-kernel-substring>

Regions are found from marker comments / function names in position_kernel.cu, so the table
follows the source as it changes."""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.path.dirname(HERE), 'nanocompore_b200', 'csrc', 'position_kernel.cu')
MARKS = [('exp', 'double exp_nonpos_fast(double'), ('(other)', 'struct PosShared {'),
         ('m-step', 'double m_step(const'), ('(other)', 'double log_prob(const Comp'),
         ('tail (LOR, chi2, counts)', 'double gmm_tail('), ('e-step', 'double log_prob_e('),
         ('accumulate', 'void accumulate('), ('load + compact', 'void position_gmm('),
         ('K=1', '// K = 1: closed form'), ('farthest pair', '// K = 2: deterministic'),
         ('Lloyd', 'for (int it = 1; it <= 20'), ('EM loop', '// EM.  Pass 0'),
         ('final pass', '// final E-step with the final parameters:'), ('(other)', '// ---- 3-variable mixtures')]
lines = open(SRC).read().splitlines()
bounds = []
for name, mark in MARKS:
    for i, ln in enumerate(lines, 1):
        if mark in ln:
            bounds.append((i, name))
            break
bounds.sort()
out = subprocess.run([sys.executable, os.path.join(HERE, 'by_line.py')] + sys.argv[1:4] + ['100000'],
                     capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0.0, 0.0])
print(out.splitlines()[0])
for ln in out.splitlines()[1:]:
    m = re.match(r'\s*([\d.]+)% inst\s+([\d.]+)% samples\s+(\S+):(\d+)', ln)
    if not m:
        continue
    inst, samp, f, l = float(m.group(1)), float(m.group(2)), m.group(3), int(m.group(4))
    if f == 'position_kernel.cu':
        key = '(other)'
        for b, name in bounds:
            if l >= b:
                key = name
    elif f == 'group_ops.cuh':
        key = 'reductions (group_ops.cuh)'
    else:
        key = f
    agg[key][0] += inst
    agg[key][1] += samp
print("| region | % of warp instructions | % of stall samples |\n|---|---:|---:|")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| {k} | {v[0]:.1f} | {v[1]:.1f} |")
