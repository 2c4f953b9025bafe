"""
fp64 operations per position of the mixture kernels, from an `ncu --set full` capture.

    python profiles/fp64_ops.py <report.ncu-rep> <positions per launch> [kernel substring ...]

Sums the predicated-on thread-level executions of the DFMA / DADD / DMUL SASS instructions (source
page of the report) over the captured launches whose name contains one of the substrings (default: the mixture kernels
`position_kernel_warp<..., 1>`), divides by the positions of the batch and writes
profiles/fp64_ops.json, which bench.py turns into the fp64 entry of its roofline.
"""
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
rep, positions = sys.argv[1], int(sys.argv[2])
subs = sys.argv[3:] or [', 1>(']
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
tot = dict(dfma=0.0, dadd=0.0, dmul=0.0)
names = []
use, hdr = False, None
for r in csv.reader(src.splitlines()):
    if r and r[0] == 'Kernel Name':
        use = any(s in r[1].replace('(int)', '') for s in subs)
        if use:
            names.append(r[1])
        hdr = None
    elif r and r[0] == 'Address':
        hdr = r
    elif use and hdr and len(r) >= len(hdr):
        op = r[hdr.index('Source')].split()
        if not op:
            continue
        nm = (op[1] if op[0].startswith('@') else op[0]).split('.')[0].lower()
        if nm in tot:
            tot[nm] += float(r[hdr.index('Predicated-On Thread Instructions Executed')] or 0)
out = {f'{k}_per_position': v / positions for k, v in tot.items()}
out.update(positions=positions, kernels=names,
           source=f"{os.path.basename(rep)} (ncu --set full --clock-control none, "
                  f"bench.py --steps 2 --warmup 1 --tx-per-step {positions // 2000})")
json.dump(out, open(os.path.join(HERE, 'fp64_ops.json'), 'w'), indent=1)
print(json.dumps(out, indent=1))
