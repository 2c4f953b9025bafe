"""
fp64 operations per position and DRAM bytes per launch of the mixture kernels, from an
`ncu --set full` capture.

    python profiles/fp64_ops.py <report.ncu-rep> <positions per launch> [kernel substring ...]

Sums the predicated-on thread-level executions of the DFMA / DADD / DMUL SASS instructions (source
page of the report) over the captured launches whose name contains one of the substrings (default:
the three split mixture kernels gmm_init / gmm_em / gmm_final of gmm_split.cu).  A kernel captured
more than once counts with the mean of its captures; the kernels are then added up (one launch of
each makes the mixture fit of a batch), divided by the positions of the batch and written to
profiles/fp64_ops.json, which bench.py turns into the fp64 entry of its roofline.  The DRAM bytes
(raw page: dram__bytes_read.sum + dram__bytes_write.sum) go to profiles/traffic.json the same way.
"""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
rep, positions = sys.argv[1], int(sys.argv[2])
subs = sys.argv[3:] or ['gmm_init_kernel', 'gmm_em_kernel', 'gmm_final_kernel']
UNITS = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9}


def clean(name):
    return name.replace('(int)', '').replace('(bool)', '').replace('ncb::', '')


src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
per = collections.OrderedDict()      # kernel name -> list of {dfma, dadd, dmul} per capture
use, hdr, cur = False, None, None
for r in csv.reader(src.splitlines()):
    if r and r[0] == 'Kernel Name':
        use = any(s in r[1] for s in subs)
        hdr = None
        if use:
            cur = dict(dfma=0.0, dadd=0.0, dmul=0.0)
            per.setdefault(clean(r[1]), []).append(cur)
    elif r and r[0] == 'Address':
        hdr = r
    elif use and hdr and len(r) >= len(hdr):
        op = r[hdr.index('Source')].split()
        if not op:
            continue
        nm = (op[1] if op[0].startswith('@') else op[0]).split('.')[0].lower()
        if nm in cur:
            cur[nm] += float(r[hdr.index('Predicated-On Thread Instructions Executed')] or 0)

raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
dram = collections.OrderedDict()
dur = collections.OrderedDict()
for v in rows[2:]:
    name = v[h.index('Kernel Name')]
    b = 0.0
    for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        b += float(v[h.index(m)].replace(',', '')) * UNITS[units[h.index(m)]]
    dram.setdefault(clean(name), []).append(b)
    i = h.index('gpu__time_duration.sum')
    dur.setdefault(clean(name), []).append(float(v[i].replace(',', '')) * UNITS[units[i]] * 1e6)

tot = dict(dfma=0.0, dadd=0.0, dmul=0.0)
kernels = {}
for name, caps in per.items():
    mean = {k: sum(c[k] for c in caps) / len(caps) for k in tot}
    for k in tot:
        tot[k] += mean[k]
    kernels[name] = {"captures": len(caps),
                     "fp64_thread_ops_per_position": {k: v / positions for k, v in mean.items()},
                     "dram_bytes_per_launch": sum(dram.get(name, [0])) / max(len(dram.get(name, [])), 1),
                     "us_per_launch_under_ncu": sum(dur.get(name, [0])) / max(len(dur.get(name, [])), 1)}
source = f"{os.path.basename(rep)} (ncu --set full --clock-control none, {positions} positions per launch; profiles/fp64_ops.py)"
out = {f'{k}_per_position': v / positions for k, v in tot.items()}
out.update(positions=positions, kernels=kernels, source=source)
json.dump(out, open(os.path.join(HERE, 'fp64_ops.json'), 'w'), indent=1)
total_dram = sum(k["dram_bytes_per_launch"] for k in kernels.values())
# every kernel of the capture, keyed by its short name (bench.py looks its roofline candidates up here)
traffic = {"kernel": "mixture kernels of a batch: " + ' + '.join(kernels), "dram_bytes_per_launch": total_dram,
           "positions": positions, "dram_bytes_per_position": total_dram / positions,
           "per_kernel": {k.split('(')[0].replace('void ', ''): sum(v) / len(v) for k, v in dram.items()},
           "us_per_launch_under_ncu": {k.split('(')[0].replace('void ', ''): sum(v) / len(v) for k, v in dur.items()},
           "source": source}
json.dump(traffic, open(os.path.join(HERE, 'traffic.json'), 'w'), indent=1)
print(json.dumps(out, indent=1))
print(json.dumps(traffic, indent=1))
