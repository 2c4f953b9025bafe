"""
Turns ncu output brought back in gpurun_out/ into the small text summaries committed here.

    python profiles/summarise.py <tag> <launches.csv> <report.ncu-rep> [bench.json]

Writes profiles/<tag>_launches.md (per-kernel share of the step, from the
`--metrics gpu__time_duration.sum --clock-control none` pass) and profiles/<tag>_<kernel>.md
(key metrics + opcode / stall breakdown of the `--set full` capture).
"""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def launches(tag, path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1e-3)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    out = [f"# {tag}: kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none)", "",
           "Cold-cache, serialised launches: compare shares, not absolutes.", "",
           "| kernel | launches | us / launch | share |", "|---|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{k[:110]}` | {a[0]} | {a[1] / a[0]:.1f} | {100 * a[1] / tot:.1f} % |")
    open(os.path.join(HERE, f'{tag}_launches.md'), 'w').write('\n'.join(out) + '\n')


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def full(tag, rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, v = rows[0], rows[1], rows[-1]
    name = v[h.index('Kernel Name')]
    out = [f"# {tag}: `{name[:100]}` (ncu --set full --clock-control none)", "", "| metric | value | unit |", "|---|---:|---|"]
    for w in WANT:
        if w in h:
            i = h.index(w)
            out.append(f"| {w} | {v[i]} | {units[i]} |")
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hh = rows[1]
    si, ii, wi = hh.index('Source'), hh.index('Instructions Executed'), hh.index('Warp Stall Sampling (All Samples)')
    stall_cols = [i for i, n in enumerate(hh) if n.startswith('stall_') and 'Not Issued' not in n]
    ops, samp, stalls = collections.Counter(), collections.Counter(), collections.Counter()
    tot = tots = 0.0
    nsass = 0
    for r in rows[2:]:
        if len(r) < len(hh):
            continue
        op = r[si].split()
        if not op:
            continue
        nsass += 1
        nm = (op[1] if op[0].startswith('@') else op[0]).split('.')[0]
        n, s = float(r[ii] or 0), float(r[wi] or 0)
        ops[nm] += n
        samp[nm] += s
        tot += n
        tots += s
        for c in stall_cols:
            stalls[hh[c]] += float(r[c] or 0)
    out += ["", f"SASS instructions in the kernel: {nsass}; warp instructions executed: {tot:.0f}", "",
            "| opcode | % of executed | % of stall samples |", "|---|---:|---:|"]
    for k, n in ops.most_common(14):
        out.append(f"| {k} | {100 * n / tot:.1f} | {100 * samp[k] / max(tots, 1):.1f} |")
    out += ["", "| stall reason | % of samples |", "|---|---:|"]
    for k, n in stalls.most_common(8):
        out.append(f"| {k} | {100 * n / max(tots, 1):.1f} |")
    short = name.split('(')[0].split('::')[-1].replace('<', '_').replace('>', '').replace(', ', '_').replace(' ', '')
    open(os.path.join(HERE, f'{tag}_{short}.md'), 'w').write('\n'.join(out) + '\n')
    # per-launch DRAM traffic for bench.py's roofline.traffic
    try:
        rd = float(v[h.index('dram__bytes_read.sum')]) * {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1}[units[h.index('dram__bytes_read.sum')]]
        wr = float(v[h.index('dram__bytes_write.sum')]) * {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1}[units[h.index('dram__bytes_write.sum')]]
        return name, rd + wr
    except Exception:
        return name, None


if __name__ == '__main__':
    tag = sys.argv[1]
    launches(tag, sys.argv[2])
    name, traffic = full(tag, sys.argv[3])
    print(name, traffic)
