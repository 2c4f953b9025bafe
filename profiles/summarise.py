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
        if 'fp64_probe_kernel' in r[ki]:      # bench.py's DFMA microbenchmark, outside the timed region
            continue
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


UNITS = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}


def short_name(name):
    name = name.replace('(int)', '').replace('(bool)', '')
    base = name.split('(')[0].replace('void ', '').split('::')[-1]
    return base.replace('<', '_').replace('>', '').replace(', ', '_').replace(' ', '')


def full(tag, rep):
    """One summary file per distinct kernel captured in the report (last capture of each)."""
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    kernels = collections.OrderedDict()
    for v in rows[2:]:
        kernels[v[h.index('Kernel Name')]] = v
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    sections = collections.OrderedDict()
    cur = None
    for r in srows:
        if r and r[0] == 'Kernel Name':
            cur = r[1]
            sections[cur] = []
        elif cur is not None:
            sections[cur].append(r)
    out_traffic = {}
    for name, v in kernels.items():
        out = [f"# {tag}: `{name[:110]}` (ncu --set full --clock-control none)", "", "| metric | value | unit |", "|---|---:|---|"]
        for w in WANT + ['sm__icc_request_hit_rate.pct', 'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed',
                         'smsp__warps_eligible.avg.per_cycle_active']:
            if w in h:
                i = h.index(w)
                out.append(f"| {w} | {v[i]} | {units[i]} |")
        sec = None
        for k in sections:
            if short_name(k) == short_name(name):
                sec = sections[k]
        if sec:
            hh = sec[0]
            si, ii, wi = hh.index('Source'), hh.index('Instructions Executed'), hh.index('Warp Stall Sampling (All Samples)')
            stall_cols = [i for i, n in enumerate(hh) if n.startswith('stall_') and 'Not Issued' not in n]
            ops, samp, stalls = collections.Counter(), collections.Counter(), collections.Counter()
            tot = tots = 0.0
            nsass = 0
            for r in sec[1:]:
                if len(r) < len(hh):
                    continue
                op = r[si].split()
                if not op:
                    continue
                nsass += 1
                nm = (op[1] if op[0].startswith('@') else op[0]).split('.')[0]
                n, s_ = float(r[ii] or 0), float(r[wi] or 0)
                ops[nm] += n
                samp[nm] += s_
                tot += n
                tots += s_
                for c in stall_cols:
                    stalls[hh[c]] += float(r[c] or 0)
            out += ["", f"SASS instructions in the kernel: {nsass}; warp instructions executed: {tot:.0f}", "",
                    "| opcode | % of executed | % of stall samples |", "|---|---:|---:|"]
            for k, n in ops.most_common(14):
                out.append(f"| {k} | {100 * n / max(tot, 1):.1f} | {100 * samp[k] / max(tots, 1):.1f} |")
            out += ["", "| stall reason | % of samples |", "|---|---:|"]
            for k, n in stalls.most_common(8):
                out.append(f"| {k} | {100 * n / max(tots, 1):.1f} |")
        open(os.path.join(HERE, f'{tag}_{short_name(name)}.md'), 'w').write('\n'.join(out) + '\n')
        try:
            rd = float(v[h.index('dram__bytes_read.sum')]) * UNITS[units[h.index('dram__bytes_read.sum')]]
            wr = float(v[h.index('dram__bytes_write.sum')]) * UNITS[units[h.index('dram__bytes_write.sum')]]
            out_traffic[name] = rd + wr
        except Exception:
            out_traffic[name] = None
    return out_traffic


if __name__ == '__main__':
    tag = sys.argv[1]
    if sys.argv[2] != '-':
        launches(tag, sys.argv[2])
    for name, traffic in full(tag, sys.argv[3]).items():
        print(traffic, name)
