"""Joins an ncu SASS-level source page (csv) with nvdisasm -g line info: executed warp
instructions and stall samples per CUDA source line.

    python profiles/by_line.py <src.csv> <disasm> <mangled-kernel-substring> [top]
"""
import collections
import csv
import re
import sys

src_csv, disasm, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
# line info: offset -> (file, line, inlined-at chain root)
lines = {}
cur = None
infn = False
for ln in open(disasm):
    if ln.startswith('.text.'):
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), m.group(3))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(\S.*?);', ln)
    if m and cur:
        lines[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ai, si, ii, wi = h.index('Address'), h.index('Source'), h.index('Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
data = [(int(r[ai], 16), float(r[ii] or 0), float(r[wi] or 0)) for r in rows[2:] if len(r) >= len(h)]
base = data[0][0]
agg = collections.defaultdict(lambda: [0.0, 0.0])
tot = tots = 0.0
for a, n, s in data:
    key = lines.get(a - base, ('?', 0, ''))[:2]
    agg[key][0] += n
    agg[key][1] += s
    tot += n
    tots += s
print(f"total warp instructions {tot:.0f}, stall samples {tots:.0f}")
for k, (n, s) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{100 * n / tot:5.1f}% inst {100 * s / max(tots, 1):5.1f}% samples  {k[0]}:{k[1]}")
