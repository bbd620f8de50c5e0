#!/usr/bin/env python
"""Per-region stall breakdown of a kernel from an .ncu-rep (SASS source page): python tools/ncu_regions.py rep [step]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; step = int(sys.argv[2]) if len(sys.argv) > 2 else 512
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr): break
    data.append(r)
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
ti = sum(f(r, 'Instructions Executed') for r in data); ts = sum(f(r, '# Samples') for r in data)
print(rows[0][1][:80], "static", len(data), "dyn", ti, "samples", ts)
keys = ['stall_no_inst', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_mio', 'stall_lg', 'stall_branch_resolving', 'stall_dispatch', 'stall_not_selected', 'stall_selected', 'stall_barrier']
print("total  " + "  ".join("%s %.1f%%" % (k[6:], 100 * sum(f(r, k) for r in data) / ts) for k in keys))
for s in range(0, len(data), step):
    blk = data[s:s + step]
    di = sum(f(r, 'Instructions Executed') for r in blk); ds = sum(f(r, '# Samples') for r in blk)
    ff = sum(1 for r in blk if 'FFMA2' in r[idx['Source']])
    print("instr %5d-%5d FFMA2 %4d dyn %5.1f%% smp %5.1f%% | " % (s, s + step, ff, 100 * di / ti, 100 * ds / ts) +
          " ".join("%s %.1f" % (k[6:10], 100 * sum(f(r, k) for r in blk) / ts) for k in keys[:9]))
