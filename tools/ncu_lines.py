#!/usr/bin/env python
"""Dynamic instruction count per CUDA source line of one kernel: joins the SASS source page of an .ncu-rep with the
line table of the same cubin (nvdisasm -g).  python tools/ncu_lines.py rep.ncu-rep libnplane.so kernel_substring [units]
`units` = aircraft processed by the launch (prints thread-instructions per aircraft)."""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
sec, lines, cur, on = None, [], ("?", 0), False
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        on = kern in m.group(1) and not lines
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s", l):
        lines.append((cur, l))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
assert len(data) == len(lines), (len(data), len(lines))
per, smp = collections.Counter(), collections.Counter()
for (c, _), r in zip(lines, data):
    per[c] += float(r[ix["Instructions Executed"]]); smp[c] += float(r[ix["# Samples"]])
tot, ts = sum(per.values()), sum(smp.values())
print("static", len(data), "dynamic warp-instr", tot, ("thread-instr/unit %.0f" % (tot * 32 / units)) if units else "")
byfile = collections.Counter()
for (f, _), v in per.items(): byfile[f] += v
print({k: "%.1f%%" % (100 * v / tot) for k, v in byfile.most_common(8)})
for k, v in per.most_common(int(os.environ.get("TOP", "60"))):
    src = ""
    for d in ("neuralplane_b200/csrc",):
        p = os.path.join(d, k[0])
        if os.path.exists(p):
            L = open(p).read().splitlines()
            if 0 < k[1] <= len(L): src = L[k[1] - 1].strip()[:90]
    print("%-18s %5d  %5.1f%% instr %5.1f%% smp %s| %s" % (k[0], k[1], 100 * v / tot, 100 * smp[k] / ts, ("%6.0f/unit " % (v * 32 / units)) if units else "", src))
