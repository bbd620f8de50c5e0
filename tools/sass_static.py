#!/usr/bin/env python
"""Static SASS size of one kernel by CUDA source line (nvdisasm -g line table): where the code bytes are.
python tools/sass_static.py libnplane.so kernel_substring [top]"""
import collections, os, re, subprocess, sys, tempfile
lib, kern = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
per, cur, on, done = collections.Counter(), ("?", 0), False, False
ops = collections.Counter()
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        on = (kern in m.group(1)) and not done
        if on:
            done = True
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m:
        per[cur] += 1
        ops[m.group(1).split(".")[0]] += 1
tot = sum(per.values())
print("static instructions", tot, "=", tot * 16 // 1024, "KB")
print("opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
byfile = collections.Counter()
for (f, _), v in per.items():
    byfile[f] += v
print({k: v for k, v in byfile.most_common(8)})
for k, v in per.most_common(top):
    src = ""
    p = os.path.join("neuralplane_b200/csrc", k[0])
    if os.path.exists(p):
        L = open(p).read().splitlines()
        if 0 < k[1] <= len(L):
            src = L[k[1] - 1].strip()[:100]
    print("%-20s %5d %6d  %4.1f%% | %s" % (k[0], k[1], v, 100 * v / tot, src))
