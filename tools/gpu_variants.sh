#!/bin/bash
mkdir -p gpurun_out
for v in neuralplane_b200/_lib/var/*.so; do
 for b in ${BLOCKS:-256}; do
  NPLANE_BLOCK=$b NPLANE_LIB=$PWD/$v timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --e2e-steps 3 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_var.json")); print("$v", "%.4g"%d["value"], "%.4f ms"%d["ms_per_step"], d["config"]["launch"])
except Exception as e: print("$v", "failed", e, open("gpurun_out/bench_var.err").read()[-800:])
PY
 done
done
