#!/bin/bash
mkdir -p gpurun_out
for v in neuralplane_b200/_lib/var/*.so neuralplane_b200/_lib/var/*.so; do
  NPLANE_LIB=$PWD/$v timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu --e2e-steps 3 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_var.json")); print("$v", "%.4g"%d["value"], "%.4f ms"%d["ms_per_step"], d["config"]["launch"])
except Exception as e: print("$v", "failed", e, open("gpurun_out/bench_var.err").read()[-800:])
PY
done
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
