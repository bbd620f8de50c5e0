#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
for b in ${BLOCKS:-256 384}; do
  NPLANE_LIB=$PWD/neuralplane_b200/_lib/libnplane_all.so NPLANE_BLOCK=$b timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --e2e-steps 3 > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_b$b.json")); print($b, "%.4g"%d["value"], d["ms_per_step"], d["config"]["launch"], "e2e %.3g"%d["e2e"]["value"])
except Exception as e: print($b, "failed", e, open("gpurun_out/bench_b$b.err").read()[-500:])
PY
done
