#!/bin/bash
for b in 384 513 385 512 513; do
  NPLANE_LIB=$PWD/neuralplane_b200/_lib/libnplane_all.so NPLANE_BLOCK=$b timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --e2e-steps 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_b.json")); print($b, "%.4g"%d["value"], "%.4f ms"%d["ms_per_step"], d["config"]["launch"])
except Exception as e: print($b, "failed", e, open("gpurun_out/bench_b.err").read()[-500:])
PY
done
NPLANE_LIB=$PWD/neuralplane_b200/_lib/libnplane_all.so NPLANE_BLOCK=513 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
