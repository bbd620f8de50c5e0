#!/bin/bash
# parity tests + bench + block sweep for the v2 step kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for b in 128 256 384 512; do
  NPLANE_LIB=$PWD/neuralplane_b200/_lib/libnplane_all.so NPLANE_BLOCK=$b timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --e2e-steps 3 > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
done
tail -n 5 gpurun_out/smoke.log; tail -n 30 gpurun_out/pytest_gpu.log
for b in 128 256 384 512; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_b$b.json")); print($b, d["value"], d["ms_per_step"], d["config"]["launch"])
except Exception as e: print($b, "failed", e, open("gpurun_out/bench_b$b.err").read()[-500:])
PY
done
