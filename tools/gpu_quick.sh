#!/bin/bash
# GPU parity tests + a short bench (the usual check after a kernel change)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --e2e-steps 5 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_quick.json")); print("%.4g a-s/s"%d["value"], "%.4f ms"%d["ms_per_step"], d["config"]["launch"], "e2e %.3g"%d["e2e"]["value"], "fma %.3f hbm %.4f"%(d["roofline"]["fma_pipe_frac"], d["roofline"]["frac"]))
except Exception as e: print("failed", e, open("gpurun_out/bench_quick.err").read()[-800:])
PY
done
