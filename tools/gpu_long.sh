#!/bin/bash
# sustained-load check: a long timed region so the clock sampler sees the steady state
timeout 600 python bench.py --steps 4000 --warmup 50 --no-cpu --e2e-steps 5 > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_long.json")); print("%.4g a-s/s"%d["value"], "%.4f ms"%d["ms_per_step"], d["clocks"], "e2e %.3g"%d["e2e"]["value"])
PY
