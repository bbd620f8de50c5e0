#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; grep real gpurun_out/bench_default.err
cat gpurun_out/bench_default.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_steps10.csv python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
grep -c "f16_step_kernel" gpurun_out/launches_bench_steps10.csv
