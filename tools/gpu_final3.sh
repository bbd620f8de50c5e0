#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh 2>&1 | tail -n 3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_steps10.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-side --e2e-steps 3 > gpurun_out/bench_under_ncu.log 2>&1
grep -c "f16_step_kernel" gpurun_out/launches_bench_steps10.csv
