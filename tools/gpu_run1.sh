#!/bin/bash
# GPU visit: smoke, parity tests, bench, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --e2e-steps 3 --no-cache > gpurun_out/bench_nocache.json 2> gpurun_out/bench_nocache.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/ncu_bench.log 2>&1
tail -n 5 gpurun_out/smoke.log; tail -n 15 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_*.json
