#!/bin/bash
# First GPU visit: smoke, parity tests, bench at several block sizes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for b in 128 256 384 512; do
  NPLANE_LIB=$PWD/neuralplane_b200/_lib/libnplane_all.so NPLANE_BLOCK=$b timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --e2e-steps 3 > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
done
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --e2e-steps 3 --no-cache > gpurun_out/bench_nocache.json 2> gpurun_out/bench_nocache.err
tail -n 5 gpurun_out/smoke.log; tail -n 15 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_*.json
