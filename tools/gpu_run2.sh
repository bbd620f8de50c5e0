#!/bin/bash
mkdir -p gpurun_out
./tools/mlp_bench > gpurun_out/mlp_bench.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/mlp_bench.txt; tail -n 40 gpurun_out/pytest_gpu.log
