#!/bin/bash
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -q -m gpu -s "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|Error|assert|k= *(1|10|100|300|1000) |vs fp64" gpurun_out/pytest_gpu.log | tail -n 60
