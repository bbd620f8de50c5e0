#!/bin/bash
# ncu full capture of the step kernel (name of the report as $1)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:f16_step -s 5 -c 2 -o gpurun_out/$1 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
