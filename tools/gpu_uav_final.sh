#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/uav_slab.jsonl
bash tools/gpu_uav_slab.sh 2>&1 | tail -n 10
bash tools/gpu_sanitize2.sh 2>&1 | grep "rc="
rm -f gpurun_out/prof_uav_slab.ncu-rep
timeout 600 ncu --set full --import-source on --clock-control none -k regex:uav_step_slab -s 3 -c 1 -f -o gpurun_out/prof_uav_slab python gpurun_out/prof_uav.py > gpurun_out/ncu_uav.log 2>&1; tail -1 gpurun_out/ncu_uav.log
