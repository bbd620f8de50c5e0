#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/uav_slab.jsonl
bash tools/gpu_uav_slab.sh 2>&1 | tail -n 10
bash tools/gpu_sanitize2.sh 2>&1 | grep "rc="
bash tools/gpu_prof_uav.sh
