#!/bin/bash
mkdir -p gpurun_out
python tools/bench_workloads.py > gpurun_out/workloads.jsonl 2> gpurun_out/workloads.err
cat gpurun_out/workloads.jsonl; tail -3 gpurun_out/workloads.err
cat > /tmp/uavprof.py <<'PY'
import torch
from neuralplane_b200 import ControlEnv
n=8_000_000
env=ControlEnv(num_envs=n, config="control", model="UAV", random_seed=0, device="cuda:0"); env.reset()
a=torch.rand((n,4),device="cuda")*2-1
for k in range(6): env.step(a)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:uav_env -s 3 -c 2 -o gpurun_out/prof_uav_r1 python /tmp/uavprof.py > gpurun_out/ncu_uav.log 2>&1; tail -2 gpurun_out/ncu_uav.log
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
