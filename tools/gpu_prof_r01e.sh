#!/bin/bash
# ncu --set full (+source) captures: K1 at n = 10^6 through bench.py, and the UAV slab kernel at n = 8e6
mkdir -p gpurun_out
rm -f gpurun_out/prof_k1_r01e.ncu-rep gpurun_out/prof_uav_slab.ncu-rep
timeout 600 ncu --set full --import-source on --clock-control none -k regex:f16_step -s 5 -c 1 -f -o gpurun_out/prof_k1_r01e python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_k1.log 2>&1; tail -1 gpurun_out/ncu_k1.log
cat > gpurun_out/prof_uav.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from neuralplane_b200 import ControlEnv
n = 8_000_000
env = ControlEnv(num_envs=n, config="control", model="UAV", random_seed=0, device="cuda:0"); env.reset()
a = torch.rand((n, 4), device="cuda") * 2 - 1
for k in range(6): env.step(a)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:uav_step_slab -s 3 -c 1 -f -o gpurun_out/prof_uav_slab python gpurun_out/prof_uav.py > gpurun_out/ncu_uav.log 2>&1; tail -1 gpurun_out/ncu_uav.log
ls -la gpurun_out/*.ncu-rep
