#!/bin/bash
# 2-GPU: bench under torchrun (nccl), reference arm under torchrun, UAV ncu capture on GPU 0
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; cat gpurun_out/bench_2gpu_ref.json
cat > gpurun_out/uavprof.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from neuralplane_b200 import ControlEnv
n=8_000_000
env=ControlEnv(num_envs=n, config="control", model="UAV", random_seed=0, device="cuda:0"); env.reset()
a=torch.rand((n,4),device="cuda")*2-1
for k in range(6): env.step(a)
torch.cuda.synchronize()
PY
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:uav_env -s 3 -c 2 -o gpurun_out/prof_uav_r1 python gpurun_out/uavprof.py > gpurun_out/ncu_uav.log 2>&1; tail -2 gpurun_out/ncu_uav.log
