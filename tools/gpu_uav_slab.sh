#!/bin/bash
# UAV step: TMA-staged slab kernel vs the per-thread kernel (NPLANE_UAV_SCALAR=1), parity tests first.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_uav.py -q -m gpu -x > gpurun_out/pytest_uav.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_uav.log
tail -n 15 gpurun_out/pytest_uav.log
for mode in slab scalar; do
if [ $mode = scalar ]; then export NPLANE_UAV_SCALAR=1; else unset NPLANE_UAV_SCALAR; fi
timeout 600 python - <<PY | tee -a gpurun_out/uav_slab.jsonl
import torch, json
from neuralplane_b200 import ControlEnv
for n in (1_000_000, 8_000_000):
  for noise in (0.01, 0.0):
    env=ControlEnv(num_envs=n, config="control", model="UAV", random_seed=0, device="cuda:0"); env.task.noise_scale=noise; env.reset()
    a=[torch.rand((n,4),device="cuda")*2-1 for _ in range(2)]
    for k in range(20): env.step(a[k%2])
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    K=200
    for k in range(K): env.step(a[k%2])
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/K
    print(json.dumps({"mode":"$mode","n":n,"noise_scale":noise,"ms_per_step":round(ms,4),"aircraft_steps_per_s":n/ms*1e3,"GBps":268*n/ms/1e6,"frac":268*n/ms/1e6/6458.4,"launch":env.launch_info()}))
    del env
PY
done
