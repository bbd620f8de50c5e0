#!/bin/bash
for v in neuralplane_b200/_lib/var/*.so; do
NPLANE_LIB=$PWD/$v python - <<PY
import torch, json
from neuralplane_b200 import ControlEnv
for n in (8_000_000,):
    env=ControlEnv(num_envs=n, config="control", model="UAV", random_seed=0, device="cuda:0"); env.reset()
    a=[torch.rand((n,4),device="cuda")*2-1 for _ in range(2)]
    for k in range(10): env.step(a[k%2])
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    K=100
    for k in range(K): env.step(a[k%2])
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/K
    print("$v", n, "%.4f ms"%ms, "%.3e a-s/s"%(n/ms*1e3), "%.0f GB/s"%(268*n/ms/1e6), "frac %.3f"%(268*n/ms/1e6/6458.4), env.launch_info())
    del env
PY
done
