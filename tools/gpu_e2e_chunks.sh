#!/bin/bash
python - <<'PY'
import time, numpy as np, torch
from neuralplane_b200 import ControlEnv, GPUVecEnv
n=1_000_000
for k in (1,2,4,6,8,12,16):
    v=GPUVecEnv([lambda: ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device="cuda:0")], pipeline_chunks=k)
    v.reset()
    acts=[(np.random.rand(n,1,4).astype(np.float32)*2-1) for _ in range(2)]
    for i in range(3): v.step(acts[i%2])
    torch.cuda.synchronize(); t0=time.perf_counter()
    K=20
    for i in range(K): v.step(acts[i%2])
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/K
    print("chunks %2d: %.3f ms/step  e2e %.3e a-s/s"%(k, dt*1e3, n/dt), flush=True)
    del v
PY
