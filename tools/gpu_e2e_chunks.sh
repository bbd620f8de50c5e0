#!/bin/bash
# numpy boundary: component times and a sweep of pipeline chunk patterns (equal and graded)
mkdir -p gpurun_out
python - <<'PY' | tee gpurun_out/e2e_chunks.txt
import time, numpy as np, torch
from neuralplane_b200 import ControlEnv, GPUVecEnv
n=1_000_000
# components
a=np.random.rand(n,4).astype(np.float32); ph=torch.empty((n,4)).pin_memory(); d=torch.empty((n,4),device="cuda")
obs_d=torch.empty((n,22),device="cuda"); obs_h=torch.empty((n,22)).pin_memory()
def t(f,K=20):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(K): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/K*1e3
print("host memcpy 16 MB numpy->pinned: %.3f ms"%t(lambda: ph.copy_(torch.from_numpy(a))))
print("H2D 16 MB: %.3f ms"%t(lambda: d.copy_(ph,non_blocking=True)))
print("D2H 88 MB: %.3f ms"%t(lambda: obs_h.copy_(obs_d,non_blocking=True)))
print("torch threads", torch.get_num_threads())
pats=[4,6,(1,2,3,4,4,4),(1,2,4,8,8,8,8),(1,2,3,4,5,5,5,5,5),(1,2,4,8,16,16,16)]
res={}
for rep in range(2):
  for k in pats:
    v=GPUVecEnv([lambda: ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device="cuda:0")], pipeline_chunks=k)
    v.reset()
    acts=[(np.random.rand(n,1,4).astype(np.float32)*2-1) for _ in range(2)]
    for i in range(3): v.step(acts[i%2])
    torch.cuda.synchronize(); t0=time.perf_counter()
    K=30
    for i in range(K): v.step(acts[i%2])
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/K
    res.setdefault(str(k),[]).append(dt*1e3)
    del v
for k,vv in res.items(): print("chunks %-22s: %s ms/step  best e2e %.3e a-s/s"%(k, " ".join("%.3f"%x for x in vv), n/min(vv)*1e3), flush=True)
PY
