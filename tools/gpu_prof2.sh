#!/bin/bash
# ncu captures of the planning / combat kernels + UAV with and without observation noise
mkdir -p gpurun_out
cat > gpurun_out/prof_modes.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from neuralplane_b200 import PlanningEnv, SingleCombatEnv
n = 400_000
e = PlanningEnv(num_envs=n, config="tracking", random_seed=0, device="cuda:0", n_substeps=50); e.reset()
for k in range(2): e.step(torch.rand((n, 3), device="cuda") * 2 - 1)
c = SingleCombatEnv(num_envs=n // 2, config="selfplay", random_seed=0, device="cuda:0"); c.reset()
for k in range(3): c.step(torch.rand((c.n, 4), device="cuda") * 2 - 1)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none -k regex:f16_step -c 5 -o gpurun_out/prof_modes_r1 python gpurun_out/prof_modes.py > gpurun_out/ncu_modes.log 2>&1; tail -1 gpurun_out/ncu_modes.log
python - <<'PY'
import torch
from neuralplane_b200 import ControlEnv
n=8_000_000
for ns in (0.01, 0.0):
    env=ControlEnv(num_envs=n, config="control", model="UAV", random_seed=0, device="cuda:0"); env.task.noise_scale=ns; env.reset()
    a=[torch.rand((n,4),device="cuda")*2-1 for _ in range(2)]
    for k in range(10): env.step(a[k%2])
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for k in range(100): env.step(a[k%2])
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/100
    print("UAV n=8e6 noise_scale=%g: %.4f ms  %.3e a-s/s  %.0f GB/s  frac %.3f"%(ns, ms, n/ms*1e3, 268*n/ms/1e6, 268*n/ms/1e6/6458.4))
    del env
PY
