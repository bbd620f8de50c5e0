#!/bin/bash
# compute-sanitizer over the code added late in round 1: the TMA-staged UAV slab kernel and the native host boundary
mkdir -p gpurun_out
cat > gpurun_out/san_new.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
from neuralplane_b200 import ControlEnv, GPUVecEnv
for n in (256, 1000, 5000 + 37):
    for task in ("control", "heading", "tracking"):
        e = ControlEnv(num_envs=n, config=task, model="UAV", random_seed=1, device="cuda:0"); e.reset()
        for k in range(4): e.step(torch.rand((n, 4), device="cuda") * 2 - 1)
        assert e.launch_info()["smem_bytes"] > 40000
        torch.cuda.synchronize()
for model, task in (("F16", "heading"), ("UAV", "control"), ("F16_tables", "heading")):
    n = 3000 + 11
    v = GPUVecEnv([lambda: ControlEnv(num_envs=n, config=task, model=model, random_seed=2, device="cuda:0")], pipeline_chunks=(1, 2, 3))
    v.reset()
    for k in range(3): out = v.step(np.random.rand(n, 1, 4).astype(np.float32) * 2 - 1)
    assert np.isfinite(out[0]).all()
print("san_new done")
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python gpurun_out/san_new.py > gpurun_out/sanitize_new.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitize_new.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python gpurun_out/san_new.py > gpurun_out/racecheck_new.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck_new.log
timeout 1200 compute-sanitizer --tool synccheck --error-exitcode 9 python gpurun_out/san_new.py > gpurun_out/synccheck_new.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/synccheck_new.log
