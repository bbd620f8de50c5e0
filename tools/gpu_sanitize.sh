#!/bin/bash
# compute-sanitizer memcheck over the ragged-size / tail-lane / multi-mode tests (small populations)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "ragged or cache or graph or gpuvecenv" > gpurun_out/sanitize_parity.log 2>&1; echo "memcheck parity rc=$?"
tail -4 gpurun_out/sanitize_parity.log
cat > gpurun_out/san_small.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from neuralplane_b200 import ControlEnv, PlanningEnv, SingleCombatEnv
for n in (1, 3, 257, 1000):
    for model, task in (("F16", "heading"), ("F16", "tracking"), ("UAV", "control")):
        e = ControlEnv(num_envs=n, config=task, model=model, random_seed=1, device="cuda:0"); e.reset()
        for k in range(3): e.step(torch.rand((n, 4), device="cuda") * 2 - 1)
        e.model.get_extended_state(); torch.cuda.synchronize()
    e = PlanningEnv(num_envs=n, config="tracking", random_seed=1, device="cuda:0", n_substeps=4); e.reset()
    for k in range(2): e.step(torch.rand((n, 3), device="cuda") * 2 - 1)
    e = SingleCombatEnv(num_envs=n, config="selfplay", random_seed=1, device="cuda:0"); e.reset()
    for k in range(2): e.step(torch.rand((e.n, 4), device="cuda") * 2 - 1)
    torch.cuda.synchronize()
print("san_small done")
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python gpurun_out/san_small.py > gpurun_out/sanitize_small.log 2>&1; echo "memcheck small rc=$?"; tail -4 gpurun_out/sanitize_small.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python gpurun_out/san_small.py > gpurun_out/racecheck_small.log 2>&1; echo "racecheck small rc=$?"; tail -4 gpurun_out/racecheck_small.log
