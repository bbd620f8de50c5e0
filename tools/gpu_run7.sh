#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/combat_exchange_check.py > gpurun_out/combat_exchange_2gpu.json 2> gpurun_out/combat_exchange_2gpu.err; echo "exchange rc=$?"; cat gpurun_out/combat_exchange_2gpu.json; tail -3 gpurun_out/combat_exchange_2gpu.err
python - <<'PY'
import torch, json
from neuralplane_b200 import SingleCombatEnv
ne=500_000
env=SingleCombatEnv(num_envs=ne, config="selfplay", random_seed=0, device="cuda:0"); env.reset()
a=[torch.rand((env.n,4),device="cuda")*2-1 for _ in range(4)]
for k in range(3): env.step(a[k])
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
K=20
for k in range(K): env.step(a[k%4])
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/K
print(json.dumps({"workload":"SingleCombat 1v1, 5e5 envs x 2 aircraft, pair-sharded, 5 sub-steps per env step","ms_per_env_step":ms,"aircraft_fdm_steps_per_s":env.n*5/(ms*1e-3),"counters":env.termination_counters()}))
PY
