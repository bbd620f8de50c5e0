#!/bin/bash
# refresh the side-workload numbers with the current build (UAV slab kernel, divc, cheaper noise) + ncu of the table env
mkdir -p gpurun_out
python tools/bench_workloads.py > gpurun_out/side_workloads_r01e.jsonl 2>gpurun_out/side.err; cat gpurun_out/side_workloads_r01e.jsonl
cat > gpurun_out/tab_bench.py <<'PY'
import sys, os, json
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from neuralplane_b200 import ControlEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 300
env = ControlEnv(num_envs=n, config="heading", model="F16_tables", random_seed=0, device="cuda:0"); env.reset()
a = [torch.rand((n, 4), device="cuda") * 2 - 1 for _ in range(4)]
for k in range(30): env.step(a[k % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
for k in range(K): env.step(a[k % 4])
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / K
print(json.dumps({"workload": "F16 heading, TABLE aero back-end, n=%d" % n, "ms_per_step": ms, "aircraft_steps_per_s": n / ms * 1e3,
                  "achieved_GBps": 276 * n / ms / 1e6, "frac_hbm": 276 * n / ms / 1e6 / 6458.4, "launch": env.launch_info()}))
PY
python gpurun_out/tab_bench.py 1000000 300 | tee gpurun_out/table_env_r01e.jsonl
python gpurun_out/tab_bench.py 4000000 100 | tee -a gpurun_out/table_env_r01e.jsonl
rm -f gpurun_out/prof_tab_r01e.ncu-rep
timeout 600 ncu --set full --import-source on --clock-control none -k regex:f16_step -s 20 -c 1 -f -o gpurun_out/prof_tab_r01e python gpurun_out/tab_bench.py 4000000 5 > gpurun_out/ncu_tab.log 2>&1; tail -1 gpurun_out/ncu_tab.log
python tools/bench_rollout.py > gpurun_out/rollout_r01e.jsonl 2>>gpurun_out/side.err; cat gpurun_out/rollout_r01e.jsonl
