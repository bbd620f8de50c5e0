#!/bin/bash
# table-backed env: parity tests, block-size sweep (libnplane_tab.so built with -DNPLANE_TAB_BLOCKS), ncu of the kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tables_env.py -q -x -s 2>&1 | tail -40
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
print(json.dumps({"workload": "F16 heading, TABLE aero back-end, n=%d" % n, "tab_block": os.environ.get("NPLANE_TAB_BLOCK", "256"),
                  "ms_per_step": ms, "aircraft_steps_per_s": n / ms * 1e3, "achieved_GBps": 276 * n / ms / 1e6,
                  "frac_hbm": 276 * n / ms / 1e6 / 6458.4, "launch": env.launch_info(), "resets": env.termination_counters()["resets"]}))
PY
for b in 128 256 384 512; do
  NPLANE_LIB=$PWD/neuralplane_b200/_lib/libnplane_tab.so NPLANE_TAB_BLOCK=$b timeout 300 python gpurun_out/tab_bench.py 2>&1 | tail -1
done
python gpurun_out/tab_bench.py 4000000 100 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none -k regex:f16_step -s 20 -c 2 -o gpurun_out/prof_tab_r1 python gpurun_out/tab_bench.py 1000000 5 > gpurun_out/ncu_tab.log 2>&1; tail -1 gpurun_out/ncu_tab.log
python bench.py --steps 300 --warmup 20 --no-cpu --e2e-steps 3 2>/dev/null | tail -1
