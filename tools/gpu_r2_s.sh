#!/bin/bash
# round 2, run S: compute-sanitizer (memcheck / racecheck / synccheck) over the round-2 kernels incl. K1c, and an ncu capture of K1c
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_r2.py > gpurun_out/s_$tool.log 2>&1; echo "$tool rc=$?"; tail -2 gpurun_out/s_$tool.log
done
cat > gpurun_out/prof_k1c.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from neuralplane_b200 import ControlEnv
n = int(sys.argv[1])
env = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device="cuda:0")
env.reset()
a = torch.rand((n, 4), device="cuda") * 2 - 1
for k in range(8): env.step(a)
torch.cuda.synchronize()
PY
LIB=neuralplane_b200/_lib/libnplane.so
for n in 3000 18944; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:f16_step_coop -s 5 -c 1 -f -o gpurun_out/prof_k1c_$n python gpurun_out/prof_k1c.py $n > gpurun_out/s_ncu_k1c_$n.log 2>&1; tail -1 gpurun_out/s_ncu_k1c_$n.log
  python tools/ncu_summary.py gpurun_out/prof_k1c_$n.ncu-rep gpurun_out/r02_coop_step_kernel_ncu_full_n$n.txt "ncu --set full --import-source on --clock-control none -k regex:f16_step_coop -s 5 -c 1   [K1c = f16_step_coop_kernel<HEADING>, n = $n]" > /dev/null
  python tools/ncu_regions.py gpurun_out/prof_k1c_$n.ncu-rep 512 >> gpurun_out/r02_coop_step_kernel_ncu_full_n$n.txt 2>&1
  TOP=40 python tools/ncu_lines.py gpurun_out/prof_k1c_$n.ncu-rep $LIB f16_step_coop_kernelILi0 $n >> gpurun_out/r02_coop_step_kernel_ncu_full_n$n.txt 2>&1
  rm -f gpurun_out/prof_k1c_$n.ncu-rep
done
head -60 gpurun_out/r02_coop_step_kernel_ncu_full_n3000.txt
