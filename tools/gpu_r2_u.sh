#!/bin/bash
# round 2, run U: the final build with K1c (4 / 8 warps): full GPU suite + smoke, sanitizers, ncu of K1c, latency sweep, both bench arms
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/u_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/u_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/u_smoke.txt
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_r2.py > gpurun_out/u_$tool.log 2>&1; echo "$tool rc=$?"; tail -1 gpurun_out/u_$tool.log
done
LIB=neuralplane_b200/_lib/libnplane.so
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
for n in 3000 18944; do
  NW=$([ $n -le 9472 ] && echo 8 || echo 4)
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:f16_step_coop -s 5 -c 1 -f -o gpurun_out/prof_k1c_$n python gpurun_out/prof_k1c.py $n > gpurun_out/u_ncu_k1c_$n.log 2>&1; tail -1 gpurun_out/u_ncu_k1c_$n.log
  python tools/ncu_summary.py gpurun_out/prof_k1c_$n.ncu-rep gpurun_out/r02_coop_step_kernel_ncu_full_n$n.txt "ncu --set full --import-source on --clock-control none -k regex:f16_step_coop -s 5 -c 1   [final K1c = f16_step_coop_kernel<HEADING, $NW warps>, n = $n; cold caches and serialised launches under ncu: 11.6 / 16.8 us free-running]" > /dev/null
  python tools/ncu_regions.py gpurun_out/prof_k1c_$n.ncu-rep 512 >> gpurun_out/r02_coop_step_kernel_ncu_full_n$n.txt 2>&1
  TOP=40 python tools/ncu_lines.py gpurun_out/prof_k1c_$n.ncu-rep $LIB f16_step_coop_kernelILi0ELi$NW $n >> gpurun_out/r02_coop_step_kernel_ncu_full_n$n.txt 2>&1
  rm -f gpurun_out/prof_k1c_$n.ncu-rep
done
python tools/small_n_sweep.py > gpurun_out/u_small_n_sweep.txt 2>&1; grep -v "^{" gpurun_out/u_small_n_sweep.txt | cut -c1-230
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default_1gpu.json 2> gpurun_out/u_bench.err; tail -c 1500 gpurun_out/r02_bench_default_1gpu.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/u_bench_ref.err; tail -c 600 gpurun_out/r02_bench_reference_arm.json
du -sh gpurun_out
