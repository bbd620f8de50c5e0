#!/bin/bash
# round 2, call K: ncu full captures of the final build (K1, K1t, UAV slab), summarised ON THE BOX (gpurun brings back <= 64 MiB)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LIB=neuralplane_b200/_lib/libnplane.so
timeout 600 ncu --set full --import-source on --clock-control none -k regex:f16_step_kernel -s 5 -c 1 -f -o gpurun_out/prof_k1_r02 python bench.py --steps 5 --warmup 3 --no-cpu --no-side > gpurun_out/k_ncu_k1.log 2>&1; tail -1 gpurun_out/k_ncu_k1.log
H="ncu --set full --import-source on --clock-control none -k regex:f16_step_kernel -s 5 -c 1 python bench.py --steps 5 --warmup 3 --no-cpu --no-side   [round 2 final build: K1 = f16_step_kernel<384,1,MODE_STEP,false,HEADING>, n = 10^6]"
python tools/ncu_summary.py gpurun_out/prof_k1_r02.ncu-rep gpurun_out/r02_f16_step_kernel_ncu_full.txt "$H" > /dev/null
python tools/ncu_regions.py gpurun_out/prof_k1_r02.ncu-rep 1024 >> gpurun_out/r02_f16_step_kernel_ncu_full.txt 2>&1
echo "" >> gpurun_out/r02_f16_step_kernel_ncu_full.txt; echo "dynamic instructions per CUDA source line (tools/ncu_lines.py; unit = aircraft):" >> gpurun_out/r02_f16_step_kernel_ncu_full.txt
TOP=45 python tools/ncu_lines.py gpurun_out/prof_k1_r02.ncu-rep $LIB f16_step_kernelILi384ELi1ELi0ELb0ELi0 1000000 >> gpurun_out/r02_f16_step_kernel_ncu_full.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r02_launches_bench_steps10.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-side > gpurun_out/k_ncu_launch.log 2>&1
cat > gpurun_out/prof_side.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from neuralplane_b200 import ControlEnv
which = sys.argv[1]
n = 8_000_000 if which == "uav" else 4_000_000
env = ControlEnv(num_envs=n, config="control" if which == "uav" else "heading", model="UAV" if which == "uav" else "F16_tables", random_seed=0, device="cuda:0")
env.reset()
a = torch.rand((n, 4), device="cuda") * 2 - 1
for k in range(6): env.step(a)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:uav_step_slab -s 3 -c 1 -f -o gpurun_out/prof_uav python gpurun_out/prof_side.py uav > gpurun_out/k_ncu_uav.log 2>&1; tail -1 gpurun_out/k_ncu_uav.log
python tools/ncu_summary.py gpurun_out/prof_uav.ncu-rep gpurun_out/r02_uav_slab_kernel_ncu_full.txt "ncu --set full ... -k regex:uav_step_slab -s 3 -c 1   [round 2: uav_step_slab_kernel with the nanosleep back-off, n = 8e6, control task, noise on]" > /dev/null
python tools/ncu_regions.py gpurun_out/prof_uav.ncu-rep 512 >> gpurun_out/r02_uav_slab_kernel_ncu_full.txt 2>&1
TOP=30 python tools/ncu_lines.py gpurun_out/prof_uav.ncu-rep $LIB uav_step_slab_kernel 8000000 >> gpurun_out/r02_uav_slab_kernel_ncu_full.txt 2>&1
rm -f gpurun_out/prof_uav.ncu-rep
timeout 600 ncu --set full --import-source on --clock-control none -k regex:f16_table_step -s 3 -c 1 -f -o gpurun_out/prof_k1t python gpurun_out/prof_side.py tab > gpurun_out/k_ncu_k1t.log 2>&1; tail -1 gpurun_out/k_ncu_k1t.log
python tools/ncu_summary.py gpurun_out/prof_k1t.ncu-rep gpurun_out/r02_table_step_kernel_ncu_full.txt "ncu --set full ... -k regex:f16_table_step -s 3 -c 1   [round 2: f16_table_step_kernel<HEADING> (K1t, one aircraft per thread), n = 4e6]" > /dev/null
python tools/ncu_regions.py gpurun_out/prof_k1t.ncu-rep 512 >> gpurun_out/r02_table_step_kernel_ncu_full.txt 2>&1
TOP=40 python tools/ncu_lines.py gpurun_out/prof_k1t.ncu-rep $LIB f16_table_step_kernelILi0 4000000 >> gpurun_out/r02_table_step_kernel_ncu_full.txt 2>&1
rm -f gpurun_out/prof_k1t.ncu-rep
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:mma3_kernel -c 4 --csv --log-file gpurun_out/k_mlp_tc_ncu_mma3.csv ./tools/mlp_tc_bench 4194304 1 > /dev/null 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default_1gpu.json 2> gpurun_out/k_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/k_bench_ref.err
du -sh gpurun_out; ls -la gpurun_out | head -30
head -30 gpurun_out/r02_f16_step_kernel_ncu_full.txt
grep -E "mma3" gpurun_out/k_mlp_tc_ncu_mma3.csv | head -6 | cut -c1-300
