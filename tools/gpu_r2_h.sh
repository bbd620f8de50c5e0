#!/bin/bash
# round 2, call H: tensor-core (mma.sync 3xTF32) vs FFMA2 micro-benchmark for the 20 -> 10 layer, with ncu instruction counts
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
./tools/mlp_tc_bench 4194304 20 > gpurun_out/h_mlp_tc_bench.txt 2>&1; cat gpurun_out/h_mlp_tc_bench.txt
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:"ffma2_kernel|mma3_kernel" -c 12 --csv --log-file gpurun_out/h_mlp_tc_ncu.csv ./tools/mlp_tc_bench 4194304 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/h_mlp_tc_ncu.csv')) if len(r)>12 and r[0].isdigit()]
for r in rows: print(r[4][:14], r[8], r[-3], r[-1])
PY
