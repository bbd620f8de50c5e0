#!/bin/bash
# block-size sweep, MLP weight-path micro-benchmark, ncu full capture of the step kernel
mkdir -p gpurun_out
./tools/mlp_bench > gpurun_out/mlp_bench.txt 2>&1
for b in 128 256 384 512; do
  NPLANE_LIB=$PWD/neuralplane_b200/_lib/libnplane_all.so NPLANE_BLOCK=$b timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --e2e-steps 3 > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:f16_step -s 5 -c 2 -o gpurun_out/prof_step_r1a python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/mlp_bench.txt
for b in 128 256 384 512; do python - <<PY
import json
d=json.load(open("gpurun_out/bench_b$b.json")); print($b, d["value"], d["ms_per_step"], d["config"]["launch"])
PY
done
