#!/bin/bash
for n in 1000000 2000000; do
for b in 320 352 384 448 512; do
  NPLANE_LIB=$PWD/neuralplane_b200/_lib/libnplane_all.so NPLANE_BLOCK=$b timeout 300 python bench.py --n $n --steps 200 --warmup 20 --no-cpu --e2e-steps 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_b.json")); print($n, $b, "%.4g"%d["value"], "%.4f ms"%d["ms_per_step"], "slabs/CTA %.2f"%($n/2/$b/148))
except Exception as e: print($b, "failed", e, open("gpurun_out/bench_b.err").read()[-500:])
PY
done
done
