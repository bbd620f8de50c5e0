#!/bin/bash
mkdir -p gpurun_out
for N in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu.json").read().strip().splitlines()[-1]); print($N, "%.4g a-s/s"%d["value"], "%.4f ms"%d["ms_per_step"], "e2e %.3g"%d["e2e"]["value"], "lines:", len(open("gpurun_out/bench_${N}gpu.json").read().strip().splitlines()))
except Exception as e: print($N, "failed", e, open("gpurun_out/bench_${N}gpu.err").read()[-1500:])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/combat_exchange_check.py > gpurun_out/combat_exchange_8gpu.json 2> gpurun_out/combat_exchange_8gpu.err; echo "exchange rc=$?"; tail -1 gpurun_out/combat_exchange_8gpu.json
