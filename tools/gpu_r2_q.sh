#!/bin/bash
# round 2, call Q (8 GPUs): final-build scaling lines N = 8, 4, 2 (weak + strong + sides)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4 2; do
  timeout 900 $TR --nproc-per-node $N --master-port 2990$N bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/q_bench_${N}gpu.err; echo "bench$N rc=$?"
done
python - <<'PY'
import json
for N in (8,4,2):
    try:
        d=json.loads(open(f'gpurun_out/r02_bench_{N}gpu.json').read().strip().splitlines()[-1])
        print(N, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'strong', json.dumps(d.get('strong'))[:260])
        for k,v in (d.get('side') or {}).items(): print('    ', k, json.dumps(v)[:1100])
    except Exception as e: print(N, 'ERR', e)
PY
tail -2 gpurun_out/q_bench_8gpu.err
