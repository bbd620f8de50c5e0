#!/bin/bash
# round 2, run X (8 GPUs): the final build -- scaling bench at N = 8 (weak + strong + side configs 4, 5 incl. the role-sharded combat
# step) and the role-sharded combat check at 8 ranks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/x_bench_8gpu.json 2> gpurun_out/x_bench_8gpu.err; echo "bench8 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29805 tools/combat_role_check.py --envs 400000 --steps 10 > gpurun_out/x_role_check_8.json 2> gpurun_out/x_role_check.err; echo "role rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/x_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'value' in d:
            print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'strong', json.dumps(d.get('strong'))[:300])
            for k,v in (d.get('side') or {}).items(): print('    ', k, json.dumps(v)[:900])
        else: print(f, json.dumps(d)[:900])
    except Exception as e: print(f, 'ERR', e)
PY
