#!/bin/bash
# round 2, run V (2 GPUs): the final build -- full GPU suite incl. the 2-GPU tests, role-sharded combat check (also at a small
# population: 128-thread local half), 2-GPU bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.log
tail -4 gpurun_out/v_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29702 tools/combat_role_check.py --envs 200000 --steps 20 > gpurun_out/v_role_check.json 2> gpurun_out/v_role_check.err; echo "role rc=$?"
timeout 600 $TR --nproc-per-node 2 --master-port 29704 tools/combat_role_check.py --envs 6000 --steps 20 > gpurun_out/v_role_check_small.json 2> gpurun_out/v_role_check_small.err; echo "role small rc=$?"
timeout 900 $TR --nproc-per-node 2 --master-port 29703 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/v_bench_2gpu.json 2> gpurun_out/v_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/v_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'value' in d: print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'strong', (d.get('strong') or {}).get('efficiency_vs_n1'), json.dumps((d.get('side') or {}).get('combat'))[:700])
        else: print(f, json.dumps(d)[:600])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/v_role_check.err gpurun_out/v_role_check_small.err gpurun_out/v_bench_2gpu.err
