#!/bin/bash
# round 2, call I (8 GPUs): scaling bench N = 8 and 4 (weak + strong + side configs 4, 5 incl. the role-sharded combat step),
# PCIe copy-only ceiling at N = 4, 8, role-sharded combat check at 8 ranks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/i_bench_8gpu.json 2> gpurun_out/i_bench_8gpu.err; echo "bench8 rc=$?"
timeout 900 $TR --nproc-per-node 4 --master-port 29802 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/i_bench_4gpu.json 2> gpurun_out/i_bench_4gpu.err; echo "bench4 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29803 tools/pcie_ceiling.py > gpurun_out/i_pcie_8.json 2> gpurun_out/i_pcie.err
timeout 300 $TR --nproc-per-node 4 --master-port 29804 tools/pcie_ceiling.py > gpurun_out/i_pcie_4.json 2>> gpurun_out/i_pcie.err
timeout 600 $TR --nproc-per-node 8 --master-port 29805 tools/combat_role_check.py --envs 400000 --steps 20 > gpurun_out/i_role_check_8.json 2> gpurun_out/i_role_check.err; echo "role rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/i_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'value' in d:
            print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'strong', json.dumps(d.get('strong'))[:300])
            for k,v in (d.get('side') or {}).items(): print('    ', k, json.dumps(v)[:1200])
        else: print(f, json.dumps(d)[:900])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/i_bench_8gpu.err gpurun_out/i_role_check.err 2>/dev/null | tail -12
