#!/bin/bash
# round 2, call B (2 GPUs): full GPU test suite (incl. the 2-GPU tests), role-sharded combat check, PCIe ceiling at N = 1, 2,
# e2e boundary A/B with the TMA bulk store of the observation tile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python tools/pcie_ceiling.py > gpurun_out/b_pcie_1.json 2> gpurun_out/b_pcie.err
timeout 300 $TR --nproc-per-node 2 --master-port 29701 tools/pcie_ceiling.py > gpurun_out/b_pcie_2.json 2>> gpurun_out/b_pcie.err
for b in mapped pipelined; do
  timeout 300 python bench.py --steps 50 --warmup 5 --no-side --no-cpu --boundary $b > gpurun_out/b_bench_$b.json 2>> gpurun_out/b_bench.err
done
NPLANE_OBS_STORE=stg timeout 300 python bench.py --steps 200 --warmup 5 --no-side --no-cpu --boundary mapped > gpurun_out/b_bench_mapped_stg.json 2>> gpurun_out/b_bench.err
timeout 600 $TR --nproc-per-node 2 --master-port 29702 tools/combat_role_check.py --envs 200000 --steps 20 > gpurun_out/b_role_check.json 2> gpurun_out/b_role_check.err; echo "role rc=$?"
timeout 900 $TR --nproc-per-node 2 --master-port 29703 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/b_bench_2gpu.json 2> gpurun_out/b_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'value' in d: print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'strong', (d.get('strong') or {}).get('efficiency_vs_n1'), json.dumps((d.get('side') or {}).get('combat'))[:900])
        else: print(f, json.dumps(d)[:700])
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/b_role_check.err gpurun_out/b_bench_2gpu.err gpurun_out/b_bench.err
