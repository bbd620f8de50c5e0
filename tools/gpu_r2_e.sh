#!/bin/bash
# round 2, call E: GPU suite with K1t / 2-v-2 / 128-thread latency variant, step time vs step index, default bench with sides
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -8 gpurun_out/e_pytest.log
timeout 600 python tools/step_time_profile.py > gpurun_out/e_step_profile.json 2> gpurun_out/e_step_profile.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench_20.json 2> gpurun_out/e_bench.err
timeout 900 python bench.py --steps 300 --warmup 10 --no-side --no-cpu > gpurun_out/e_bench_300.json 2>> gpurun_out/e_bench.err
NPLANE_TAB_KERNEL=pairs timeout 600 python - > gpurun_out/e_tab_pairs.json 2>> gpurun_out/e_bench.err <<'PY'
import json, torch, bench
dev = torch.device("cuda:0")
print(json.dumps(bench.side_tables(dev, 6458.4, torch.cuda.synchronize)))
PY
python - <<'PY'
import json
for f in ('e_bench_20','e_bench_300'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'])
        for k,v in (d.get('side') or {}).items(): print('   ', k, json.dumps(v)[:400])
    except Exception as e: print(f,'ERR',e)
print(open('gpurun_out/e_step_profile.json').read()[:3000])
print(open('gpurun_out/e_tab_pairs.json').read()[:600])
PY
