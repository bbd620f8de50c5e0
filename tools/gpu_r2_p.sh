#!/bin/bash
# round 2, call P: final verification on one GPU -- pytest -m gpu (as the driver runs it), smoke(), bench both arms
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p_pytest.log; tail -3 gpurun_out/p_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/p_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/p_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default_1gpu.json 2> gpurun_out/p_bench.err; echo "bench rc=$?"
timeout 900 python bench.py > gpurun_out/r02_bench_noflags_1gpu.json 2>> gpurun_out/p_bench.err; echo "bench(default flags) rc=$?"
python - <<'PY'
import json
for f in ('r02_bench_reference_arm','r02_bench_default_1gpu','r02_bench_noflags_1gpu'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d.get('clocks'), d.get('cpu_baseline'), d.get('reference_cuda_eager'))
        for k,v in (d.get('side') or {}).items(): print('   ', k, json.dumps(v)[:420])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/p_bench.err
