#!/bin/bash
# round 2, call D: full GPU suite, K1 with the cross-step record (cache on / off), ncu full capture + launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -8 gpurun_out/d_pytest.log
timeout 600 python bench.py --steps 300 --warmup 10 --no-side --no-cpu > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
timeout 600 python bench.py --steps 300 --warmup 10 --no-side --no-cpu --no-cache > gpurun_out/d_bench_nocache.json 2>> gpurun_out/d_bench.err
NPLANE_OBS_STORE=stg timeout 600 python bench.py --steps 300 --warmup 10 --no-side --no-cpu > gpurun_out/d_bench_stg.json 2>> gpurun_out/d_bench.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:f16_step -s 5 -c 1 -f -o gpurun_out/prof_k1_r02d python bench.py --steps 5 --warmup 3 --no-cpu --no-side > gpurun_out/d_ncu_k1.log 2>&1; tail -1 gpurun_out/d_ncu_k1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/d_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-side > gpurun_out/d_ncu_launch.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/d_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['e2e'].get('boundary'))
    except Exception as e: print(f, 'ERR', e)
PY
ls -la gpurun_out/*.ncu-rep
