#!/bin/bash
# round 2, call A: GPU tests of the new boundary / plug-in surface + first bench with the mapped boundary
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a_env.txt 2>&1
nproc >> gpurun_out/a_env.txt; lscpu | grep "Model name" >> gpurun_out/a_env.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
for b in mapped pipelined copy; do
  timeout 300 python bench.py --steps 50 --warmup 5 --no-side --no-cpu --boundary $b > gpurun_out/a_bench_$b.json 2>> gpurun_out/a_bench.err
done
NPLANE_MAPPED_ACTIONS=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-side --no-cpu --boundary mapped > gpurun_out/a_bench_mapped_h2d.json 2>> gpurun_out/a_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/a_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], d.get('clocks'), {k:(v if not isinstance(v,dict) else '...') for k,v in (d.get('side') or {}).items()})
    except Exception as e: print(f, 'ERR', e)
PY
