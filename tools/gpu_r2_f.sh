#!/bin/bash
# round 2, call F: same-box A/B of K1 builds (why did the free-running step time go from 0.47 to 0.55 ms?)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for rep in 1 2; do
for v in vb v0 v1 v2 v3; do
  NPLANE_LIB=$PWD/build/ab/$v.so timeout 300 python bench.py --steps 100 --warmup 10 --no-side --no-cpu > gpurun_out/f_$v.$rep.json 2>> gpurun_out/f.err
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/f_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'ms %.4f'%d['ms_per_step'], 'value %.4g'%d['value'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/f.err
