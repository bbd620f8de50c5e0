#!/bin/bash
# round 2, call O: chunk pattern sweep of the pipelined numpy boundary (same box)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for pat in 1,2,3,4,4,4 1,1,2,3,4,4,4 1,2,4,6,6,6 1,2,4,8,8 2,3,4,4,4,4 1,2,3,4,5,5,5,5,5 1,1,2,2,4,4,4,4,4 1,3,6,8,8 1,2,3,4,4,4; do
  NPLANE_PIPELINE=$pat timeout 300 python bench.py --steps 30 --warmup 5 --no-side --no-cpu --e2e-steps 60 > gpurun_out/o_tmp.json 2>> gpurun_out/o.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/o_tmp.json').read().strip().splitlines()[-1]); print('pattern $pat', 'e2e ms %.4f'%d['e2e']['ms_per_step'], '%.4g'%d['e2e']['value'])
PY
done
tail -3 gpurun_out/o.err
