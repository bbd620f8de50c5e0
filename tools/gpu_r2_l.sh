#!/bin/bash
# round 2, call L: pipelined numpy boundary v2 (zero-copy actions, reward / flags written by the kernel) vs v1, same box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "boundary or vec_env or ring" > gpurun_out/l_pytest.log 2>&1; tail -3 gpurun_out/l_pytest.log
for rep in 1 2; do
NPLANE_HOST_PIPE_V1=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-side --no-cpu --e2e-steps 40 > gpurun_out/l_v1.$rep.json 2>> gpurun_out/l.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-side --no-cpu --e2e-steps 40 > gpurun_out/l_v2.$rep.json 2>> gpurun_out/l.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/l_v*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'e2e %.4g'%d['e2e']['value'], 'ms %.4f'%d['e2e']['ms_per_step'], d['e2e']['boundary'])
PY
tail -3 gpurun_out/l.err
