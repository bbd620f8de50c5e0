#!/bin/bash
# round 2, call N: do the three warps of a scheduler run their MLP and tail phases in lockstep?  phase-shift them at kernel start
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for ns in 0 5000 10000 15000 20000 30000 0; do
  NPLANE_STAGGER_NS=$ns timeout 300 python bench.py --steps 100 --warmup 10 --no-side --no-cpu > gpurun_out/n_stagger_$ns.json 2>> gpurun_out/n.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/n_stagger_$ns.json').read().strip().splitlines()[-1]); print('stagger_ns', $ns, 'ms %.4f'%d['ms_per_step'], 'value %.4g'%d['value'])
PY
done
tail -3 gpurun_out/n.err
