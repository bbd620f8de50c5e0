#!/bin/bash
# same-box A/B of two libraries (NPLANE_LIB): headline K1 throughput + e2e, K1c latency.  usage: tools/ab_headline.sh libA.so libB.so
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for rep in 1 2; do
  for l in "$@"; do
    echo "== $l"
    NPLANE_LIB=$PWD/$l python bench.py --steps 100 --warmup 20 --no-cpu --no-side | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g  ms %.4f  e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
    NPLANE_LIB=$PWD/$l python tools/k1c_ab.py
  done
done 2>&1 | tee gpurun_out/ab_headline.txt
