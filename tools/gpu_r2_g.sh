#!/bin/bash
# round 2, call G: same-box A/B, second batch (K1: staging / shared Philox / shared powf / unrolled passes; K1t: CTA shapes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for rep in 1 2; do
for v in w0 w1_nostage w2_philox w3_pow w4_unrollpass; do
  NPLANE_LIB=$PWD/build/ab/$v.so timeout 300 python bench.py --steps 100 --warmup 10 --no-side --no-cpu > gpurun_out/g_$v.$rep.json 2>> gpurun_out/g.err
done
done
for v in w0 t1_256x3_nostage t2_256x4_nostage t3_512x1 t4_384x2_nostage; do
NPLANE_LIB=$PWD/build/ab/$v.so timeout 300 python - > gpurun_out/g_tab_$v.json 2>> gpurun_out/g.err <<'PY'
import json, torch, bench
print(json.dumps(bench.side_tables(torch.device("cuda:0"), 6458.4, torch.cuda.synchronize)))
PY
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/g_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'ms %.4f'%d['ms_per_step'], ('value %.4g'%d['value']) if 'value' in d else ('a-s/s %.4g frac %.3f'%(d['aircraft_steps_per_s'], d['frac'])))
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/g.err
