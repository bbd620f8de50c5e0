#!/bin/bash
# round 2, call M: K1t with the guess-and-walk cell search (vs the scan) and CTA shapes, same box; table tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tables_env.py tests/test_gpu_tables.py -m gpu -q > gpurun_out/m_pytest.log 2>&1; tail -3 gpurun_out/m_pytest.log
for v in t0_scan t_default t1_512x1 t2_256x2 t3_320x2; do
NPLANE_LIB=$PWD/build/ab/$v.so timeout 300 python - > gpurun_out/m_tab_$v.json 2>> gpurun_out/m.err <<'PY'
import json, torch, bench
print(json.dumps(bench.side_tables(torch.device("cuda:0"), 6458.4, torch.cuda.synchronize)))
PY
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/m_tab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'ms %.4f'%d['ms_per_step'], 'a-s/s %.4g frac %.3f'%(d['aircraft_steps_per_s'], d['frac']), d['launch'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/m.err
