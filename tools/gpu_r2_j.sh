#!/bin/bash
# round 2, call J: final build -- GPU suite, driver-style bench (both arms), ncu full captures (K1, K1t, UAV slab) + launch list,
# compute-sanitizer over the round-2 kernels, planning / combat shard A/B (384 vs 512-thread CTAs at 125 k aircraft)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
tail -4 gpurun_out/j_pytest.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/j_bench_ref.json 2> gpurun_out/j_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 500 --warmup 20 --no-side --no-cpu > gpurun_out/j_bench_500.json 2>> gpurun_out/j_bench.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:f16_step_kernel -s 5 -c 1 -f -o gpurun_out/prof_k1_r02j python bench.py --steps 5 --warmup 3 --no-cpu --no-side > gpurun_out/j_ncu_k1.log 2>&1; tail -1 gpurun_out/j_ncu_k1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/j_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-side > gpurun_out/j_ncu_launch.log 2>&1
cat > gpurun_out/prof_side.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from neuralplane_b200 import ControlEnv
which = sys.argv[1]
n = 8_000_000 if which == "uav" else 4_000_000
env = ControlEnv(num_envs=n, config="control" if which == "uav" else "heading", model="UAV" if which == "uav" else "F16_tables", random_seed=0, device="cuda:0")
env.reset()
a = torch.rand((n, 4), device="cuda") * 2 - 1
for k in range(6): env.step(a)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:uav_step_slab -s 3 -c 1 -f -o gpurun_out/prof_uav_r02j python gpurun_out/prof_side.py uav > gpurun_out/j_ncu_uav.log 2>&1; tail -1 gpurun_out/j_ncu_uav.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:f16_table_step -s 3 -c 1 -f -o gpurun_out/prof_k1t_r02j python gpurun_out/prof_side.py tab > gpurun_out/j_ncu_k1t.log 2>&1; tail -1 gpurun_out/j_ncu_k1t.log
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:mma3_kernel -c 4 --csv --log-file gpurun_out/j_mlp_tc_ncu_mma3.csv ./tools/mlp_tc_bench 4194304 1 > /dev/null 2>&1
python - > gpurun_out/j_shard_ab.json 2>> gpurun_out/j_bench.err <<'PY'
import json, os, torch
from neuralplane_b200 import PlanningEnv, SingleCombatEnv
import bench
dev = torch.device("cuda:0")
out = {}
for name, mk, aw in (("planning_125k", lambda: PlanningEnv(num_envs=125_000, config="tracking", random_seed=0, device=dev), 3),
                     ("combat_62500_pairs", lambda: SingleCombatEnv(num_envs=62_500, config="selfplay", random_seed=0, device=dev), 4)):
    for blk in ("auto", "384"):
        if blk != "auto": os.environ["NPLANE_BLOCK"] = blk
        env = mk()
        os.environ.pop("NPLANE_BLOCK", None)
        env.reset()
        acts = [torch.rand((env.n, aw), device=dev) * 2 - 1 for _ in range(2)]
        ms = bench.timed_steps(lambda k: env.step(acts[k % 2]), 10, 3, torch.cuda.synchronize) / 10
        out[f"{name}_{blk}"] = {"ms_per_env_step": ms, "launch": env.launch_info()}
print(json.dumps(out))
PY
cat gpurun_out/j_shard_ab.json
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_r2.py > gpurun_out/j_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/j_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_r2.py > gpurun_out/j_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/j_racecheck.log
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_r2.py > gpurun_out/j_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/j_synccheck.log
python - <<'PY'
import json
for f in ('j_bench_ref','j_bench','j_bench_500'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d.get('clocks'), d.get('cpu_baseline'), d.get('reference_cuda_eager'))
        for k,v in (d.get('side') or {}).items(): print('   ', k, json.dumps(v)[:500])
    except Exception as e: print(f,'ERR',e)
PY
ls -la gpurun_out/*.ncu-rep
