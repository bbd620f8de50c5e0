#!/usr/bin/env python
"""Benchmark of the hot path: aircraft-steps/sec of ControlEnv.step (F16, heading) at 10^6 aircraft per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n AIRCRAFT_PER_GPU] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W      (N > 1)

One "step" = one env.step() over the whole population = ONE launch of f16_step_kernel.  Prints ONE JSON line
(rank 0).  See DESIGN.md "Measurement" for the definitions of value / e2e / roofline / cpu_baseline and of the
extra keys `strong` (fixed 10^6 population sharded over the ranks) and `side` (BASELINE configs 3-5).
"""
import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_STEP = 276          # 100 B read + 176 B written per aircraft-step
ALGO_FLOP_PER_STEP = 52.6e3        # two full nlplant evaluations (reference arithmetic)
METRIC = "aircraft-steps/sec at N=10^6 (F16 Heading)"
UNIT = "aircraft-steps/s"
WORKLOAD = "F16 Heading task, ControlEnv, num_agents=10^6 per GPU, random-policy rollout (BASELINE configs[1])"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")   # the unmodified reference's envs/ (baseline/install_ref.py), when shipped


def workload_config(n):
    """The `config` object of BOTH arms (ours and --impl reference): what is computed, nothing about how."""
    return {"workload": WORKLOAD, "aircraft_per_gpu": n, "noise_scale": 0.01,
            "actions": "uniform(-1, 1)^4 random policy, pre-drawn; in-step episodic resets",
            "l2": "per-step working set (state+obs+cache >= 300 MB at n=10^6) exceeds the 126 MB L2",
            "sharding": "contiguous global index ranges per rank, no collective"}


def measured_traffic(kernel="f16_step_kernel"):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum at n = 10^6), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))[kernel]
        return float(d["dram_bytes_per_launch"]), d
    except Exception:
        return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons of one GPU sampled IN PROCESS through NVML every millisecond by a thread, so that even a
    10 ms timed region (--steps 20) is covered; `nvidia-smi -lms 25` as a subprocess (round 1) never emitted inside it."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, torch_device, period_s=0.001):
        self.samples, self.reason_bits, self.max_mhz, self.h = [], 0, None, None
        self._stop = threading.Event()
        self.period = period_s
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(torch_device).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_device.index or 0)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _read(self):
        nv = self.nv
        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.reason_bits |= int(get(self.h))
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            try:
                self._read()
            except Exception:
                break
            time.sleep(self.period)

    def start(self):
        if self.h is not None:
            self.t.start()
        return self

    def stop(self):
        if self.h is not None:
            try:
                self._read()               # one sample taken with the timed work still on the device's recent history
            except Exception:
                pass
            self._stop.set()
            self.t.join(timeout=2)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": [name for bit, name in self.REASONS if self.reason_bits & bit], "samples": len(self.samples),
               "how": "in-process NVML thread, 1 ms period, over the timed region"}
        return out


# ---------------------------------------------------------------------------------------------------------------------------
# CPU side: the reference itself (baseline/_ref, when shipped) or the oracle port
# ---------------------------------------------------------------------------------------------------------------------------
def reference_available():
    """NPLANE_BENCH_PORT=1 forces the oracle-port fallback (what a box without baseline/_ref runs)."""
    return os.path.exists(os.path.join(REF_DIR, "envs", "control_env.py")) and not os.environ.get("NPLANE_BENCH_PORT")


def make_reference_env(n, device):
    """The UNMODIFIED reference ControlEnv (baseline/_ref/envs, with the gym / torchdiffeq stubs of SURVEY App. F)."""
    for p in (os.path.join(REF_DIR, "envs"), REF_DIR, os.path.join(REF_DIR, "_shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import envs.control_env as control_env
    with contextlib.redirect_stdout(io.StringIO()):
        return control_env.ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device=device)


def time_reference_env(n, device, warmup, steps, threads=None):
    """measure_env.py:65-78 on `device`: env.step(action) in a loop, stdout silenced (the termination conditions print)."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    env = make_reference_env(n, device)
    g = torch.Generator().manual_seed(1)
    acts = [(torch.rand((n, 4), generator=g) * 2 - 1).to(device) for _ in range(3)]
    sync = (lambda: torch.cuda.synchronize()) if str(device).startswith("cuda") else (lambda: None)
    with contextlib.redirect_stdout(io.StringIO()):
        env.reset()
        for k in range(warmup):
            env.step(acts[k % 3])
        sync()
        t0 = time.perf_counter()
        for k in range(steps):
            env.step(acts[k % 3])
        sync()
        el = time.perf_counter() - t0
    return n * steps / el, el


def time_port(n, warmup, steps, threads=None):
    """The oracle (torch-CPU restatement of the reference step, bit-identical to it: tests/test_oracle_golden.py)."""
    import torch
    from oracle import tapes
    from oracle.f16_oracle import F16EnvOracle
    if threads:
        torch.set_num_threads(threads)
    env = F16EnvOracle(n, "heading")
    env.reset(torch.from_numpy(tapes.reset_draw_tape(1, 0, n)))
    acts = [torch.from_numpy(tapes.action_tape(1, k, n, 1.0)) for k in range(1, 4)]
    draws = torch.from_numpy(tapes.reset_draw_tape(1, 1, n))
    noise = torch.randn(n, 22)
    for k in range(warmup):
        env.step(acts[k % 3], draws, noise)
    t0 = time.perf_counter()
    for k in range(steps):
        env.step(acts[k % 3], draws, noise)
    el = time.perf_counter() - t0
    return n * steps / el, el


def cpu_arm(n, warmup, steps):
    """(value, elapsed, kind, cores, sample text) of the CPU implementation on every host thread."""
    thr = os.cpu_count()
    if reference_available():
        v, el = time_reference_env(n, "cpu", warmup, steps, thr)
        kind, what = "reference", "the unmodified reference ControlEnv(device='cpu') from baseline/_ref"
    else:
        v, el = time_port(n, warmup, steps, thr)
        kind, what = "port", "oracle port of the reference step (torch CPU ops; baseline/_ref not shipped)"
    return v, el, kind, thr, f"{steps} steps x {n} aircraft after {warmup} warm-up steps in {el:.1f} s: {what}, {thr} threads"


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, same config / steps /
    warm-up as our arm.  Rank 0 only.  Also reports the unmodified reference with device='cuda:0' (eager PyTorch on the
    same B200) under `reference_cuda_eager` -- labelled, not the headline."""
    if rank != 0:
        return
    n, W, K = args.n, max(args.warmup, 1), args.steps
    note = None
    if not reference_available():
        n, note = min(n, args.cpu_n), "baseline/_ref not shipped: oracle port on a reduced population"
    v, el, kind, thr, sample = cpu_arm(n, W, K)
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * el / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.n),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": thr, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if note:
        out["note"] = note
        out["config"] = dict(out["config"], sample_aircraft=n)
    if reference_available() and not args.no_side:
        try:
            import torch
            if torch.cuda.is_available():
                ve, ee = time_reference_env(args.n, "cuda:0", 2, 5)
                out["reference_cuda_eager"] = {"value": ve, "unit": UNIT, "ms_per_step": 1e3 * ee / 5, "steps": 5, "warmup": 2,
                                               "what": "the unmodified reference ControlEnv(device='cuda:0'): eager PyTorch on this B200"}
        except Exception as e:
            out["reference_cuda_eager"] = {"error": repr(e)[:300]}
        # the population the reference itself trains at (scripts/train_heading.sh:13): our arm reports it under side.small_n_latency
        try:
            import torch
            small = {"workload": "the unmodified reference ControlEnv, num_agents=3000 (scripts/train_heading.sh)", "steps": 20, "warmup": 3}
            vs, es = time_reference_env(3000, "cpu", 3, 20, threads=os.cpu_count())
            small["cpu"] = {"us_per_step": 1e6 * es / 20, "aircraft_steps_per_s": vs, "cores": os.cpu_count()}
            if torch.cuda.is_available():
                vs, es = time_reference_env(3000, "cuda:0", 3, 20)
                small["cuda_eager"] = {"us_per_step": 1e6 * es / 20, "aircraft_steps_per_s": vs}
            out["reference_small_n"] = small
        except Exception as e:
            out["reference_small_n"] = {"error": repr(e)[:300]}
    args.emit(out)


# ---------------------------------------------------------------------------------------------------------------------------
# side measurements (outside the headline's timed region)
# ---------------------------------------------------------------------------------------------------------------------------
def timed_steps(step, K, W, barrier):
    import torch
    for k in range(W):
        step(k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        step(k)
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def side_uav_roofline(dev, hbm_peak, barrier, n=8_000_000, K=100, W=20):
    """The path's HBM-bound kernel (UAV aircraft plug-in, BASELINE configs[2]'s second-model slot) against the same measured
    HBM peak, timed live with CUDA events."""
    import torch
    from neuralplane_b200 import ControlEnv
    env = ControlEnv(num_envs=n, config="control", model="UAV", random_seed=0, device=dev)
    env.reset()
    acts = [torch.rand((n, 4), device=dev) * 2 - 1 for _ in range(2)]
    ms = timed_steps(lambda k: env.step(acts[k % 2]), K, W, barrier) / K
    achieved = 268.0 * n / (ms * 1e-3) / 1e9
    return {"kernel": "uav_step_slab_kernel", "workload": f"UAV Control task, ControlEnv, num_agents={n}, random policy, noise_scale "
            f"{float(env.task.noise_scale)}; working set 2.1 GB per step >> L2", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
            "unit": "GB/s", "frac": achieved / hbm_peak, "algorithmic_bytes_per_aircraft_step": 268, "ms_per_step": ms,
            "aircraft_steps_per_s": n / (ms * 1e-3), "steps": K, "warmup": W, "launch": env.launch_info()}


def side_uav_config3(dev, barrier, n=100_000, K=200, W=20):
    """BASELINE configs[2] as written: Control task, num_agents = 10^5, second aircraft model (the reference's is UAV)."""
    import torch
    from neuralplane_b200 import ControlEnv
    env = ControlEnv(num_envs=n, config="control", model="UAV", random_seed=0, device=dev)
    env.reset()
    acts = [torch.rand((n, 4), device=dev) * 2 - 1 for _ in range(2)]
    ms = timed_steps(lambda k: env.step(acts[k % 2]), K, W, barrier) / K
    return {"workload": "UAV Control task, ControlEnv, num_agents=10^5 (BASELINE configs[2])", "aircraft_steps_per_s": n / (ms * 1e-3),
            "ms_per_step": ms, "steps": K, "note": "26.8 MB per step: L2-resident at this size, so not a roofline line"}


def side_planning(dev, rank, world, barrier, max_over_ranks, n_total=1_000_000, K=3, W=1):
    """BASELINE configs[3]: PlanningEnv (tracking task) with the fused PID low-level controller, 10^6 aircraft in total
    sharded over the ranks, 50 FDM sub-steps per env step."""
    import torch
    from neuralplane_b200 import PlanningEnv
    from neuralplane_b200.sharding import shard_range
    base, n = shard_range(n_total, rank, world)
    env = PlanningEnv(num_envs=n, config="tracking", random_seed=0, device=dev, index_base=base)
    env.reset()
    acts = [torch.rand((n, 3), device=dev) * 2 - 1 for _ in range(2)]
    ms = max_over_ranks(timed_steps(lambda k: env.step(acts[k % 2]), K, W, barrier)) / K
    return {"workload": "F16 Tracking task, PlanningEnv, fused PID low-level controller, num_agents=10^6 in total sharded by rank "
                        "(BASELINE configs[3])", "aircraft_total": n_total, "sub_steps": env.n_substeps,
            "env_steps_per_s": n_total / (ms * 1e-3), "fdm_steps_per_s": n_total * env.n_substeps / (ms * 1e-3),
            "ms_per_env_step": ms, "steps": K, "scaling": "strong"}


def side_combat(dev, rank, world, barrier, max_over_ranks, pairs_total=500_000, K=5, W=2):
    """BASELINE configs[4]: SingleCombat 1-v-1, 5 x 10^5 pairs in total.  Pair-sharded (both aircraft of a pair on one rank,
    no exchange) and, at N >= 2, role-sharded (egos and opponents on different ranks; the partner records cross NVLink inside
    the step, exchange inside the timed region)."""
    import torch
    from neuralplane_b200 import SingleCombatEnv
    from neuralplane_b200.sharding import shard_range
    base, n = shard_range(2 * pairs_total, rank, world)
    env = SingleCombatEnv(num_envs=n // 2, config="selfplay", random_seed=0, device=dev, index_base=base)
    env.reset()
    acts = [torch.rand((n, 4), device=dev) * 2 - 1 for _ in range(2)]
    ms = max_over_ranks(timed_steps(lambda k: env.step(acts[k % 2]), K, W, barrier)) / K
    out = {"workload": "SingleCombat 1v1 self-play, 5x10^5 pairs in total (BASELINE configs[4])", "pairs_total": pairs_total,
           "pair_sharded": {"env_steps_per_s": 2 * pairs_total / (ms * 1e-3), "fdm_steps_per_s": 2 * pairs_total * env.n_substeps / (ms * 1e-3),
                            "ms_per_env_step": ms, "exchange": "none (the pair lives in one thread)"}}
    del env
    if world >= 2 and world % 2 == 0:
        try:
            from neuralplane_b200.envs.singlecombat_env import role_sharded_bench
            out["combat_role_sharded"] = role_sharded_bench(dev, rank, world, pairs_total, K, W, barrier, max_over_ranks)
        except Exception as e:
            out["combat_role_sharded"] = {"error": repr(e)[:300]}
    return out


def side_tables(dev, hbm_peak, barrier, n=4_000_000, K=100, W=20):
    """The table aero back-end (ControlEnv(model='F16_tables'), SURVEY f-3) through K1t, one aircraft per thread."""
    import torch
    from neuralplane_b200 import ControlEnv
    env = ControlEnv(num_envs=n, config="heading", model="F16_tables", random_seed=0, device=dev)
    env.reset()
    acts = [torch.rand((n, 4), device=dev) * 2 - 1 for _ in range(2)]
    ms = timed_steps(lambda k: env.step(acts[k % 2]), K, W, barrier) / K
    achieved = 276.0 * n / (ms * 1e-3) / 1e9
    return {"kernel": "f16_table_step_kernel", "workload": f"F16 Heading task with the table aero back-end, num_agents={n}, random policy",
            "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "algorithmic_bytes_per_aircraft_step": 276, "ms_per_step": ms, "aircraft_steps_per_s": n / (ms * 1e-3), "steps": K,
            "launch": env.launch_info()}


def side_small_n(dev, barrier, n=3000, K=200, W=20):
    """The population the reference trains at (scripts/train_heading.sh:13: 3 000 envs): per-step latency of the device-
    resident step under CUDA-graph replay (no launch overhead from Python)."""
    import torch
    from neuralplane_b200 import ControlEnv
    env = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device=dev)
    env.reset()
    a = torch.rand((n, 4), device=dev) * 2 - 1
    for _ in range(3):
        env.step(a)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            env.step(a)
    ms = timed_steps(lambda k: g.replay(), K // 10, 2, barrier) / (K // 10 * 10)
    # the same population through the numpy boundary the reference's runners call (host actions in, host obs out: one mapped launch)
    import numpy as np
    from neuralplane_b200 import GPUVecEnv
    venv = GPUVecEnv([lambda: ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device=dev)])
    venv.reset()
    ha = (np.random.default_rng(0).random((n, 1, 4), dtype=np.float32) * 2 - 1)
    for _ in range(20):
        venv.step(ha)
    t0 = time.perf_counter()
    for _ in range(K):
        venv.step(ha)
    e2e_us = 1e6 * (time.perf_counter() - t0) / K
    return {"workload": f"F16 Heading task, ControlEnv, num_agents={n} (the reference's training population)",
            "us_per_step": ms * 1e3, "aircraft_steps_per_s": n / (ms * 1e-3), "how": "device-resident step under CUDA-graph replay",
            "e2e_us_per_step": e2e_us, "e2e_aircraft_steps_per_s": n / (e2e_us * 1e-6), "e2e_boundary": venv.boundary,
            "launch": env.launch_info(), "planning_10000_envs": side_plan_small(dev, barrier)}


def side_plan_small(dev, barrier, n=10_000, K=20, W=5):
    """scripts/train_tracking.sh: PlanningEnv at 10 000 envs, 50 FDM sub-steps per env step under the fused PID controller."""
    import torch
    from neuralplane_b200 import PlanningEnv
    env = PlanningEnv(num_envs=n, config="tracking", random_seed=0, device=dev)
    env.reset()
    acts = [torch.rand((n, 3), device=dev) * 2 - 1 for _ in range(2)]
    ms = timed_steps(lambda k: env.step(acts[k % 2]), K, W, barrier) / K
    return {"workload": f"F16 Tracking task, PlanningEnv, fused PID low-level controller, num_agents={n} (the reference's training population)",
            "sub_steps": env.n_substeps, "ms_per_env_step": ms, "us_per_fdm_substep": ms * 1e3 / env.n_substeps,
            "fdm_steps_per_s": n * env.n_substeps / (ms * 1e-3), "launch": env.launch_info()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--n", type=int, default=1_000_000, help="aircraft per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-n", type=int, default=100_000, help="population of the oracle-port fallback (no baseline/_ref)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cache", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the side measurements (configs 3-5, UAV roofline, small n)")
    ap.add_argument("--boundary", default=None, help="host boundary of the e2e leg: mapped | pipelined | copy")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner) go to stderr instead
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)

    args.emit = emit
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from neuralplane_b200 import ControlEnv, GPUVecEnv
    from neuralplane_b200.sharding import max_over_ranks, reduce_counters, shard_range

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W, K, n = max(args.warmup, 3), args.steps, args.n

    # rank r owns the contiguous global index range [r*n, (r+1)*n): no data-path collective (aircraft are independent)
    venv = GPUVecEnv([lambda: ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device=dev,
                                         index_base=rank * n, use_coef_cache=not args.no_cache)], boundary=args.boundary)
    env = venv.gpu_vec_env
    env.reset()
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    actions = [torch.rand((n, 4), device=dev, generator=g) * 2 - 1 for _ in range(8)]   # random policy, pre-drawn

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(W):
        env.step(actions[k % 8])
    barrier()
    sampler = ClockSampler(dev).start() if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        env.step(actions[k % 8])
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    ms_max = max_over_ranks(ms)
    info = env.launch_info()
    venv_boundary = venv.boundary

    # end to end through the numpy boundary the runners call (GPUVecEnv.step): host numpy in, host numpy out
    Ke = max(3, min(args.e2e_steps, K))
    host_actions = [a.cpu().numpy().reshape(n, 1, 4) for a in actions[:2]]
    for k in range(3):
        venv.step(host_actions[k % 2])
    barrier()
    t0 = time.perf_counter()
    for k in range(Ke):
        venv.step(host_actions[k % 2])
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n * Ke / e2e_s
    counters = reduce_counters(env.termination_counters())

    # strong scaling: the SAME 10^6-aircraft population sharded over the ranks (BASELINE: "at N=10^6 ... @1/2/4/8 B200")
    strong = None
    n_total = n
    if world > 1:
        base, n_loc = shard_range(n_total, rank, world)
        senv = ControlEnv(num_envs=n_loc, config="heading", model="F16", random_seed=0, device=dev, index_base=base)
        senv.reset()
        sacts = [a[:n_loc].contiguous() for a in actions[:4]]
        sms = max_over_ranks(timed_steps(lambda k: senv.step(sacts[k % 4]), K, W, barrier))
        n1_value = n * K / (ms * 1e-3)              # this rank stepping the whole 10^6 alone (the weak leg above)
        strong = {"aircraft_total": n_total, "aircraft_per_gpu": n_loc, "value": n_total * K / (sms * 1e-3), "unit": UNIT,
                  "ms_per_step": sms / K, "efficiency_vs_n1": (n_total * K / (sms * 1e-3)) / (world * n1_value),
                  "n1_value": n1_value, "launch": senv.launch_info(),
                  "how": "fixed 10^6 population, shard_range per rank, device-timed max over ranks; n1_value = rank 0 stepping all "
                         "10^6 aircraft alone in the same run"}
        del senv, sacts
    else:
        strong = {"aircraft_total": n_total, "aircraft_per_gpu": n, "value": n * K / (ms * 1e-3), "unit": UNIT,
                  "ms_per_step": ms / K, "efficiency_vs_n1": 1.0, "how": "N = 1: the headline itself"}

    side = {}
    if not args.no_side:
        hbm_peak, _ = peaks()
        del venv, env, actions
        torch.cuda.empty_cache()
        jobs = [("planning", lambda: side_planning(dev, rank, world, barrier, max_over_ranks)),
                ("combat", lambda: side_combat(dev, rank, world, barrier, max_over_ranks))]
        if world == 1:
            jobs += [("uav_roofline", lambda: side_uav_roofline(dev, hbm_peak, barrier)),
                     ("uav_config3", lambda: side_uav_config3(dev, barrier)),
                     ("table_backend", lambda: side_tables(dev, hbm_peak, barrier)),
                     ("small_n_latency", lambda: side_small_n(dev, barrier))]
        for name, fn in jobs:            # every rank runs the same list (the sharded ones hold collectives)
            try:
                side[name] = fn()
            except Exception as e:       # a side line must never take the headline down
                side[name] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()

    if rank == 0:
        hbm_peak, peak_src = peaks()
        value = world * n * K / (ms_max * 1e-3)
        per_launch_s = ms * 1e-3 / K
        achieved = ALGO_BYTES_PER_STEP * n / per_launch_s / 1e9
        traffic, traffic_src = measured_traffic()
        if traffic is not None and n != int(traffic_src.get("n", n)):
            traffic = traffic * n / float(traffic_src["n"])      # per-aircraft traffic is size independent
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(n),
            "launch": dict(info, coef_cache=not args.no_cache),
            "e2e": {"value": e2e_value, "unit": UNIT, "steps": Ke, "ms_per_step": 1e3 * e2e_s / Ke,
                    "h2d_bytes_per_step": n * 4 * 4, "d2h_bytes_per_step": n * 22 * 4 + n * 4 + 3 * n,
                    "boundary": venv_boundary,
                    "how": "GPUVecEnv.step(numpy) -> numpy through pinned host buffers (bytes counted from those buffers): 'pipelined' "
                           "= upload / kernel / download of aircraft chunks on three streams (np_env_step_host), 'mapped' = the "
                           "kernel reads / writes device-mapped host memory directly; wall clock, max over ranks"},
            "gpu_launches": K,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic,
                         "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "traffic_source": (traffic_src or {}).get("source"),
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_STEP * n, "peak_source": peak_src,
                         "kernel": "f16_step_kernel", "algorithmic_bytes_per_aircraft_step": ALGO_BYTES_PER_STEP,
                         "fp32_note": "the step is fp32-pipe bound, not HBM bound (DESIGN.md section 3): FFMA2 is half rate, so the "
                                      "7 550 MACs/aircraft-step that remain after the exact table conversion cap K1 at "
                                      "about 4.9e9 aircraft-steps/s (148 SM x 128 MAC/clk x 1.965 GHz)",
                         "fma_pipe_frac": (7550.0 * n / per_launch_s) / (148 * 128 * 1.965e9),
                         "reference_equivalent_gflops": ALGO_FLOP_PER_STEP * n / per_launch_s / 1e9},
            "strong": strong,
            "side": side,
            "termination_counters": counters,
        }
        out["side_rooflines"] = [side[k] for k in ("uav_roofline", "table_backend") if k in side and "error" not in side[k]]
        if world == 1 and not args.no_cpu:
            try:
                nc = n if reference_available() else args.cpu_n
                v, el, kind, thr, sample = cpu_arm(nc, 1, 5 if reference_available() else 50)
                out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": thr, "kind": kind, "sample": sample}
            except Exception as e:
                out["cpu_baseline"] = {"error": repr(e)[:300]}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
