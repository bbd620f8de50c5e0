#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Inputs are the deterministic tapes of oracle/tapes.py (so the fixtures hold outputs only);
outputs are what the reference's own ControlEnv / F16Model / F16Dynamics return on CPU.
Also snapshots the head of the reference's recorded trajectory renders/result/*.npy
(the only trajectory fixture the reference ships).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from ref_harness import REF_ROOT, RefEnv, import_reference  # noqa: E402
from oracle import tapes  # noqa: E402

CHECKPOINTS = [1, 2, 3, 5, 10, 20, 50, 100, 200, 300, 500, 700, 1000]


def snap(ref, obs, rew, done, bad, exc):
    sn = ref.snapshot()
    tgt = [sn[k] for k in ("target_altitude", "target_heading", "target_vt") if k in sn] \
        if "target_heading" in sn and "target_pitch" not in sn and "target_npos" not in sn else None
    if "target_pitch" in sn:
        tgt = [sn["target_pitch"], sn["target_heading"], sn["target_vt"]]
    if "target_npos" in sn:
        tgt = [sn["target_npos"], sn["target_epos"], sn["target_altitude"]]
    return dict(s=sn["s"].numpy().copy(), u=sn["u"].numpy().copy(), tgt=torch.stack(tgt, 1).numpy(),
                step_count=sn["step_count"].numpy().astype(np.int32), obs=obs.numpy().copy(),
                reward=rew.numpy().copy(), done=done.numpy().copy(), bad=bad.numpy().copy(), exc=exc.numpy().copy())


def trajectory(task, n, steps, scale, seed, name):
    """Reference trajectory + the reference's own sensitivity: a twin run whose state is perturbed by 1 ulp
    (s * (1 + 2^-23)) after every episodic reset; the twin's distance to the unperturbed run at each checkpoint is
    the fp32 noise level any re-implementation should be judged against."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from _metrics import state_rel_err
    ref = RefEnv(n, task, "F16", seed=0, noise_scale=0.0)
    twin = RefEnv(n, task, "F16", seed=0, noise_scale=0.0)
    d0 = torch.from_numpy(tapes.reset_draw_tape(seed, 0, n))
    obs0 = ref.reset(d0)
    twin.reset(d0)
    twin.env.model.s = twin.env.model.s * (1.0 + 2.0 ** -23)
    out = {"obs0": obs0.numpy().copy(), "meta": np.array([n, steps, seed], dtype=np.int64), "scale": np.float32(scale)}
    n_bad, n_done, rsum = [], [], []
    for k in range(1, steps + 1):
        a = torch.from_numpy(tapes.action_tape(seed, k, n, scale))
        d = torch.from_numpy(tapes.reset_draw_tape(seed, k, n))
        obs, rew, done, bad, exc = ref.step(a, d)
        twin.step(a, d, perturb_after_reset=1.0 + 2.0 ** -23)
        if k in CHECKPOINTS:
            same = (twin.env.step_count == ref.env.step_count).numpy()
            e = state_rel_err(twin.env.model.s.numpy()[same], ref.env.model.s.numpy()[same])
            out[f"k{k}_selfnoise"] = np.array([np.median(e), e.max(), same.mean()], dtype=np.float64)
        n_bad.append(int(bad.sum())); n_done.append(int(done.sum())); rsum.append(float(rew.double().sum()))
        if k in CHECKPOINTS:
            for key, v in snap(ref, obs, rew, done, bad, exc).items():
                out[f"k{k}_{key}"] = v
    out["n_bad"] = np.array(n_bad, dtype=np.int32)
    out["n_done"] = np.array(n_done, dtype=np.int32)
    out["reward_sum"] = np.array(rsum, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "bad events", sum(n_bad), "done events", sum(n_done))


def add_truth(name, task):
    """Append the float64 'truth' trajectory to a trajectory fixture: the oracle restatement (bit-identical to the
    reference in float32, tests/test_oracle_golden.py) run in double on the same tapes.  Stored per checkpoint:
    k{k}_s64 and k{k}_ref_vs_truth = (median, max) of the reference-fp32 state error against it."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from _metrics import state_rel_err
    from oracle.f16_oracle import F16EnvOracle
    path = os.path.join(HERE, name)
    g = dict(np.load(path))
    n, steps, seed = [int(x) for x in g["meta"]]
    scale = float(g["scale"])
    o64 = F16EnvOracle(n, task, dtype=torch.float64)
    o64.reset(torch.from_numpy(tapes.reset_draw_tape(seed, 0, n)).double())
    for k in range(1, steps + 1):
        o64.step(torch.from_numpy(tapes.action_tape(seed, k, n, scale)).double(),
                 torch.from_numpy(tapes.reset_draw_tape(seed, k, n)).double())
        if f"k{k}_s" in g:
            same = o64.step_count.numpy() == g[f"k{k}_step_count"]
            e = state_rel_err(g[f"k{k}_s"][same], o64.s.numpy()[same])
            g[f"k{k}_s64"] = o64.s.numpy().copy()
            g[f"k{k}_step_count64"] = o64.step_count.numpy().astype(np.int32)
            g[f"k{k}_ref_vs_truth"] = np.array([np.median(e), e.max(), same.mean()])
    np.savez_compressed(path, **g)
    print(name, "truth added; ref-vs-truth median at last checkpoint %.2e" % g[f"k{max(c for c in CHECKPOINTS if c <= steps)}_ref_vs_truth"][0])


def done_branch(name):
    """Drive the target-reached (`done`) branch of UnreachHeading (unreach_heading.py:38-53), which random
    actions never reach: after the initial reset put every target on the current state and jump step_count
    to just below min_check_interval; record the next 4 steps (done fires, +200, then the reset)."""
    n, seed = 32, 21
    ref = RefEnv(n, "heading", "F16", seed=0, noise_scale=0.0)
    ref.reset(torch.from_numpy(tapes.reset_draw_tape(seed, 0, n)))
    e = ref.env
    e.task.target_altitude[:] = e.model.s[:, 2]
    e.task.target_heading[:] = e.model.s[:, 5]
    e.task.target_vt[:] = e.model.s[:, 6]
    e.task.target_altitude[::4] += 500.0          # a quarter stay off-target
    e.step_count[:] = 298
    e.step_count[1::8] = 2499                      # late but on-target: neither done nor bad_done
    e.step_count[::8] = 2499                       # late and off-target: bad_done (unreach heading)
    out = {"meta": np.array([n, 4, seed], dtype=np.int64)}
    for k in range(1, 5):
        a = torch.from_numpy(tapes.action_tape(seed, k, n, 0.02))
        d = torch.from_numpy(tapes.reset_draw_tape(seed, k, n))
        obs, rew, done, bad, exc = ref.step(a, d)
        for key, v in snap(ref, obs, rew, done, bad, exc).items():
            out[f"k{k}_{key}"] = v
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "done at k2:", int(out["k2_done"].sum()), "bad at k1:", int(out["k1_bad"].sum()))


def nlplant_kat(name, n=2048, seed=5):
    """Known answers of F16Dynamics.nlplant, the 43 coefficient nets and the derived getters on random states."""
    import_reference()
    from models.F16_model import F16Model
    cfg = type("Cfg", (), {"init_state": {"init_T": 2000}})
    m = F16Model(cfg, n, torch.device("cpu"), 0)
    s, u = tapes.random_envelope_states(seed, n)
    m.s = torch.from_numpy(s); m.u = torch.from_numpy(u)
    with torch.no_grad():
        xdot = m.get_extended_state()
        ax, ay, az = m.get_acceleration()
        h = m.dynamics.hifi_F16
        alpha, beta, el = m.s[:, 7] * (180.0 / torch.pi), m.s[:, 8] * (180.0 / torch.pi), m.u[:, 1]
        d = np.load(os.path.join(HERE, "..", "..", "neuralplane_b200", "data", "f16_aero.npz"))
        meths = {a.lower(): a for a in dir(h) if a.startswith("_") and not a.startswith("__")}
        coefs = []
        for nm, row in zip(d["names"], d["desc"]):
            nm = str(nm)
            meth = meths["_" + nm.lower()]
            args = [(alpha, beta, el)[int(c)] for c in row[1:1 + int(row[0])]]
            coefs.append(getattr(h, meth)(*args).numpy())
        out = dict(xdot=xdot[:, :12].numpy(), accel=torch.stack((ax, ay, az), 1).numpy(),
                   coefs=np.stack(coefs, 1), eas2tas=m.get_EAS2TAS().numpy(), G=m.get_G().numpy(),
                   meta=np.array([n, seed], dtype=np.int64))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, out["xdot"].shape, out["coefs"].shape)


def recorded_trajectory(name, head=400):
    """Head of the reference's recorded F-16 trajectory (writer: renders/render_ppo.py:37,98-102,153-186)."""
    keys = ["npos", "epos", "altitude", "roll", "pitch", "yaw", "vt", "alpha", "beta", "G", "T", "el", "ail", "rud"]
    out = {k: np.load(os.path.join(REF_ROOT, "renders", "result", k + ".npy"))[:head].astype(np.float32) for k in keys}
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: v.shape for k, v in out.items()}["npos"])


def eval_metrics_golden(name, head=400):
    """The nine flight metrics of renders/evaluate_result.py:29-43, computed by executing THOSE reference lines on the
    head of the reference's own recording (the float32 snapshot of ref_recorded_trajectory.npz, so the test input is
    exactly the fixture)."""
    src = open(os.path.join(REF_ROOT, "renders", "evaluate_result.py"), encoding="utf-8").read().splitlines()
    lines = [ln for ln in src[28:43] if "=" in ln and not ln.lstrip().startswith("#")]
    rec = np.load(os.path.join(HERE, "ref_recorded_trajectory.npz"))
    ns = {"np": np}
    for k in ("G", "vt", "pitch", "alpha", "beta", "altitude"):
        ns[k + "_buf"] = rec[k][:head].astype(np.float64)
    exec("\n".join(lines), ns)
    names = ("G", "TAS", "RoC", "AOA", "ASM", "SSM", "OSM", "AOASM", "AOSSM")
    np.savez_compressed(os.path.join(HERE, name), names=np.array(names), values=np.array([ns[k] for k in names], dtype=np.float64),
                        head=np.array([head]))
    print(name, {k: float(ns[k]) for k in names})


def rollout_golden(name, T=12, N=6, A=2, D=22, act=4, seed=23):
    """The reference's OWN ReplayBuffer (algorithms/utils/buffer.py) driven through insert / compute_returns /
    after_update for the four (use_gae, use_proper_time_limits) variants, with masks derived by the reference runner's
    insert logic (runner/F16sim_runner.py:141-157, executed on a stand-in object carrying the attributes it reads)."""
    import types
    import_reference()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import gym.spaces
    from algorithms.utils.buffer import ReplayBuffer
    rng = np.random.default_rng(seed)
    obs_space = gym.spaces.Box(low=-10, high=10., shape=(D,))
    act_space = gym.spaces.Box(low=-1, high=1., shape=(act,))
    tape = {
        "obs0": rng.standard_normal((N, A, D)).astype(np.float32),
        "obs": rng.standard_normal((T, N, A, D)).astype(np.float32),
        "actions": rng.uniform(-1, 1, (T, N, A, act)).astype(np.float32),
        "rewards": rng.standard_normal((T, N, A, 1)).astype(np.float32),
        "dones": rng.random((T, N, A, 1)) < 0.08,
        "bad_dones": rng.random((T, N, A, 1)) < 0.10,
        "exceed": rng.random((T, N, A, 1)) < 0.03,
        "logp": rng.standard_normal((T, N, A, 1)).astype(np.float32),
        "values": rng.standard_normal((T, N, A, 1)).astype(np.float32),
        "rnn_a": rng.standard_normal((T, N, A, 1, 8)).astype(np.float32),
        "rnn_c": rng.standard_normal((T, N, A, 1, 8)).astype(np.float32),
        "next_value": rng.standard_normal((N, A, 1)).astype(np.float32),
    }
    out = {"meta": np.array([T, N, A, D, act, seed]), "gamma": np.array([0.99]), "gae_lambda": np.array([0.95])}
    out.update({"tape_" + k: v for k, v in tape.items()})
    # the reference runner's insert(), bound to a stand-in `self`
    from runner.F16sim_runner import F16SimRunner
    for use_gae in (True, False):
        for proper in (True, False):
            args = types.SimpleNamespace(buffer_size=T, n_rollout_threads=N, gamma=0.99, use_proper_time_limits=proper, use_gae=use_gae,
                                         gae_lambda=0.95, recurrent_hidden_size=8, recurrent_hidden_layers=1)
            buf = ReplayBuffer(args, A, obs_space, act_space)
            buf.obs[0] = tape["obs0"].copy()
            runner = types.SimpleNamespace(buffer=buf, n_rollout_threads=N, num_agents=A)
            for t in range(T):
                F16SimRunner.insert(runner, [tape["obs"][t], tape["actions"][t], tape["rewards"][t], tape["dones"][t], tape["bad_dones"][t],
                                             tape["exceed"][t], tape["logp"][t], tape["values"][t], tape["rnn_a"][t].copy(), tape["rnn_c"][t].copy()])
            buf.compute_returns(tape["next_value"])
            key = f"gae{int(use_gae)}_proper{int(proper)}_"
            out[key + "returns"] = buf.returns.copy()
            out[key + "advantages"] = buf.advantages.copy()
            if use_gae and proper:
                for k in ("obs", "actions", "rewards", "masks", "bad_masks", "action_log_probs", "value_preds", "rnn_states_actor", "rnn_states_critic"):
                    out["buf_" + k] = getattr(buf, k).copy()
                out["buf_step_after"] = np.array([buf.step])
                torch.manual_seed(5)                                    # recurrent_generator draws torch.randperm
                for bi, batch in enumerate(ReplayBuffer.recurrent_generator(buf, 2, 4)):
                    for name_, arr in zip(("obs", "actions", "masks", "logp", "adv", "returns", "values", "rnn_a", "rnn_c"), batch):
                        out[f"gen{bi}_{name_}"] = np.asarray(arr).copy()
                buf.after_update()
                out["after_obs0"], out["after_masks0"], out["after_bad_masks0"] = buf.obs[0].copy(), buf.masks[0].copy(), buf.bad_masks[0].copy()
                out["after_rnn_a0"] = buf.rnn_states_actor[0].copy()
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: v.shape for k, v in out.items() if k.endswith("returns")})


class RefPidPlanner:
    """PlanningEnv.step (envs/planning_env.py:144-177) assembled from the reference's OWN parts, with the low-level
    GRU PPO actor (whose checkpoint is not in the repository, planning_env.py:16) replaced by the reference's PID
    stack: RollController / PitchController / YawController (`Controller.stabilize`, algorithms/pid/controller.py:
    69-74), `L1Controller.update_heading_hold` + `nav_roll` for the heading target (controller.py:114-124), the
    pitch target fed straight to the pitch loop, and a TAS loop built from the reference `PID` class with the gains
    of algorithms/pid/config/speedcontroller.yaml (the reference's SpeedController class itself is not runnable:
    it reads attributes it never defines, speedController.py:24-45).  Everything that is stepped -- F16Model,
    TrackingTask, terminations, rewards, the controllers -- is unmodified reference code."""

    N_SUB = 50

    def __init__(self, n):
        import_reference()
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        from algorithms.pid.controller import Controller
        import algorithms.pid.pid as pid
        self.ref = RefEnv(n, "tracking", "F16", seed=0, noise_scale=0.0)
        env = self.ref.env
        with contextlib.redirect_stdout(io.StringIO()):
            self.ctl = Controller(dt=env.model.dt, n=n, device="cpu")
        self.speed_pid = pid.PID(Kp=5, Ki=25, Kd=0, Kff=80, Kimax=100, dt=env.model.dt, n=n, device="cpu")
        self.speed_last_out = torch.zeros((n, 1))

    def step(self, action, draws):
        ref, env, ctl = self.ref, self.ref.env, self.ctl
        with contextlib.redirect_stdout(io.StringIO()):
            with ref._inject(draws):
                env.reset()                                                     # :145
            action = torch.clamp(action, -1, 1)
            roll, pitch, yaw = env.model.get_posture()
            vt = env.model.get_vt()
            target_pitch = pitch + action[:, 0] * 0.3                           # :150-152
            target_heading = yaw + action[:, 1] * 0.3
            target_vt = vt + action[:, 2] * 30
            for i in range(self.N_SUB):
                ctl.pitch_dem = target_pitch.reshape(-1, 1)
                ctl.update_heading_hold(target_heading.reshape(-1, 1), env)
                TAS = env.model.get_TAS().reshape(-1, 1)
                limit = torch.abs(self.speed_last_out) >= 100
                self.speed_pid.update_all(target_vt.reshape(-1, 1) * 0.3048 / 340, TAS * 0.3048 / 340, limit)
                out = self.speed_pid.get_ff() + self.speed_pid.get_p() + self.speed_pid.get_i() + self.speed_pid.get_d()
                self.speed_last_out = out
                ctl.throttle_dem = torch.clamp(out / 100, 0, 1)
                ctl.stabilize(env)
                ego_actions = ctl.get_action()
                env.model.update(ego_actions)                                   # :160
                reset = (env.is_done.bool() | env.bad_done.bool()) | env.exceed_time_limit.bool()
                env.model.s[reset] = env.model.recent_s[reset]                  # :162-166
                env.step_count += 1
                obs = env.obs()
                done, bad_done, exceed, _ = env.done({})
                reward = env.reward()
        self.targets = torch.stack((target_pitch, target_heading, target_vt), 1)
        return obs, reward, done, bad_done, exceed

    def pid_state(self):
        c = self.ctl
        rows = []
        for p, last in ((c.roll_controller.rate_pid, c.roll_controller.last_out), (c.pitch_controller.rate_pid, c.pitch_controller.last_out),
                        (c.yaw_controller.rate_pid, c.yaw_controller.last_out), (self.speed_pid, self.speed_last_out)):
            rows += [p.error.reshape(-1), p.integrator.reshape(-1), last.reshape(-1)]
        return torch.stack(rows, 1)


def planning_trajectory(n, steps, scale, seed, name):
    """RefPidPlanner trajectory: one record per planning step (= 50 reference sub-steps)."""
    pl = RefPidPlanner(n)
    ref = pl.ref
    obs0 = ref.reset(torch.from_numpy(tapes.reset_draw_tape(seed, 0, n)))
    out = {"obs0": obs0.numpy().copy(), "meta": np.array([n, steps, seed], dtype=np.int64), "scale": np.float32(scale)}
    for k in range(1, steps + 1):
        a = torch.from_numpy(tapes.action_tape(seed, k, n, scale, num_actions=3))
        d = torch.from_numpy(tapes.reset_draw_tape(seed, k, n))
        obs, rew, done, bad, exc = pl.step(a, d)
        for key, v in snap(ref, obs, rew, done, bad, exc).items():
            out[f"k{k}_{key}"] = v
        out[f"k{k}_pid"] = pl.pid_state().numpy().copy()
        out[f"k{k}_targets"] = pl.targets.numpy().copy()
        print(name, "planning step", k, "bad", int(bad.sum()), "done", int(done.sum()), flush=True)
    np.savez_compressed(os.path.join(HERE, name), **out)


def geodetic_golden(name):
    """Values of the reference's enu_to_geodetic (envs/utils/utils.py:140-142) on 200 random ENU points."""
    import_reference()
    from utils.utils import enu_to_geodetic as ref_e2g
    rng = np.random.RandomState(0)
    pts = np.concatenate([rng.uniform(-3e5, 3e5, (200, 2)), rng.uniform(0, 2e4, (200, 1))], 1)
    out = np.array([ref_e2g(float(e), float(n), float(u), 0, 0, 0) for e, n, u in pts])
    np.savez_compressed(os.path.join(HERE, name), enu=pts, llh=out)
    print(name, out.shape)


class RefCombat:
    """SingleCombatEnv (envs/singlecombat_env.py) is stale at this commit and cannot be constructed,
    so its step is re-assembled here from the reference code that still runs, called UNMODIFIED:
      * SingleCombatEnv.obs / SingleCombatEnv.reward   -- the unbound methods (singlecombat_env.py:64-181) on this object
      * get_AO_TA_R, orientation_fn, distance_fn        -- envs/utils/utils.py:156-249 (blood model, :263-271)
      * Overload, LowAltitude, HighSpeed, LowSpeed, ExtremeState, Crash, Timeout, Shutdown -- termination classes
      * F16Model.update / getters, Controller.stabilize -- the current model plug-in and PID stack
    Orchestration re-derived from singlecombat_env.py:183-274 (documented in oracle/combat_oracle.py): env-level reset
    of a pair when either aircraft is flagged, 5 FDM sub-steps per env step with the demand low-pass, the ORIGINAL
    4-D action kept for all 5 sub-steps (the reference overwrites it inside the loop, App. D.10), each sub-step applied
    through F16Model.update([a0, -el/45, -ail/45, -rud/45]) (the current plug-in API, as renders/render_control.py
    does), flags OR-accumulated, blood updated after the 5th sub-step."""

    def __init__(self, num_envs):
        import_reference()
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        import envs.singlecombat_env as sc
        from envs.utils.utils import parse_config
        from models.F16_model import F16Model
        from algorithms.pid.controller import Controller
        from termination_conditions.overload import Overload
        from termination_conditions.low_altitude import LowAltitude
        from termination_conditions.high_speed import HighSpeed
        from termination_conditions.low_speed import LowSpeed
        from termination_conditions.extreme_state import ExtremeState
        from termination_conditions.crash import Crash
        from termination_conditions.timeout import Timeout
        from termination_conditions.shutdown import Shutdown
        self.sc = sc
        self.config = parse_config("selfplay")
        self.config.init_state = {"init_T": getattr(self.config, "init_T", 2000)}
        self.device = torch.device("cpu")
        self.num_envs, self.num_agents = num_envs, 2
        self.n = 2 * num_envs
        self.target_dist = getattr(self.config, "target_dist", 3)
        self.model = F16Model(self.config, self.n, self.device, 0)
        with contextlib.redirect_stdout(io.StringIO()):
            self.controller = Controller(dt=self.model.dt, n=self.n, device="cpu")
        self.conditions = [Overload(self.config), LowAltitude(self.config), HighSpeed(self.config), LowSpeed(self.config),
                           ExtremeState(self.config), Crash(self.config, "cpu"), Timeout(self.config), Shutdown(self.config, "cpu")]
        self.blood = 100 * torch.ones(self.n)
        self.step_count = torch.zeros(self.n, dtype=torch.int64)
        self.is_done = torch.ones(self.n, dtype=torch.bool)          # force the initial reset of everyone
        self.bad_done = torch.ones(self.n, dtype=torch.bool)
        self.exceed_time_limit = torch.ones(self.n, dtype=torch.bool)

    # attributes the stale methods read (singlecombat_env.py:88-119, :142-147)
    @property
    def s(self):
        return self.model.s

    @property
    def velocity(self):
        return torch.stack(self.model.get_velocity(), 1)

    @property
    def es(self):
        return self.model.get_extended_state()

    def reset_done_envs(self, draws):
        """singlecombat_env.py:207-238 with the draws taken from the tape: columns npos, epos, altitude, heading, vt."""
        c = self.config
        flag = (self.is_done | self.bad_done) | self.exceed_time_limit
        env_reset = flag.reshape(self.num_envs, 2).any(dim=1)
        m = env_reset.repeat_interleave(2)
        d = draws
        self.model.s[m, :] = 0
        self.model.u[m, :] = 0
        self.model.s[m, 0] = d[m, 0] * (c.max_npos - c.min_npos) + c.min_npos
        self.model.s[m, 1] = d[m, 1] * (c.max_epos - c.min_epos) + c.min_epos
        self.model.s[m, 2] = d[m, 2] * (c.max_altitude - c.min_altitude) + c.min_altitude
        self.model.s[m, 5] = d[m, 3] * (c.max_heading - c.min_heading) + c.min_heading
        self.model.s[m, 6] = d[m, 4] * (c.max_vt - c.min_vt) + c.min_vt
        self.model.u[m, 0] = c.init_T
        self.blood[m] = 100
        self.step_count[m] = 0
        self.is_done[:] = False
        self.bad_done[:] = False
        self.exceed_time_limit[:] = False

    def reset(self, draws):
        self.is_done[:] = True
        with contextlib.redirect_stdout(io.StringIO()):
            self.reset_done_envs(draws)
            return self.sc.SingleCombatEnv.obs(self)

    def step(self, action, draws):
        ctl = self.controller
        with contextlib.redirect_stdout(io.StringIO()):
            self.reset_done_envs(draws)
            a = torch.clamp(action, -1, 1)
            for i in range(5):
                ctl.roll_dem = 0.9 * ctl.roll_dem + 0.1 * a[:, 1].reshape(-1, 1) * 4 * torch.pi / 9     # :246
                ctl.pitch_dem = 0.9 * ctl.pitch_dem + 0.1 * a[:, 2].reshape(-1, 1) * torch.pi / 12      # :247
                ctl.stabilize(self)                                                                    # :251
                ego = torch.hstack((a[:, 0].reshape(-1, 1), -ctl.el / 45, -ctl.ail / 45, -ctl.rud / 45))
                self.model.update(ego)
                self.step_count += 1
                for cond in self.conditions:
                    bad, done, exc, _ = cond.get_termination(None, self, {})
                    self.bad_done = self.bad_done | bad
                    self.is_done = self.is_done | done
                    self.exceed_time_limit = self.exceed_time_limit | exc
            obs = self.sc.SingleCombatEnv.obs(self)
            reward = self.sc.SingleCombatEnv.reward(self)
            from envs.utils.utils import get_AO_TA_R, orientation_fn, distance_fn                         # :263-271
            ego_i = torch.arange(self.num_envs) * 2
            enm_i = ego_i + 1
            es = self.es
            AO, TA, R = get_AO_TA_R(self.s[ego_i, :3], self.s[enm_i, :3], es[ego_i, :3], es[enm_i, :3])
            self.blood[enm_i] -= orientation_fn(AO) * distance_fn(R * 0.3048 / 1000)
            self.blood[ego_i] -= orientation_fn(torch.pi - TA) * distance_fn(R * 0.3048 / 1000)
        return obs, reward, self.is_done.clone(), self.bad_done.clone(), self.exceed_time_limit.clone()

    def ctrl_state(self):
        c = self.controller
        rows = [c.roll_dem.reshape(-1), c.pitch_dem.reshape(-1)]
        for p, last in ((c.roll_controller.rate_pid, c.roll_controller.last_out), (c.pitch_controller.rate_pid, c.pitch_controller.last_out),
                        (c.yaw_controller.rate_pid, c.yaw_controller.last_out)):
            rows += [p.error.reshape(-1), p.integrator.reshape(-1), last.reshape(-1)]
        return torch.stack(rows, 1)


def combat_trajectory(num_envs, steps, seed, name, close=False):
    """RefCombat trajectory.  close=True re-positions every pair nose-to-tail at 1-3 km after the first reset so that
    the blood / Shutdown / Crash branches fire within the fixture."""
    rc = RefCombat(num_envs)
    n = rc.n
    obs0 = rc.reset(torch.from_numpy(tapes.reset_draw_tape(seed, 0, n)))
    if close:
        ego, enm = torch.arange(num_envs) * 2, torch.arange(num_envs) * 2 + 1
        gap = torch.linspace(100.0, 9000.0, num_envs)
        rc.model.s[enm, 0] = rc.model.s[ego, 0] + gap
        rc.model.s[enm, 1] = rc.model.s[ego, 1] + 0.02 * gap
        rc.model.s[enm, 2] = rc.model.s[ego, 2] + 10.0
        rc.model.s[:, 5] = 0.0
        rc.blood[enm[::3]] = 0.3                               # nearly shot down: Shutdown(done) fires soon
        rc.blood[ego[1::7]] = 0.2
    out = {"obs0": obs0.numpy().copy(), "meta": np.array([num_envs, steps, seed], dtype=np.int64),
           "s_start": rc.model.s.numpy().copy(), "blood_start": rc.blood.numpy().copy()}
    for k in range(1, steps + 1):
        a = torch.from_numpy(tapes.action_tape(seed, k, n, 1.0 if not close else 0.2))
        d = torch.from_numpy(tapes.reset_draw_tape(seed, k, n))
        obs, rew, done, bad, exc = rc.step(a, d)
        out[f"k{k}_s"] = rc.model.s.numpy().copy(); out[f"k{k}_u"] = rc.model.u.numpy().copy()
        out[f"k{k}_obs"] = obs.numpy().copy(); out[f"k{k}_reward"] = rew.numpy().copy()
        out[f"k{k}_done"] = done.numpy().copy(); out[f"k{k}_bad"] = bad.numpy().copy(); out[f"k{k}_exc"] = exc.numpy().copy()
        out[f"k{k}_blood"] = rc.blood.numpy().copy(); out[f"k{k}_step_count"] = rc.step_count.numpy().astype(np.int32)
        out[f"k{k}_ctrl"] = rc.ctrl_state().numpy().copy()
        print(name, "step", k, "bad", int(bad.sum()), "done", int(done.sum()), "min blood %.2f" % float(rc.blood.min()), flush=True)
    np.savez_compressed(os.path.join(HERE, name), **out)


def uav_trajectory(task, n, steps, scale, seed, name):
    """ControlEnv(model='UAV') from the unmodified reference with num_controls = 3 (the one fix it needs to survive
    its second step): state / obs / reward / flags at every checkpoint."""
    ref = RefEnv(n, task, "UAV", seed=0, noise_scale=0.0)
    ref.env.model.num_controls = 3
    ref.env.model.u = torch.zeros(n, 3)
    obs0 = ref.reset(torch.from_numpy(tapes.reset_draw_tape(seed, 0, n)))
    out = {"obs0": obs0.numpy().copy(), "meta": np.array([n, steps, seed], dtype=np.int64), "scale": np.float32(scale)}
    n_bad = []
    for k in range(1, steps + 1):
        a = torch.from_numpy(tapes.action_tape(seed, k, n, scale))
        d = torch.from_numpy(tapes.reset_draw_tape(seed, k, n))
        obs, rew, done, bad, exc = ref.step(a, d)
        n_bad.append(int(bad.sum()))
        if k in CHECKPOINTS or k == steps:
            for key, v in snap(ref, obs, rew, done, bad, exc).items():
                out[f"k{k}_{key}"] = v
    out["n_bad"] = np.array(n_bad, dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "bad events", sum(n_bad))


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "combat":       # regenerate only the combat fixtures
        combat_trajectory(24, 30, 18, "combat_traj.npz")
        combat_trajectory(24, 30, 19, "combat_close_traj.npz", close=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "geodetic":     # regenerate only the geodetic (acmi) fixture
        geodetic_golden("geodetic_golden.npz")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "rollout":      # regenerate only the rollout-buffer fixture
        rollout_golden("rollout_returns.npz")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "metrics":      # regenerate only the flight-metrics fixture
        eval_metrics_golden("eval_metrics.npz")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "planning":     # regenerate only the planning fixture
        planning_trajectory(48, 16, 1.0, 17, "planning_pid_traj.npz")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "uav":          # regenerate only the UAV fixtures
        uav_trajectory("control", 64, 400, 1.0, 15, "uav_control_traj.npz")
        uav_trajectory("heading", 64, 400, 1.0, 16, "uav_heading_traj.npz")
        sys.exit(0)
    recorded_trajectory("ref_recorded_trajectory.npz")
    eval_metrics_golden("eval_metrics.npz")
    nlplant_kat("f16_nlplant_kat.npz")
    done_branch("heading_done_branch.npz")
    for task, n, steps, scale, seed, name in (("heading", 128, 1000, 0.3, 11, "heading_traj_a03.npz"),
                                              ("heading", 128, 1000, 1.0, 12, "heading_traj_a10.npz"),
                                              ("control", 64, 300, 1.0, 13, "control_traj.npz"),
                                              ("tracking", 64, 300, 1.0, 14, "tracking_traj.npz")):
        trajectory(task, n, steps, scale, seed, name)
        add_truth(name, task)
    uav_trajectory("control", 64, 400, 1.0, 15, "uav_control_traj.npz")
    uav_trajectory("heading", 64, 400, 1.0, 16, "uav_heading_traj.npz")
    planning_trajectory(48, 16, 1.0, 17, "planning_pid_traj.npz")
    combat_trajectory(24, 30, 18, "combat_traj.npz")
    combat_trajectory(24, 30, 19, "combat_close_traj.npz", close=True)
    geodetic_golden("geodetic_golden.npz")
    rollout_golden("rollout_returns.npz")
