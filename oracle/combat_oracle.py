"""CPU oracle for the 1-v-1 combat step.  TEST INFRASTRUCTURE ONLY.

The reference's SingleCombatEnv (envs/singlecombat_env.py) is stale at the surveyed commit: it cannot be constructed
(positional-argument mix-up, attributes BaseEnv no longer has -- SURVEY section 0), so there is no runnable
reference STEP.  What still runs is pinned: tests/golden/combat*_traj.npz come from a harness
(tests/golden/make_golden.py RefCombat) that calls, unmodified, the reference's SingleCombatEnv.obs and .reward
methods, get_AO_TA_R / orientation_fn / distance_fn, all eight termination classes, F16Model and Controller.stabilize.
The ORCHESTRATION is re-derived from singlecombat_env.py:183-274 -- **parity unpinned for the orchestration**:
  1. pairs are adjacent, ego = 2e, enemy = 2e + 1 (:98-99, crash.py:32-33);
  2. env-level reset (:207-238): a pair is re-initialised when either aircraft carries a flag; draws in the order
     npos, epos, altitude, heading, vt (:220-224); u = [init_T, 0, 0, 0, 0]; blood = 100; step_count = 0; then ALL
     flags are cleared; the controller state is not reset (the reference never resets it);
  3. step (:240-274): clamp the 4-D action [throttle, roll_dem, pitch_dem, yaw]; 5 FDM sub-steps, each: demand
     low-pass roll_dem <- 0.9 roll_dem + 0.1 a1 4pi/9, pitch_dem <- 0.9 pitch_dem + 0.1 a2 pi/12 (:246-247);
     Controller.stabilize with yaw_rate_dem = 0 (a3 only sets `yaw_dem`, which the current stabilize_yaw ignores,
     controller.py:60-65); surfaces sign-flipped (:253-255) and applied through the current plug-in API
     F16Model.update([a0, -el/45, -ail/45, -rud/45]) (the T formula of :252 is exactly update's throttle lag);
     step_count += 1; the eight terminations OR-accumulate;
     FIXED: the reference overwrites `action` with the 5-D control vector inside the loop (App. D.10); the original
     4-D action is used for all five sub-steps here;
  4. obs (:64-138) and reward (:140-181) at the final state; then the blood model (:263-271).
"""
import torch

from .f16_oracle import AeroNets, eas2tas, euler_step, lowpass_controls, nlplant, wrap_pi, body_accel
from .planning_oracle import GAINS, _Pid

COMBAT_CFG = dict(  # envs/configs/selfplay.yaml
    dt=0.02, airspeed=0, altitude_limit=2500.0, acceleration_limit=600.0, max_velocity=3, min_velocity=0.01,
    min_alpha=-20, max_alpha=45, min_beta=-30, max_beta=30, distance_limit=200, max_steps=2000, init_T=2000,
    target_dist=3, max_altitude=20000, min_altitude=19000, max_vt=1200, min_vt=1000, max_heading=0.5, min_heading=-0.5,
    max_npos=5000, min_npos=-5000, max_epos=5000, min_epos=-5000,
)


def ao_ta_r(ego_pos, enm_pos, ego_vel, enm_vel, two_d=False):
    """get_AO_TA_R / get2d_AO_TA_R (envs/utils/utils.py:156-206) incl. the side flag."""
    if two_d:
        ego_pos, enm_pos, ego_vel, enm_vel = ego_pos[:, :-1], enm_pos[:, :-1], ego_vel[:, :-1], enm_vel[:, :-1]
    ego_v = torch.linalg.norm(ego_vel, dim=1)
    enm_v = torch.linalg.norm(enm_vel, dim=1)
    delta = enm_pos - ego_pos
    dist = torch.linalg.norm(delta, dim=1)
    AO = torch.arccos(torch.clamp(torch.sum(delta * ego_vel, dim=1) / (dist * ego_v + 1e-8), -1, 1))
    TA = torch.arccos(torch.clamp(torch.sum(delta * enm_vel, dim=1) / (dist * enm_v + 1e-8), -1, 1))
    side = torch.sign(ego_vel[:, 0] * delta[:, 1] - ego_vel[:, 1] * delta[:, 0])      # z of cross(v_xy, dpos_xy)
    return AO, TA, dist, side


def orientation_reward_v2(AO, TA):
    """utils.py:215-217."""
    return 1 / (50 * AO / torch.pi + 2) + 1 / 2 \
        + torch.min((torch.arctanh(1 - torch.max(1.9 * TA / torch.pi, 1e-4 * torch.ones_like(TA)))) / (2 * torch.pi),
                    torch.zeros_like(TA)) + 0.5


def range_reward_v3(target_dist, R):
    """utils.py:230-231 (target_dist is unused by v3)."""
    return 1 * (R < 5) + (R >= 5) * torch.clamp(-0.032 * R ** 2 + 0.284 * R + 0.38, 0, 1) + torch.clamp(torch.exp(-0.16 * R), 0, 0.2)


def orientation_fn(AO):
    """utils.py:235-243."""
    m3 = (AO >= 0) & (AO <= torch.pi / 6)
    m4 = (AO <= 0) & (AO >= -torch.pi / 6)
    return (1 - 6 * AO / torch.pi) * m3 + (1 + 6 * AO / torch.pi) * m4


def distance_fn(R):
    """utils.py:245-249."""
    return 1 * (R <= 1) + (3 - R) / 2 * ((R > 1) & (R <= 3))


class CombatOracle:
    N_SUB = 5                # singlecombat_env.py:244
    GROUP = 2                # aircraft that share an env-level reset (singlecombat_env.py:207-238)
    REWARD_SCALE = 0.01      # singlecombat_env.py:176-177

    def __init__(self, num_envs, cfg=None, aero=None, dtype=torch.float32):
        """num_envs counts DUELS (ego, enemy pairs)."""
        self.num_envs, self.n, self.dtype = num_envs, 2 * num_envs, dtype
        self.cfg = dict(COMBAT_CFG)
        if cfg:
            self.cfg.update(cfg)
        n = self.n
        self.aero = aero if aero is not None else AeroNets(dtype=dtype)
        self.s = torch.zeros(n, 12, dtype=dtype)
        self.u = torch.zeros(n, 5, dtype=dtype)
        self.blood = 100 * torch.ones(n, dtype=dtype)
        self.step_count = torch.zeros(n, dtype=torch.int64)
        self.is_done = torch.ones(n, dtype=torch.bool)
        self.bad_done = torch.ones(n, dtype=torch.bool)
        self.exceed_time_limit = torch.ones(n, dtype=torch.bool)
        self.roll_dem = torch.zeros(n, dtype=dtype)
        self.pitch_dem = torch.zeros(n, dtype=dtype)
        self.pid = {k: _Pid(GAINS[k], self.cfg["dt"], n, dtype) for k in ("roll", "pitch", "yaw")}
        self.last_out = {k: torch.zeros(n, dtype=dtype) for k in ("roll", "pitch", "yaw")}
        self.ego = torch.arange(num_envs) * 2
        self.enm = self.ego + 1

    # -- reset ---------------------------------------------------------------------------------
    def reset_done_envs(self, draws):
        c = self.cfg
        flag = (self.is_done | self.bad_done) | self.exceed_time_limit
        m = flag.reshape(-1, self.GROUP).any(dim=1).repeat_interleave(self.GROUP)
        d = draws.to(self.dtype)
        self.s[m, :] = 0
        self.u[m, :] = 0
        self.s[m, 0] = d[m, 0] * (c["max_npos"] - c["min_npos"]) + c["min_npos"]
        self.s[m, 1] = d[m, 1] * (c["max_epos"] - c["min_epos"]) + c["min_epos"]
        self.s[m, 2] = d[m, 2] * (c["max_altitude"] - c["min_altitude"]) + c["min_altitude"]
        self.s[m, 5] = d[m, 3] * (c["max_heading"] - c["min_heading"]) + c["min_heading"]
        self.s[m, 6] = d[m, 4] * (c["max_vt"] - c["min_vt"]) + c["min_vt"]
        self.u[m, 0] = c["init_T"]
        self.blood[m] = 100
        self.step_count[m] = 0
        self.is_done[:] = False
        self.bad_done[:] = False
        self.exceed_time_limit[:] = False

    def reset(self, draws):
        self.is_done[:] = True
        self.reset_done_envs(draws)
        return self.obs()

    # -- controller (Controller.stabilize, controller.py:43-74; see planning_oracle.py for the loops) -----------
    def _rate_out(self, name, desired_rate, rate, scaler, e2t, strict):
        p = self.pid[name]
        limit = (torch.abs(self.last_out[name]) > 45) if strict else (torch.abs(self.last_out[name]) >= 45)
        p.update_all(desired_rate * scaler * scaler, rate * scaler * scaler, limit)
        ff, pp, ii, dd = p.terms()
        out = ff / (scaler * e2t + 1e-8) + pp + ii + dd
        out = 180 * out / torch.pi
        self.last_out[name] = out
        return torch.clamp(out, -45, 45)

    def stabilize(self):
        s, c = self.s, self.cfg
        gravity = 32.174
        xdot = nlplant(self.aero, s, self.u)
        roll, pitch = s[:, 3], s[:, 4]
        TAS = s[:, 6] + c["airspeed"] * torch.ones_like(s[:, 6])
        e2t = eas2tas(s[:, 2])
        scale_min, scale_max = min(0.5, 1000 / (2 * 2300)), max(2.0, 1000 / (0.7 * 100))
        scaler = torch.clamp(1000 / (TAS + 1e-8), scale_min, scale_max)
        ail = self._rate_out("roll", wrap_pi(self.roll_dem - roll) / 0.5, xdot[:, 3], scaler, e2t, False)
        desired = wrap_pi(self.pitch_dem - pitch) / 0.5
        m1 = torch.abs(roll) < (torch.pi / 2)
        m2 = roll >= (torch.pi / 2)
        m3 = roll <= (-torch.pi / 2)
        r1 = torch.clamp(roll, -4 * torch.pi / 9, 4 * torch.pi / 9)
        r2 = torch.clamp(roll, 5 * torch.pi / 9, torch.pi)
        r3 = torch.clamp(roll, -torch.pi, -5 * torch.pi / 9)
        inverted = ~m1
        rollc = m1 * r1 + m2 * r2 + m3 * r3
        mp = torch.abs(pitch) <= (7 * torch.pi / 18)
        rate_offset = mp * torch.cos(pitch) * torch.abs(gravity / TAS * torch.tan(rollc) * torch.sin(rollc) * e2t) * 1
        rate_offset = rate_offset * ~inverted - rate_offset * inverted
        desired1 = desired + rate_offset
        desired = ~inverted * desired1 + inverted * (rate_offset - desired)
        roll_wrapped = torch.abs(roll)
        mk = roll_wrapped > (torch.pi / 2)
        roll_wrapped = mk * (torch.pi - roll_wrapped) + (~mk) * roll_wrapped
        mk = (roll_wrapped > (5 * torch.pi / 18)) & (torch.abs(pitch) < (7 * torch.pi / 18))
        roll_prop = (roll_wrapped - 5 * torch.pi / 18) / (4 * torch.pi / 18)
        roll_prop = roll_prop * mk
        desired = desired * (1 - roll_prop)
        el = self._rate_out("pitch", desired, xdot[:, 4], scaler, e2t, True)
        rud = self._rate_out("yaw", torch.zeros_like(roll), xdot[:, 5], scaler, e2t, False)
        return el, ail, rud

    # -- obs / reward (singlecombat_env.py:64-181) ---------------------------------------------------------
    def _geometry_inputs(self):
        es = nlplant(self.aero, self.s, self.u)
        return self.s[self.ego, :3], self.s[self.enm, :3], es[self.ego, :3], es[self.enm, :3]

    def obs(self):
        s = self.s
        sa, ca, sb, cb = torch.sin(s[:, 7]), torch.cos(s[:, 7]), torch.sin(s[:, 8]), torch.cos(s[:, 8])
        vel = torch.stack((s[:, 6] * cb * ca, s[:, 6] * sb, s[:, 6] * cb * sa), 1)       # F16Model.get_velocity
        ego, enm = self.ego, self.enm
        inter = lambda a, b: torch.stack((a, b), 1).reshape(-1)                          # hstack(...).reshape(-1, 1)
        dvx = inter(vel[enm, 0] - vel[ego, 0], vel[ego, 0] - vel[enm, 0]) * 0.3048 / 340
        dalt = inter(s[enm, 2] - s[ego, 2], s[ego, 2] - s[enm, 2]) * 0.3048 / 1000
        AO, TA, dist, side = ao_ta_r(*self._geometry_inputs(), two_d=True)
        cols = [s[:, 2] * 0.3048 / 5000, torch.sin(s[:, 3]), torch.cos(s[:, 3]), torch.sin(s[:, 4]), torch.cos(s[:, 4]),
                vel[:, 0] * 0.3048 / 340, vel[:, 1] * 0.3048 / 340, vel[:, 2] * 0.3048 / 340, s[:, 6] * 0.3048 / 340,
                dvx, dalt, inter(AO, torch.pi - TA), inter(TA, torch.pi - AO), inter(dist, dist) * 0.3048 / 10000,
                inter(side, -side)]
        return torch.stack(cols, 1)

    def reward(self):
        AO, TA, dist, _ = ao_ta_r(*self._geometry_inputs())
        rr = range_reward_v3(self.cfg["target_dist"], dist * 0.3048 / 1000)
        ego_r = orientation_reward_v2(AO, TA) * rr
        enm_r = orientation_reward_v2(torch.pi - TA, torch.pi - AO) * rr
        return self.REWARD_SCALE * torch.stack((ego_r, enm_r), 1).reshape(-1)

    # -- terminations (the eight classes of singlecombat_env.py:48-58) ---------------------------------------
    def terminations(self):
        s, c = self.s, self.cfg
        ax, ay, az = body_accel(self.aero, s, self.u)
        bad = (torch.sqrt(ax ** 2 + ay ** 2 + az ** 2) - c["acceleration_limit"]) > 0
        bad = bad | ((s[:, 2] - c["altitude_limit"]) < 0)
        vel = (s[:, 6] + c["airspeed"] * torch.ones_like(s[:, 6])) * 0.3048 / 340
        bad = bad | ((vel - c["max_velocity"]) >= 0) | ((vel - c["min_velocity"]) <= 0)
        a_deg, b_deg = s[:, 7] * 180 / torch.pi, s[:, 8] * 180 / torch.pi
        bad = bad | ((a_deg < c["min_alpha"]) | (a_deg > c["max_alpha"])) | ((b_deg < c["min_beta"]) | (b_deg > c["max_beta"]))
        ego, enm = self.ego, self.enm
        d2 = (s[ego, 0] - s[enm, 0]) ** 2 + (s[ego, 1] - s[enm, 1]) ** 2 + (s[ego, 2] - s[enm, 2]) ** 2   # crash.py:40-42
        crash = (d2 <= c["distance_limit"] ** 2).repeat_interleave(2)
        m1, m2 = self.blood[ego] <= 0, self.blood[enm] <= 0                                              # shutdown.py:30-40
        done = (m2 & (~m1)).repeat_interleave(2)
        bad = bad | crash | m1.repeat_interleave(2)
        exc = (self.step_count - c["max_steps"]) >= 0                                                    # timeout.py:29
        return done, bad, exc

    # -- step ------------------------------------------------------------------------------------------------
    def step(self, action, draws):
        self.reset_done_envs(draws)
        a = torch.clamp(action.to(self.dtype), -1, 1)
        for _ in range(self.N_SUB):
            self.roll_dem = 0.9 * self.roll_dem + 0.1 * a[:, 1] * 4 * torch.pi / 9
            self.pitch_dem = 0.9 * self.pitch_dem + 0.1 * a[:, 2] * torch.pi / 12
            el, ail, rud = self.stabilize()
            ego_action = torch.stack((a[:, 0], -el / 45, -ail / 45, -rud / 45), 1)
            self.u = lowpass_controls(self.u, ego_action)
            self.s = euler_step(self.aero, self.s, self.u, self.cfg["dt"])
            self.step_count += 1
            done, bad, exc = self.terminations()
            self.is_done = self.is_done | done
            self.bad_done = self.bad_done | bad
            self.exceed_time_limit = self.exceed_time_limit | exc
        obs = self.obs()
        reward = self.reward()
        AO, TA, R, _ = ao_ta_r(*self._geometry_inputs())
        self.blood[self.enm] -= orientation_fn(AO) * distance_fn(R * 0.3048 / 1000)
        self.blood[self.ego] -= orientation_fn(torch.pi - TA) * distance_fn(R * 0.3048 / 1000)
        return obs, reward, self.is_done.clone(), self.bad_done.clone(), self.exceed_time_limit.clone()

    def ctrl_state(self):
        rows = [self.roll_dem, self.pitch_dem]
        for k in ("roll", "pitch", "yaw"):
            rows += [self.pid[k].error, self.pid[k].integrator, self.last_out[k]]
        return torch.stack(rows, 1)


MULTI_CFG = dict(COMBAT_CFG, max_npos=10000, min_npos=-10000, max_epos=10000, min_epos=-10000)   # multiple_selfplay.yaml


class MultiCombatOracle(CombatOracle):
    """MultipleCombatEnv (envs/multiplecombat_env.py) with the pairing restated as two adjacent duels per env -- see the
    docstring of neuralplane_b200/envs/multiplecombat_env.py for why and for what is kept from the file: per-duel 1-v-1
    formulas, env-level reset over all four agents (:207-238), ONE FDM step per env step (:258), no 0.01 reward factor
    (:176-177).  `num_envs` counts ENVS of four aircraft.  Parity UNPINNED for this orchestration."""
    N_SUB = 1
    GROUP = 4
    REWARD_SCALE = 1.0

    def __init__(self, num_envs, cfg=None, aero=None, dtype=torch.float32):
        c = dict(MULTI_CFG)
        if cfg:
            c.update(cfg)
        super().__init__(2 * num_envs, c, aero, dtype)
