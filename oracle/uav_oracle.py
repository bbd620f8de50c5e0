"""CPU oracle for ControlEnv(model='UAV').  TEST INFRASTRUCTURE ONLY (same rules as oracle/f16_oracle.py).

Restates, in torch CPU ops and the reference's operand order, `UAVModel` (envs/models/UAV_model.py:10-176),
`UAVDynamics.nlplant` (envs/models/UAV/UAV_dynamics.py:15-84) and the task layer as it sees this model THROUGH THE
GETTERS (heading_task.py:49-152, control_task.py:49-152, tracking_task.py:48-155, termination_conditions/*.py,
reward_functions/*.py).  Pinned bit-exact in float32 by tests/golden/uav_*_traj.npz, generated from the unmodified
reference with the one fix it needs to run more than one step: `num_controls = 3` (the reference's update() shrinks
`u` to the three force columns the dynamics read, while the shipped yaml says 5 -- SURVEY App. D.9).
"""
import torch

from .f16_oracle import DEFAULT_CFG, wrap_pi


def uav_nlplant(s, u):
    """UAVDynamics.nlplant (UAV_dynamics.py:15-84): xdot[:, :12] for s[n,12] (SI), forces u[n,3]."""
    UAV_M = 300
    M, N, L_bar = 1.0, 1.0, 1.0
    I_x, I_y, I_z, I_xz = 1.0, 1.0, 1.0, 0
    g = 9.81
    phi, theta, psi = s[:, 3], s[:, 4], s[:, 5]
    U, V, W = s[:, 6], s[:, 7], s[:, 8]
    P, Q, R = s[:, 9], s[:, 10], s[:, 11]
    F_x, F_y, F_z = u[:, 0], u[:, 1], u[:, 2]
    st, ct, tt = torch.sin(theta), torch.cos(theta), torch.tan(theta)
    sphi, cphi = torch.sin(phi), torch.cos(phi)
    spsi, cpsi = torch.sin(psi), torch.cos(psi)
    xd = [None] * 12
    xd[0] = U * (ct * cpsi) + V * (sphi * st * cpsi - cphi * spsi) + W * (sphi * spsi + cphi * st * cpsi)
    xd[1] = U * (ct * spsi) + V * (sphi * st * spsi + cphi * cpsi) + W * (-sphi * cpsi + cphi * st * spsi)
    xd[2] = U * st - V * (sphi * ct) - W * (cphi * ct)
    xd[3] = P + (R * cphi + Q * sphi) * tt
    xd[4] = Q * cphi - R * sphi
    xd[5] = (R * cphi + Q * sphi) / ct
    xd[6] = V * R - W * Q - g * st + F_x / UAV_M
    xd[7] = -U * R + W * P + g * ct * sphi + F_y / UAV_M
    xd[8] = U * Q - V * P + g * ct * cphi + F_z / UAV_M
    b0 = L_bar - Q * R * (I_z - I_y) + P * Q * I_xz
    b1 = N - P * Q * (I_y - I_x) - Q * R * I_xz
    b2 = M - P * R * (I_x - I_z) - (P ** 2 - R ** 2) * I_xz
    xd[9] = (b0 * I_z + b1 * I_xz) / (I_z * I_x - I_xz ** 2)
    xd[10] = b2 / I_y
    xd[11] = (b0 * I_xz + b1 * I_x) / (I_z * I_x - I_xz ** 2)
    return torch.stack(xd, dim=1)


class UAVEnvOracle:
    """ControlEnv(config=task, model='UAV') restated.  draws[n,5]: [0] altitude, [1] vt (UAV_model.py:41-42),
    [2:5] the task's target draws; noise[n,22] standard normals or None."""

    def __init__(self, n, task="control", cfg=None, dtype=torch.float32):
        assert task in ("heading", "control", "tracking")
        self.n, self.task, self.dtype = n, task, dtype
        self.cfg = dict(DEFAULT_CFG)
        if cfg:
            self.cfg.update(cfg)
        self.s = torch.zeros(n, 12, dtype=dtype)
        self.u = torch.zeros(n, 3, dtype=dtype)
        self.tgt = torch.zeros(n, 3, dtype=dtype)
        self.step_count = torch.zeros(n, dtype=torch.int64)
        self.is_done = torch.ones(n, dtype=torch.bool)
        self.bad_done = torch.ones(n, dtype=torch.bool)
        self.exceed_time_limit = torch.ones(n, dtype=torch.bool)

    # -- getters (UAV_model.py:63-134) -------------------------------------------------------
    def position(self):
        return self.s[:, 0] / 0.3048, self.s[:, 1] / 0.3048, self.s[:, 2] / 0.3048

    def vt(self):
        U, V, W = self.s[:, 6], self.s[:, 7], self.s[:, 8]
        return torch.sqrt(U ** 2 + V ** 2 + W ** 2) / 0.3048

    def tas(self):
        vt = self.vt()
        return vt + self.cfg["airspeed"] * torch.ones_like(vt)

    def eas2tas(self):
        alt = self.s[:, 2] / 0.3048
        tfac = 1 - .703e-5 * alt
        return torch.sqrt(1 / torch.pow(tfac, 4.14))

    def acceleration(self):
        xdot = uav_nlplant(self.s, self.u)
        s = self.s
        vel_u, vel_v, vel_w = s[:, 6] / 0.3048, s[:, 7] / 0.3048, s[:, 8] / 0.3048
        u_dot, v_dot, w_dot = xdot[:, 6] / 0.3048, xdot[:, 7] / 0.3048, xdot[:, 8] / 0.3048
        ax = u_dot + s[:, 10] * vel_w - s[:, 11] * vel_v
        ay = v_dot + s[:, 11] * vel_u - s[:, 9] * vel_w
        az = w_dot + s[:, 9] * vel_v - s[:, 10] * vel_u
        return ax, ay, az

    # -- reset (env_base.py:83-97, UAV_model.py:32-45, task.reset) --------------------------------
    def reset(self, draws, noise=None):
        c = self.cfg
        m = (self.is_done | self.bad_done) | self.exceed_time_limit
        d = draws.to(self.dtype)
        self.s[m, :] = 0
        self.u[m, :] = 0
        self.s[m, 2] = (d[m, 0] * (c["max_altitude"] - c["min_altitude"]) + c["min_altitude"]) * 0.3048
        self.s[m, 6] = (d[m, 1] * (c["max_vt"] - c["min_vt"]) + c["min_vt"]) * 0.3048
        self.u[m, 0] = c["init_T"]
        npos, epos, alt = self.position()
        pitch, hdg, vt = self.s[:, 4], self.s[:, 5], self.vt()
        if self.task == "heading":
            self.tgt[m, 0] = alt[m] + 1000
            self.tgt[m, 1] = wrap_pi(hdg[m] + 2 * torch.pi / 3)
            self.tgt[m, 2] = vt[m] + 0
        elif self.task == "control":
            self.tgt[m, 0] = wrap_pi(pitch[m] + 2 * (d[m, 2] - 0.5) * c["max_pitch_increment"])
            self.tgt[m, 1] = wrap_pi(hdg[m] + 2 * (d[m, 3] - 0.5) * c["max_heading_increment"])
            self.tgt[m, 2] = vt[m] + 2 * (d[m, 4] - 0.5) * c["max_velocities_u_increment"]
        else:
            dist = d[m, 2] * (c["max_distance"] - c["min_distance"]) + c["min_distance"]
            th1 = d[m, 3] * torch.pi / 3 - torch.pi / 6
            th2 = d[m, 4] * torch.pi / 3 - torch.pi / 6
            self.tgt[m, 0] = npos[m] + dist * torch.cos(th1) * torch.cos(th2)
            self.tgt[m, 1] = epos[m] + dist * torch.cos(th1) * torch.sin(th2)
            self.tgt[m, 2] = alt[m] + dist * torch.sin(th1)
        self.step_count[m] = 0
        self.is_done[:] = False
        self.bad_done[:] = False
        self.exceed_time_limit[:] = False
        return self.obs(noise)

    # -- obs (heading_task.py:71-152 through the getters) -------------------------------------------
    def obs(self, noise=None):
        s, tgt, c = self.s, self.tgt, self.cfg
        npos, epos, alt = self.position()
        roll, pitch, hdg, vt = s[:, 3], s[:, 4], s[:, 5], self.vt()
        e2t = self.eas2tas()
        EAS = self.tas() / e2t
        zero = torch.zeros_like(s[:, 0])
        if self.task == "heading":
            o0, o1, o2 = (alt - tgt[:, 0]) * 0.3048 / 1000, wrap_pi(hdg - tgt[:, 1]), (vt - tgt[:, 2]) * 0.3048 / 340
        elif self.task == "control":
            o0, o1, o2 = wrap_pi(pitch - tgt[:, 0]), wrap_pi(hdg - tgt[:, 1]), (vt - tgt[:, 2]) * 0.3048 / 340
        else:
            o0, o1, o2 = (npos - tgt[:, 0]) * 0.3048 / 1000, (epos - tgt[:, 1]) * 0.3048 / 1000, (alt - tgt[:, 2]) * 0.3048 / 1000
        cols = [o0, o1, o2, alt * 0.3048 / 5000, torch.sin(roll), torch.cos(roll), torch.sin(pitch), torch.cos(pitch),
                EAS * 0.3048 / 340, torch.sin(zero), torch.cos(zero), torch.sin(zero), torch.cos(zero),
                s[:, 9], s[:, 10], s[:, 11], zero / 0.225 / 76300 * 0.3048, zero / 45, zero / 45, zero / 45, zero / 45, e2t]
        o = torch.stack(cols, dim=1)
        if noise is not None:
            o = o + noise.to(o.dtype) * c["noise_scale"]
        return o

    # -- terminations + reward (task_base.py:60-96) ---------------------------------------------------
    def terminations(self):
        s, tgt, c = self.s, self.tgt, self.cfg
        ax, ay, az = self.acceleration()
        acc = torch.sqrt(ax ** 2 + ay ** 2 + az ** 2)
        overload = (acc - c["acceleration_limit"]) > 0
        npos, epos, alt = self.position()
        vt = self.vt()
        low_alt = (alt - c["altitude_limit"]) < 0
        vel = self.tas() * 0.3048 / 340
        hi, lo = (vel - c["max_velocity"]) >= 0, (vel - c["min_velocity"]) <= 0
        a_deg = torch.zeros_like(alt) * 180 / torch.pi
        ext = ((a_deg < c["min_alpha"]) | (a_deg > c["max_alpha"])) | ((a_deg < c["min_beta"]) | (a_deg > c["max_beta"]))
        late = self.step_count >= c["max_check_interval"]
        if self.task == "heading":
            off = ((torch.abs(wrap_pi(s[:, 5] - tgt[:, 1])) >= torch.pi / 36) | (torch.abs(alt - tgt[:, 0]) >= 100)) \
                | (torch.abs(vt - tgt[:, 2]) >= 20)
            done = ((~off) & (~late)) & (self.step_count >= c["min_check_interval"])
        elif self.task == "control":
            off = ((torch.abs(wrap_pi(s[:, 5] - tgt[:, 1])) >= torch.pi / 36) | (torch.abs(s[:, 4] - tgt[:, 0]) >= torch.pi / 36)) \
                | (torch.abs(vt - tgt[:, 2]) >= 20)
            done = (~off) & (~late)
        else:
            off = ((torch.abs(npos - tgt[:, 0]) >= 100) | (torch.abs(epos - tgt[:, 1]) >= 100)) | (torch.abs(alt - tgt[:, 2]) >= 100)
            done = (~off) & (~late)
        bad = overload | low_alt | hi | lo | ext | (late & off)
        self.last_accel = acc
        return done, bad, torch.zeros_like(bad)

    def reward(self):
        s, tgt = self.s, self.tgt
        npos, epos, alt = self.position()
        vt = self.vt()
        if self.task == "heading":
            d0, d1, d2 = (alt - tgt[:, 0]) * 0.3048 / 1000, wrap_pi(s[:, 5] - tgt[:, 1]) / torch.pi, (vt - tgt[:, 2]) * 0.3048 / 340
            r = -d0 ** 2 + -d1 ** 2 + -d2 ** 2
        elif self.task == "control":
            d0, d1, d2 = wrap_pi(s[:, 4] - tgt[:, 0]) / torch.pi, wrap_pi(s[:, 5] - tgt[:, 1]) / torch.pi, (vt - tgt[:, 2]) * 0.3048 / 340
            r = -d0 ** 2 + -d1 ** 2 + -d2 ** 2
        else:
            d0, d1, d2 = (npos - tgt[:, 0]) * 0.3048 / 1000, (epos - tgt[:, 1]) * 0.3048 / 1000, (alt - tgt[:, 2]) * 0.3048 / 1000
            r = 0.1 * (-d0 ** 2 + -d1 ** 2 + -d2 ** 2)
        total = torch.zeros(self.n, dtype=self.dtype)
        total += r
        total += -200 * self.bad_done + 200 * self.is_done
        return total

    # -- step (env_base.py:99-109; UAV_model.py:51-62) ----------------------------------------------------
    def step(self, action, draws, noise=None):
        self.reset(draws)
        a = torch.clamp(action.to(self.dtype), -1, 1)
        self.u = torch.stack([0.9 * self.u[:, j] + 0.1 * a[:, j] * 27000 for j in range(3)], dim=1)
        h = torch.tensor([0., self.cfg["dt"]], dtype=self.dtype)
        self.s = self.s + (h[1] - h[0]) * uav_nlplant(self.s, self.u)
        self.step_count += 1
        obs = self.obs(noise)
        done, bad, exc = self.terminations()
        self.is_done = self.is_done | done
        self.bad_done = self.bad_done | bad
        self.exceed_time_limit = self.exceed_time_limit | exc
        reward = self.reward()
        return obs, reward, self.is_done.clone(), self.bad_done.clone(), self.exceed_time_limit.clone()
