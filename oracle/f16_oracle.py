"""CPU oracle for the NeuralPlane F-16 ControlEnv hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (neuralplane_b200/) never does.

It is a restatement, in plain torch CPU tensor ops (the numerical library the
reference itself runs on), of the reference's algorithm for
`ControlEnv.step()` / `reset()` with the F16 model and the heading / control /
tracking tasks.  Every function cites the reference file:line it follows
(paths relative to the reference repo root).  Arithmetic is written in the
reference's operand order with Python-float constants, so in float32 it is
bit-identical to the reference on the same torch build (pinned by
tests/test_oracle_golden.py against fixtures generated from the unmodified
reference by tests/golden/make_golden.py, and against the reference's own
recorded trajectory renders/result/*.npy).

Third-party arithmetic: the reference integrates with torchdiffeq==0.2.3
(requirement.txt:44, not in the tree) `odeint_adjoint(method='euler')` on the
grid t=[0, dt]; the fixed-grid Euler solver takes exactly one step
y1 = y0 + (t1 - t0) * f(t0, y0).  Restated in `euler_step` below; pinned by the
trajectory fixture, not by any reference test.

`dtype=torch.float64` evaluates the same formulas in double (weights cast up)
and serves as the "truth" line when comparing fp32 implementations.
"""
import math
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
AERO_NPZ = os.path.join(_HERE, "..", "neuralplane_b200", "data", "f16_aero.npz")

# yaml defaults of envs/configs/{heading,control,tracking}.yaml (identical where shared)
DEFAULT_CFG = dict(
    airspeed=0, noise_scale=0.01, dt=0.02,
    altitude_limit=2500.0, acceleration_limit=300.0, max_velocity=3, min_velocity=0.01,
    min_alpha=-20, max_alpha=45, min_beta=-30, max_beta=30,
    max_heading_increment=3, max_pitch_increment=3, max_velocities_u_increment=300.0,
    max_distance=2000, min_distance=2000,
    max_check_interval=2500, min_check_interval=300,
    init_T=2000, max_altitude=20000, min_altitude=19000, max_vt=1200, min_vt=1000,
)


# --------------------------------------------------------------------------------------
# aero-coefficient MLPs          envs/models/F16/hifi_F16_AeroData.py
# --------------------------------------------------------------------------------------
class AeroNets:
    """The 43 ReLU MLP surrogates (hifi_F16_AeroData.py:12-29, 41-129) from the packed data file."""

    def __init__(self, path=AERO_NPZ, dtype=torch.float32):
        d = np.load(path)
        self.names = [str(x) for x in d["names"]]
        self.desc = d["desc"]
        self.norm = d["norm"]
        self.dtype = dtype
        blob = torch.from_numpy(d["blob"].copy())
        self.layers = []
        for row in self.desc:
            n_layers, dims, off = int(row[4]), [int(x) for x in row[5:10]], int(row[10])
            ls = []
            for l in range(n_layers):
                i, o = dims[l], dims[l + 1]
                W = blob[off:off + i * o].reshape(o, i).to(dtype)
                b = blob[off + i * o: off + i * o + o].to(dtype)
                off += i * o + o
                ls.append((W, b))
            self.layers.append(ls)

    def index(self, name):
        return self.names.index(name)

    def eval_net(self, k, alpha_deg, beta_deg, el_deg):
        """One `_X(alpha, beta, el)` method (pattern hifi_F16_AeroData.py:149-166):
        z-score each used input (:32-33), hstack, Linear/ReLU stack (:25-29), y*std+mean (:36-37)."""
        row, nm = self.desc[k], self.norm[k]
        srcs = (alpha_deg, beta_deg, el_deg)
        cols = []
        for j in range(int(row[0])):
            x = srcs[int(row[1 + j])]
            cols.append(((x - float(nm[j])) / float(nm[3 + j])).reshape(-1, 1))
        h = cols[0] if len(cols) == 1 else torch.hstack(cols)
        h = h.to(self.dtype)
        ls = self.layers[k]
        for li, (W, b) in enumerate(ls):
            h = torch.nn.functional.linear(h, W, b)
            if li + 1 < len(ls):
                h = torch.relu(h)
        return h.reshape(-1) * float(nm[7]) + float(nm[6])

    def coeffs(self, alpha_deg, beta_deg, el_deg):
        """All coefficients the dynamics use, by name (group fns hifi_F16_AeroData.py:748-819)."""
        out = {}
        for k, name in enumerate(self.names):
            if int(self.desc[k][11]) == 0:
                continue  # delta_Czq_lef: evaluated at :786 but dropped at F16_dynamics.py:167-175
            out[name] = self.eval_net(k, alpha_deg, beta_deg, el_deg)
        return out


# --------------------------------------------------------------------------------------
# dynamics                        envs/models/F16/F16_dynamics.py
# --------------------------------------------------------------------------------------
def atmos_qbar(alt, vt):
    """Dynamic pressure of F16Dynamics.atmos (F16_dynamics.py:22-35); mach/ps are unused by nlplant."""
    rho0 = 2.377e-3
    tfac = 1 - .703e-5 * alt
    rho = rho0 * pow(tfac, 4.14)
    return .5 * rho * pow(vt, 2)


def nlplant(aero, s, u):
    """xdot[:, :12] of F16Dynamics.nlplant (F16_dynamics.py:37-229) for state s[n,12], control u[n,5]."""
    g, m, B, S, cbar = 32.17, 636.94, 30.0, 300.0, 11.32
    xcgr, xcg, Heng = 0.35, 0.30, 0.0
    Jy, Jxz, Jz, Jx = 55814.0, 982.0, 63100.0, 9496.0
    r2d = 180.0 / torch.pi

    alt, phi, theta, psi = s[:, 2], s[:, 3], s[:, 4], s[:, 5]
    vt = s[:, 6]
    alpha = s[:, 7] * r2d
    beta = s[:, 8] * r2d
    P, Q, R = s[:, 9], s[:, 10], s[:, 11]
    sa, ca = torch.sin(s[:, 7]), torch.cos(s[:, 7])
    sb, cb = torch.sin(s[:, 8]), torch.cos(s[:, 8])
    st, ct, tt = torch.sin(theta), torch.cos(theta), torch.tan(theta)
    sphi, cphi = torch.sin(phi), torch.cos(phi)
    spsi, cpsi = torch.sin(psi), torch.cos(psi)
    vt = (vt <= 0.01) * 0.01 + (vt > 0.01) * vt                      # :104

    T, el, ail, rud, lef = u[:, 0], u[:, 1], u[:, 2], u[:, 3], u[:, 4]
    dail = ail / 21.5
    drud = rud / 30.0
    dlef = (1 - lef / 25.0)
    qbar = atmos_qbar(alt, vt)

    U = vt * ca * cb
    V = vt * sb
    W = vt * sa * cb
    xd = [None] * 12
    xd[0] = U * (ct * cpsi) + V * (sphi * cpsi * st - cphi * spsi) + W * (cphi * st * cpsi + sphi * spsi)
    xd[1] = U * (ct * spsi) + V * (sphi * spsi * st + cphi * cpsi) + W * (cphi * st * spsi - sphi * cpsi)
    xd[2] = U * st - V * (sphi * ct) - W * (cphi * ct)
    xd[3] = P + tt * (Q * sphi + R * cphi)
    xd[4] = Q * cphi - R * sphi
    xd[5] = (Q * sphi + R * cphi) / ct

    c = aero.coeffs(alpha, beta, el)
    k2 = (cbar / (2 * vt))
    b2 = (B / (2 * vt))
    dXdQ = k2 * (c["Cxq"] + c["delta_Cxq_lef"] * dlef)
    Cx_tot = c["Cx"] + c["delta_Cx_lef"] * dlef + dXdQ * Q
    dZdQ = k2 * (c["Czq"] + c["delta_Cz_lef"] * dlef)                # sic: delta_Cz_lef (:199)
    Cz_tot = c["Cz"] + c["delta_Cz_lef"] * dlef + dZdQ * Q
    dMdQ = k2 * (c["Cmq"] + c["delta_Cmq_lef"] * dlef)
    delta_Cm_ds = torch.zeros_like(alpha)                             # hifi_other_coeffs :811-818
    Cm_tot = c["Cm"] * c["eta_el"] + Cz_tot * (xcgr - xcg) + c["delta_Cm_lef"] * dlef + dMdQ * Q \
        + c["delta_Cm"] + delta_Cm_ds
    dYdail = c["delta_Cy_a20"] + c["delta_Cy_a20_lef"] * dlef
    dYdR = b2 * (c["Cyr"] + c["delta_Cyr_lef"] * dlef)
    dYdP = b2 * (c["Cyp"] + c["delta_Cyp_lef"] * dlef)
    Cy_tot = c["Cy"] + c["delta_Cy_lef"] * dlef + dYdail * dail + c["delta_Cy_r30"] * drud + dYdR * R + dYdP * P
    dNdail = c["delta_Cn_a20"] + c["delta_Cn_a20_lef"] * dlef
    dNdR = b2 * (c["Cnr"] + c["delta_Cnr_lef"] * dlef)
    dNdP = b2 * (c["Cnp"] + c["delta_Cnp_lef"] * dlef)
    Cn_tot = c["Cn"] + c["delta_Cn_lef"] * dlef - Cy_tot * (xcgr - xcg) * (cbar / B) + dNdail * dail \
        + c["delta_Cn_r30"] * drud + dNdR * R + dNdP * P + c["delta_Cnbeta"] * beta
    dLdail = c["delta_Cl_a20"] + c["delta_Cl_a20_lef"] * dlef
    dLdR = b2 * (c["Clr"] + c["delta_Clr_lef"] * dlef)
    dLdP = b2 * (c["Clp"] + c["delta_Clp_lef"] * dlef)
    Cl_tot = c["Cl"] + c["delta_Cl_lef"] * dlef + dLdail * dail + c["delta_Cl_r30"] * drud + dLdR * R + dLdP * P \
        + c["delta_Clbeta"] * beta
    Udot = R * V - Q * W - g * st + qbar * S * Cx_tot / m + T / m
    Vdot = P * W - R * U + g * ct * sphi + qbar * S * Cy_tot / m
    Wdot = Q * U - P * V + g * ct * cphi + qbar * S * Cz_tot / m
    xd[6] = (U * Udot + V * Vdot + W * Wdot) / vt
    xd[7] = (U * Wdot - W * Udot) / (U * U + W * W)
    xd[8] = (Vdot * vt - V * xd[6]) / (vt * vt * cb)
    L_tot = Cl_tot * qbar * S * B
    M_tot = Cm_tot * qbar * S * cbar
    N_tot = Cn_tot * qbar * S * B
    denom = Jx * Jz - Jxz * Jxz
    xd[9] = (Jz * L_tot + Jxz * N_tot - (Jz * (Jz - Jy) + Jxz * Jxz) * Q * R + Jxz * (Jx - Jy + Jz) * P * Q
             + Jxz * Q * Heng) / denom
    xd[10] = (M_tot + (Jz - Jx) * P * R - Jxz * (P * P - R * R) - R * Heng) / Jy
    xd[11] = (Jx * N_tot + Jxz * L_tot + (Jx * (Jx - Jy) + Jxz * Jxz) * P * Q - Jxz * (Jx - Jy + Jz) * Q * R
              + Jx * Q * Heng) / denom
    return torch.stack(xd, dim=1)


def euler_step(aero, s, u, dt):
    """torchdiffeq 0.2.3 fixed-grid Euler on t=[0,dt] as called at F16_model.py:64-67 (dt is an f32 tensor)."""
    h = torch.tensor([0., dt], dtype=s.dtype)
    return s + (h[1] - h[0]) * nlplant(aero, s, u)


def lowpass_controls(u, action):
    """F16Model.update control lag (F16_model.py:52-62); lef column forced to zero."""
    a = torch.clamp(action, -1, 1)
    T = 0.9 * u[:, 0] + 0.1 * a[:, 0] * 0.225 * 76300 / 0.3048
    el = 0.9 * u[:, 1] + 0.1 * a[:, 1] * 45
    ail = 0.9 * u[:, 2] + 0.1 * a[:, 2] * 45
    rud = 0.9 * u[:, 3] + 0.1 * a[:, 3] * 45
    return torch.stack((T, el, ail, rud, torch.zeros_like(T)), dim=1)


def eas2tas(alt):
    """F16Model.get_EAS2TAS (F16_model.py:156-162)."""
    tfac = 1 - .703e-5 * alt
    return torch.sqrt(1 / torch.pow(tfac, 4.14))


def body_accel(aero, s, u):
    """F16Model.get_acceleration (F16_model.py:132-148): needs a full nlplant(s,u)."""
    xdot = nlplant(aero, s, u)
    sina, cosa = torch.sin(s[:, 7]), torch.cos(s[:, 7])
    sinb, cosb = torch.sin(s[:, 8]), torch.cos(s[:, 8])
    vel_u = s[:, 6] * cosb * cosa
    vel_v = s[:, 6] * sinb
    vel_w = s[:, 6] * cosb * sina
    u_dot = cosb * cosa * xdot[:, 6] - s[:, 6] * sinb * cosa * xdot[:, 8] - s[:, 6] * cosb * sina * xdot[:, 7]
    v_dot = sinb * xdot[:, 6] + s[:, 6] * cosb * xdot[:, 8]
    w_dot = cosb * sina * xdot[:, 6] - s[:, 6] * sinb * sina * xdot[:, 8] + s[:, 6] * cosb * cosa * xdot[:, 7]
    ax = u_dot + s[:, 10] * vel_w - s[:, 11] * vel_v
    ay = v_dot + s[:, 11] * vel_u - s[:, 9] * vel_w
    az = w_dot + s[:, 9] * vel_v - s[:, 10] * vel_u
    return ax, ay, az


def load_factors(aero, s, u):
    """F16Model.get_accels / get_G (F16_model.py:150-182), grav = 32.174 (sic, vs 32.17 in the EoM)."""
    grav = 32.174
    ax, ay, az = body_accel(aero, s, u)
    nx = 1.0 / grav * ax + torch.sin(s[:, 4])
    ny = 1.0 / grav * ay - torch.cos(s[:, 4]) * torch.sin(s[:, 3])
    nz = -1.0 / grav * az + torch.cos(s[:, 4]) * torch.cos(s[:, 3])
    return nx, ny, nz


def wrap_2pi(angle):
    """envs/utils/utils.py:144-148 (torch `%` is Python-style: result takes the divisor's sign)."""
    res = angle % (2 * torch.pi)
    res = res + 2 * torch.pi * (res < 0)
    return res


def wrap_pi(angle):
    """envs/utils/utils.py:150-154."""
    res = wrap_2pi(angle)
    res = res - 2 * torch.pi * (res > torch.pi)
    return res


# --------------------------------------------------------------------------------------
# env = BaseEnv + F16Model + task     envs/env_base.py, envs/models/F16_model.py, envs/tasks/*
# --------------------------------------------------------------------------------------
class F16EnvOracle:
    """ControlEnv(config=task, model='F16') restated (env_base.py:15-109, control_env.py:19-35).

    Reset randomness is explicit: `draws[n, 5]` uniforms in [0,1); a lane that resets uses
    draws[i,0] for altitude, draws[i,1] for vt (F16_model.py:41-42) and draws[i,2:5] for the
    task's target draws (control_task.py:59-61, tracking_task.py:57-60).  Observation noise
    (heading_task.py:152) is `noise[n,22]` standard normals scaled by noise_scale, or omitted.
    """

    def __init__(self, n, task="heading", cfg=None, aero=None, dtype=torch.float32):
        assert task in ("heading", "control", "tracking")
        self.n, self.task, self.dtype = n, task, dtype
        self.cfg = dict(DEFAULT_CFG)
        if cfg:
            self.cfg.update(cfg)
        self.aero = aero if aero is not None else AeroNets(dtype=dtype)
        self.s = torch.zeros(n, 12, dtype=dtype)
        self.u = torch.zeros(n, 5, dtype=dtype)
        self.tgt = torch.zeros(n, 3, dtype=dtype)      # heading: alt,psi,vt | control: theta,psi,vt | tracking: n,e,alt
        self.step_count = torch.zeros(n, dtype=torch.int64)
        self.is_done = torch.ones(n, dtype=torch.bool)  # env_base.py:31-33: everyone resets first
        self.bad_done = torch.ones(n, dtype=torch.bool)
        self.exceed_time_limit = torch.ones(n, dtype=torch.bool)

    # -- reset ---------------------------------------------------------------------------
    def reset(self, draws, noise=None):
        """BaseEnv.reset (env_base.py:83-97) -> F16Model.reset (F16_model.py:33-45) -> task.reset."""
        c = self.cfg
        m = (self.is_done | self.bad_done) | self.exceed_time_limit
        d = draws.to(self.dtype)
        self.s[m, :] = 0
        self.u[m, :] = 0
        self.s[m, 2] = d[m, 0] * (c["max_altitude"] - c["min_altitude"]) + c["min_altitude"]
        self.s[m, 6] = d[m, 1] * (c["max_vt"] - c["min_vt"]) + c["min_vt"]
        self.u[m, 0] = c["init_T"]
        s = self.s
        if self.task == "heading":                      # heading_task.py:49-69 (constant increments)
            self.tgt[m, 0] = s[m, 2] + 1000
            self.tgt[m, 1] = wrap_pi(s[m, 5] + 2 * torch.pi / 3)
            self.tgt[m, 2] = s[m, 6] + 0
        elif self.task == "control":                    # control_task.py:49-68
            self.tgt[m, 0] = wrap_pi(s[m, 4] + 2 * (d[m, 2] - 0.5) * c["max_pitch_increment"])
            self.tgt[m, 1] = wrap_pi(s[m, 5] + 2 * (d[m, 3] - 0.5) * c["max_heading_increment"])
            self.tgt[m, 2] = s[m, 6] + 2 * (d[m, 4] - 0.5) * c["max_velocities_u_increment"]
        else:                                           # tracking_task.py:48-71
            dist = d[m, 2] * (c["max_distance"] - c["min_distance"]) + c["min_distance"]
            th1 = d[m, 3] * torch.pi / 3 - torch.pi / 6
            th2 = d[m, 4] * torch.pi / 3 - torch.pi / 6
            self.tgt[m, 0] = s[m, 0] + dist * torch.cos(th1) * torch.cos(th2)
            self.tgt[m, 1] = s[m, 1] + dist * torch.cos(th1) * torch.sin(th2)
            self.tgt[m, 2] = s[m, 2] + dist * torch.sin(th1)
        self.step_count[m] = 0
        self.is_done[:] = False
        self.bad_done[:] = False
        self.exceed_time_limit[:] = False
        return self.obs(noise)

    # -- observation ---------------------------------------------------------------------
    def obs(self, noise=None):
        """HeadingTask.get_obs (heading_task.py:71-152); control_task.py:70-152; tracking_task.py:73-155."""
        s, u, tgt, c = self.s, self.u, self.tgt, self.cfg
        alt, roll, pitch, hdg, vt = s[:, 2], s[:, 3], s[:, 4], s[:, 5], s[:, 6]
        e2t = eas2tas(alt)
        EAS = (vt + c["airspeed"] * torch.ones_like(vt)) / e2t          # F16_model.py:96-103
        if self.task == "heading":
            o0 = (alt - tgt[:, 0]) * 0.3048 / 1000
            o1 = wrap_pi(hdg - tgt[:, 1])
            o2 = (vt - tgt[:, 2]) * 0.3048 / 340
        elif self.task == "control":
            o0 = wrap_pi(pitch - tgt[:, 0])
            o1 = wrap_pi(hdg - tgt[:, 1])
            o2 = (vt - tgt[:, 2]) * 0.3048 / 340
        else:
            o0 = (s[:, 0] - tgt[:, 0]) * 0.3048 / 1000
            o1 = (s[:, 1] - tgt[:, 1]) * 0.3048 / 1000
            o2 = (alt - tgt[:, 2]) * 0.3048 / 1000
        cols = [o0, o1, o2, alt * 0.3048 / 5000, torch.sin(roll), torch.cos(roll), torch.sin(pitch),
                torch.cos(pitch), EAS * 0.3048 / 340, torch.sin(s[:, 7]), torch.cos(s[:, 7]),
                torch.sin(s[:, 8]), torch.cos(s[:, 8]), s[:, 9], s[:, 10], s[:, 11],
                u[:, 0] / 0.225 / 76300 * 0.3048, u[:, 1] / 45, u[:, 2] / 45, u[:, 3] / 45, u[:, 4] / 45, e2t]
        o = torch.stack(cols, dim=1)
        if noise is not None:
            o = o + noise.to(o.dtype) * c["noise_scale"]
        return o

    # -- terminations ----------------------------------------------------------------------
    def terminations(self):
        """BaseTask.get_termination (task_base.py:75-96) over the six conditions in list order
        (heading_task.py:39-47): overload.py:37-42, low_altitude.py:29-30, high_speed.py:29-30,
        low_speed.py:29-30, extreme_state.py:32-36, unreach_{heading,posture,target}.py."""
        s, u, tgt, c = self.s, self.u, self.tgt, self.cfg
        ax, ay, az = body_accel(self.aero, s, u)
        acc = torch.sqrt(ax ** 2 + ay ** 2 + az ** 2)
        bad = (acc - c["acceleration_limit"]) > 0
        causes = {"overload": bad.clone()}
        low_alt = (s[:, 2] - c["altitude_limit"]) < 0
        vel = (s[:, 6] + c["airspeed"] * torch.ones_like(s[:, 6])) * 0.3048 / 340
        hi = (vel - c["max_velocity"]) >= 0
        lo = (vel - c["min_velocity"]) <= 0
        a_deg = s[:, 7] * 180 / torch.pi
        b_deg = s[:, 8] * 180 / torch.pi
        ext = ((a_deg < c["min_alpha"]) | (a_deg > c["max_alpha"])) | ((b_deg < c["min_beta"]) | (b_deg > c["max_beta"]))
        late = self.step_count >= c["max_check_interval"]
        if self.task == "heading":                       # unreach_heading.py:33-53
            early_ok = self.step_count >= c["min_check_interval"]
            off = ((torch.abs(wrap_pi(s[:, 5] - tgt[:, 1])) >= torch.pi / 36)
                   | (torch.abs(s[:, 2] - tgt[:, 0]) >= 100)) | (torch.abs(s[:, 6] - tgt[:, 2]) >= 20)
            done = ((~off) & (~late)) & early_ok
        elif self.task == "control":                     # unreach_posture.py:33-55 (pitch error not wrapped)
            off = ((torch.abs(wrap_pi(s[:, 5] - tgt[:, 1])) >= torch.pi / 36)
                   | (torch.abs(s[:, 4] - tgt[:, 0]) >= torch.pi / 36)) | (torch.abs(s[:, 6] - tgt[:, 2]) >= 20)
            done = (~off) & (~late)
        else:                                            # unreach_target.py:31-47
            off = ((torch.abs(s[:, 0] - tgt[:, 0]) >= 100) | (torch.abs(s[:, 1] - tgt[:, 1]) >= 100)) \
                | (torch.abs(s[:, 2] - tgt[:, 2]) >= 100)
            done = (~off) & (~late)
        unreach = late & off
        causes.update(low_altitude=low_alt, high_speed=hi, low_speed=lo, extreme_state=ext, unreach=unreach, reached=done)
        bad = bad | low_alt | hi | lo | ext | unreach
        self.last_causes = causes
        self.last_accel = acc
        return done, bad, torch.zeros_like(bad)

    # -- reward ----------------------------------------------------------------------------
    def reward(self):
        """task_base.py:60-73; heading_reward.py:17-36 / posture_reward.py:17-35 / position_reward.py:17-34;
        event_driven_reward.py:28 (+-200 on the ACCUMULATED env flags)."""
        s, tgt = self.s, self.tgt
        if self.task == "heading":
            d0 = (s[:, 2] - tgt[:, 0]) * 0.3048 / 1000
            d1 = wrap_pi(s[:, 5] - tgt[:, 1]) / torch.pi
            d2 = (s[:, 6] - tgt[:, 2]) * 0.3048 / 340
            r = -d0 ** 2 + -d1 ** 2 + -d2 ** 2
        elif self.task == "control":
            d0 = wrap_pi(s[:, 4] - tgt[:, 0]) / torch.pi
            d1 = wrap_pi(s[:, 5] - tgt[:, 1]) / torch.pi
            d2 = (s[:, 6] - tgt[:, 2]) * 0.3048 / 340
            r = -d0 ** 2 + -d1 ** 2 + -d2 ** 2
        else:
            d0 = (s[:, 0] - tgt[:, 0]) * 0.3048 / 1000
            d1 = (s[:, 1] - tgt[:, 1]) * 0.3048 / 1000
            d2 = (s[:, 2] - tgt[:, 2]) * 0.3048 / 1000
            r = 0.1 * (-d0 ** 2 + -d1 ** 2 + -d2 ** 2)
        total = torch.zeros(self.n, dtype=self.dtype)
        total += r
        total += -200 * self.bad_done + 200 * self.is_done
        return total

    # -- step ------------------------------------------------------------------------------
    def step(self, action, draws, noise=None):
        """BaseEnv.step (env_base.py:99-109)."""
        self.reset(draws)                                   # :100 (its obs is discarded)
        self.u = lowpass_controls(self.u, action.to(self.dtype))     # F16_model.py:51-63
        self.s = euler_step(self.aero, self.s, self.u, self.cfg["dt"])  # :64-67
        self.step_count += 1                                # :102
        obs = self.obs(noise)                               # :103
        done, bad, exc = self.terminations()                # :105 -> env_base.py:70-75
        self.is_done = self.is_done | done
        self.bad_done = self.bad_done | bad
        self.exceed_time_limit = self.exceed_time_limit | exc
        reward = self.reward()                              # :106
        return obs, reward, self.is_done.clone(), self.bad_done.clone(), self.exceed_time_limit.clone()
