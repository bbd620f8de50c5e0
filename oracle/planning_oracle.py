"""CPU oracle for the PlanningEnv-shaped step with the PID low-level controller.  TEST INFRASTRUCTURE ONLY.

Restates, in torch CPU ops and the reference's operand order:
  * PlanningEnv.step (envs/planning_env.py:144-177): reset, clamp, targets, 50 x {controller -> F16Model.update ->
    freeze already-terminated aircraft at recent_s -> step_count -> obs / done / reward}, returning the last sub-step;
  * the low-level controller that replaces the reference's GRU PPO actor (its checkpoint is not in the repository,
    planning_env.py:16): Controller.stabilize (algorithms/pid/controller.py:43-74) = RollController /
    PitchController / YawController rate loops (rollController.py:26-49, pitchController.py:30-94,
    yawController.py:69-84) on the PID class (pid.py:17-41), L1 heading hold for the heading target
    (L1Controller.py:230-271, controller.py:114-124), the pitch target fed to the pitch loop, and a TAS loop built
    from the same PID class with the gains of algorithms/pid/config/speedcontroller.yaml.
Pinned by tests/golden/planning_pid_traj.npz, which tests/golden/make_golden.py generates by running exactly those
reference classes (RefPidPlanner).  Not restated: the PID class returns zeros for EVERY aircraft when any target or
measurement is NaN/Inf (pid.py:18-21), a population-wide side effect with no per-aircraft meaning.
"""
import torch

from .f16_oracle import F16EnvOracle, eas2tas, euler_step, lowpass_controls, nlplant, wrap_pi

GAINS = {  # algorithms/pid/config/{roll,pitch,yaw,speed}controller.yaml: Kp Ki Kd Kff Kimax
    "roll": (10, 0.3, 0, 0.3, 0.666), "pitch": (10, 0.3, 0, 0.3, 0.666), "yaw": (1, 0.3, 0.05, 0.3, 0.666),
    "speed": (5, 25, 0, 80, 100),
}


class _Pid:
    """pid.py:5-56 without the population-wide NaN guard."""

    def __init__(self, gains, dt, n, dtype):
        self.Kp, self.Ki, self.Kd, self.Kff, self.Kimax = gains
        self.dt, self.first = dt, True
        self.error = torch.zeros(n, dtype=dtype)
        self.integrator = torch.zeros(n, dtype=dtype)
        self.derivative = torch.zeros(n, dtype=dtype)
        self.target = torch.zeros(n, dtype=dtype)

    def update_all(self, target, measurement, limit):
        if self.first:
            self.first = False
            self.target = target
            self.error = target - measurement
            self.derivative = torch.zeros_like(self.error)
            self.integrator = torch.zeros_like(self.error)
        else:
            last_error = self.error
            self.target = target
            self.error = target - measurement
            self.derivative = (self.error - last_error) / self.dt
        if self.Ki != 0 and self.dt > 0:                                           # update_i
            self.integrator = self.integrator + self.error * self.Ki * self.dt * (~limit | (self.error * self.dt < 0))
            self.integrator = torch.clamp(self.integrator, -self.Kimax, self.Kimax)
        else:
            self.integrator = torch.zeros_like(self.error)

    def terms(self):
        return self.target * self.Kff, self.error * self.Kp, self.integrator, self.derivative * self.Kd


class PlanningOracle(F16EnvOracle):
    N_SUB = 50

    def __init__(self, n, cfg=None, aero=None, dtype=torch.float32):
        c = dict(noise_scale=0.0)
        if cfg:
            c.update(cfg)
        super().__init__(n, "tracking", c, aero, dtype)
        dt = self.cfg["dt"]
        self.pid = {k: _Pid(GAINS[k], dt, n, dtype) for k in ("roll", "pitch", "yaw", "speed")}
        self.last_out = {k: torch.zeros(n, dtype=dtype) for k in ("roll", "pitch", "yaw", "speed")}

    # -- rate loop shared by the three attitude controllers (e.g. rollController.py:26-41) -------------
    def _rate_out(self, name, desired_rate, rate, scaler, e2t, strict):
        p = self.pid[name]
        limit = (torch.abs(self.last_out[name]) > 45) if strict else (torch.abs(self.last_out[name]) >= 45)
        p.update_all(desired_rate * scaler * scaler, rate * scaler * scaler, limit)
        ff, pp, ii, dd = p.terms()
        out = ff / (scaler * e2t + 1e-8) + pp + ii + dd
        out = 180 * out / torch.pi
        self.last_out[name] = out
        return torch.clamp(out, -45, 45)

    def controller(self, target_pitch, target_heading, target_vt):
        """One low-level control decision -> action[n,4] (controller.py:140-148)."""
        s, c = self.s, self.cfg
        gravity = 32.174
        xdot = nlplant(self.aero, s, self.u)          # get_ground_speed / get_euler_angular_velocity (F16_model.py:75-91)
        roll, pitch, yaw = s[:, 3], s[:, 4], s[:, 5]
        TAS = s[:, 6] + c["airspeed"] * torch.ones_like(s[:, 6])
        e2t = eas2tas(s[:, 2])
        # L1 heading hold (L1Controller.py:230-252, :267-271; controller.py:114-124)
        omegaA = 4.4428 / 17
        target_bearing = wrap_pi(target_heading)
        Nu = wrap_pi(target_bearing - wrap_pi(yaw))
        groundSpeed = torch.sqrt(xdot[:, 0] * xdot[:, 0] + xdot[:, 1] * xdot[:, 1])
        VomegaA = groundSpeed * omegaA
        Nu = torch.clamp(Nu, -torch.pi / 2, torch.pi / 2)
        latAccDem = 2 * torch.sin(Nu) * VomegaA
        roll_dem = torch.cos(pitch) * torch.atan(latAccDem / gravity)
        roll_dem = torch.clamp(roll_dem, -torch.pi / 2, torch.pi / 2)
        roll_dem = torch.clamp(roll_dem, -torch.pi / 4, torch.pi / 4)
        yaw_rate_dem = gravity * torch.tan(roll_dem) / TAS * e2t
        # TAS loop (PID class, speedcontroller.yaml gains)
        p = self.pid["speed"]
        p.update_all(target_vt * 0.3048 / 340, TAS * 0.3048 / 340, torch.abs(self.last_out["speed"]) >= 100)
        ff, pp, ii, dd = p.terms()
        out = ff + pp + ii + dd
        self.last_out["speed"] = out
        throttle = torch.clamp(out / 100, 0, 1)
        # Controller.stabilize (controller.py:43-74)
        scale_min, scale_max = min(0.5, 1000 / (2 * 2300)), max(2.0, 1000 / (0.7 * 100))
        scaler = torch.clamp(1000 / (TAS + 1e-8), scale_min, scale_max)
        ail = self._rate_out("roll", wrap_pi(roll_dem - roll) / 0.5, xdot[:, 3], scaler, e2t, False)
        # pitch loop with turn coordination (pitchController.py:47-94)
        desired = wrap_pi(target_pitch - pitch) / 0.5
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
        pitch_wrapped = torch.abs(pitch)
        mk = roll_wrapped > (torch.pi / 2)
        roll_wrapped = mk * (torch.pi - roll_wrapped) + (~mk) * roll_wrapped
        mk = (roll_wrapped > (5 * torch.pi / 18)) & (pitch_wrapped < (7 * torch.pi / 18))
        roll_prop = (roll_wrapped - 5 * torch.pi / 18) / (4 * torch.pi / 18)
        roll_prop = roll_prop * mk
        desired = desired * (1 - roll_prop)
        el = self._rate_out("pitch", desired, xdot[:, 4], scaler, e2t, True)
        rud = self._rate_out("yaw", yaw_rate_dem, xdot[:, 5], scaler, e2t, False)
        return torch.stack((throttle, -el / 45, -ail / 45, -rud / 45), dim=1)

    def pid_state(self):
        rows = []
        for k in ("roll", "pitch", "yaw", "speed"):
            rows += [self.pid[k].error, self.pid[k].integrator, self.last_out[k]]
        return torch.stack(rows, 1)

    def plan_step(self, action, draws, noise=None):
        """PlanningEnv.step (planning_env.py:144-177)."""
        self.reset(draws)
        a = torch.clamp(action.to(self.dtype), -1, 1)
        target_pitch = self.s[:, 4] + a[:, 0] * 0.3
        target_heading = self.s[:, 5] + a[:, 1] * 0.3
        target_vt = self.s[:, 6] + a[:, 2] * 30
        self.targets = torch.stack((target_pitch, target_heading, target_vt), 1)
        for _ in range(self.N_SUB):
            ego = self.controller(target_pitch, target_heading, target_vt)
            recent_s = self.s
            self.u = lowpass_controls(self.u, ego)
            self.s = euler_step(self.aero, self.s, self.u, self.cfg["dt"])
            frozen = (self.is_done | self.bad_done) | self.exceed_time_limit
            self.s[frozen] = recent_s[frozen]
            self.step_count += 1
            obs = self.obs(noise)
            done, bad, exc = self.terminations()
            self.is_done = self.is_done | done
            self.bad_done = self.bad_done | bad
            self.exceed_time_limit = self.exceed_time_limit | exc
            reward = self.reward()
        return obs, reward, self.is_done.clone(), self.bad_done.clone(), self.exceed_time_limit.clone()
