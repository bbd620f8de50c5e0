"""UAV plug-in (reference: envs/models/UAV_model.py:10-176) over SoA device buffers: a force-driven 6-DoF rigid body
(SI state, body velocities in slots 6..8, three body forces as controls).  `s` ([n,12]) and `u` ([n,3]) are
transposed views of the field-major buffers the step kernel works on; the dynamics run in libnplane.so
(uav_env_kernel / np_uav_nlplant).

The reference's `update()` shrinks `u` from the yaml's num_controls: 5 to 3 columns and crashes on its next reset
(SURVEY App. D.9); this model has the 3 controls the dynamics read (UAV_dynamics.py:50-52).
"""
import torch

from .. import _soa
from ... import _native as nv
from .model_base import BaseModel


class UAVModel(BaseModel):
    model_id = nv.MODEL_IDS["UAV"]

    def __init__(self, config, n, device, random_seed, ld=None):
        super().__init__(config, n, device, random_seed)
        self.num_states = getattr(self.config, 'num_states', 12)
        self.num_controls = 3
        self.dt = getattr(self.config, 'dt', 0.02)
        self.solver = getattr(self.config, 'solver', 'euler')
        self.airspeed = getattr(self.config, 'airspeed', 0)
        if self.solver != 'euler' or self.num_states != 12:
            raise NotImplementedError("the native UAV step implements solver='euler', 12 states")
        self.max_altitude = getattr(self.config, 'max_altitude', 20000)
        self.min_altitude = getattr(self.config, 'min_altitude', 19000)
        self.max_vt = getattr(self.config, 'max_vt', 1200)
        self.min_vt = getattr(self.config, 'min_vt', 1000)
        self.init_state = self.config.init_state
        self.aero = None
        self.ld = ld if ld is not None else _soa.pitch(n)
        self._s = torch.zeros((12, self.ld), device=device)
        self._u = torch.zeros((5, self.ld), device=device)      # rows 0..2 = Fx Fy Fz (the env binds a 5-row block)
        self._xdot = torch.zeros((12, self.ld), device=device)
        self.s = self._s.t()[:n]
        self.u = self._u.t()[:n, :3]
        self._recent_s = torch.zeros((12, self.ld), device=device)
        self.recent_s = self._recent_s.t()[:n]   # the state the last update() started from (UAV_model.py:56)

    def reset(self, env):
        env.reset()

    def update(self, action):
        """UAVModel.update (UAV_model.py:51-62): clamp, force low-pass, one explicit Euler step -- one native launch
        (np_uav_update), bit-identical to what env.step() does to the same (s, u, action)."""
        if not torch.is_tensor(action):
            action = torch.as_tensor(action, dtype=torch.float32, device=self.device)
        if action.dim() != 2 or action.shape[0] != self.n or action.shape[1] < 3:
            raise ValueError(f"action must have shape [{self.n}, >=3], got {tuple(action.shape)}")
        if action.shape[1] != 4 or action.dtype != torch.float32 or not action.is_contiguous() or action.device != self._s.device:
            a4 = torch.zeros((self.n, 4), dtype=torch.float32, device=self._s.device)
            a4[:, :min(4, action.shape[1])] = action[:, :4].to(device=self._s.device, dtype=torch.float32)
            action = a4
        st = nv.lib().np_uav_update(self._s.data_ptr(), self._u.data_ptr(), self._recent_s.data_ptr(), action.data_ptr(), self.n,
                                    self.ld, float(self.dt), torch.cuda.current_stream(self._s.device).cuda_stream)
        nv.check(st, "np_uav_update")

    def get_extended_state(self):
        """UAVDynamics.nlplant at the current (s, u): [n,12] view (the reference returns 15 columns, the last 3 zero)."""
        st = nv.lib().np_uav_nlplant(self._s.data_ptr(), self._u.data_ptr(), self._xdot.data_ptr(), self.n, self.ld,
                                     torch.cuda.current_stream(self.device).cuda_stream)
        nv.check(st, "np_uav_nlplant")
        return self._xdot.t()[:self.n]

    # -- getters (UAV_model.py:63-176) -------------------------------------------------------------------------
    def get_state(self):
        return self.s

    def get_control(self):
        return self.u

    def get_position(self):
        return self.s[:, 0] / 0.3048, self.s[:, 1] / 0.3048, self.s[:, 2] / 0.3048

    def get_ground_speed(self):
        es = self.get_extended_state()
        return es[:, 0] / 0.3048, es[:, 1] / 0.3048

    def get_climb_rate(self):
        return self.get_extended_state()[:, 2] / 0.3048

    def get_posture(self):
        return self.s[:, 3], self.s[:, 4], self.s[:, 5]

    def get_euler_angular_velocity(self):
        es = self.get_extended_state()
        return es[:, 3], es[:, 4], es[:, 5]

    def get_vt(self):
        U, V, W = self.s[:, 6], self.s[:, 7], self.s[:, 8]
        return torch.sqrt(U ** 2 + V ** 2 + W ** 2) / 0.3048

    def get_TAS(self):
        vt = self.get_vt()
        return vt + self.airspeed * torch.ones_like(vt)

    def get_EAS(self):
        return self.get_TAS() / self.get_EAS2TAS()

    def get_AOA(self):
        return torch.zeros_like(self.s[:, 0])

    def get_AOS(self):
        return torch.zeros_like(self.s[:, 0])

    def get_angular_velocity(self):
        return self.s[:, 9], self.s[:, 10], self.s[:, 11]

    def get_thrust(self):
        return torch.zeros_like(self.u[:, 0])

    def get_control_surface(self):
        z = torch.zeros_like(self.u[:, 0])
        return z, z.clone(), z.clone(), z.clone()

    def get_velocity(self):
        return self.s[:, 6] / 0.3048, self.s[:, 7] / 0.3048, self.s[:, 8] / 0.3048

    def get_acceleration(self):
        xdot = self.get_extended_state()
        vel_u, vel_v, vel_w = self.get_velocity()
        s = self.s
        ax = xdot[:, 6] / 0.3048 + s[:, 10] * vel_w - s[:, 11] * vel_v
        ay = xdot[:, 7] / 0.3048 + s[:, 11] * vel_u - s[:, 9] * vel_w
        az = xdot[:, 8] / 0.3048 + s[:, 9] * vel_v - s[:, 10] * vel_u
        return ax, ay, az

    def get_accels(self):
        grav = 32.174
        ax, ay, az = self.get_acceleration()
        s = self.s
        nx_cg = 1.0 / grav * ax + torch.sin(s[:, 4])
        ny_cg = 1.0 / grav * ay - torch.cos(s[:, 4]) * torch.sin(s[:, 3])
        nz_cg = -1.0 / grav * az + torch.cos(s[:, 4]) * torch.cos(s[:, 3])
        return nx_cg, ny_cg, nz_cg

    def get_G(self):
        nx_cg, ny_cg, nz_cg = self.get_accels()
        return torch.sqrt(nx_cg ** 2 + ny_cg ** 2 + nz_cg ** 2)

    def get_EAS2TAS(self):
        tfac = 1 - .703e-5 * (self.s[:, 2] / 0.3048)
        return torch.sqrt(1 / torch.pow(tfac, 4.14))

    def get_atmos(self):
        alt, vt = self.s[:, 2] / 0.3048, self.get_vt()
        tfac = 1 - .703e-5 * alt
        temp = torch.where(alt >= 35000.0, torch.full_like(alt, 390.0), 519.0 * tfac)
        rho = 2.377e-3 * torch.pow(tfac, 4.14)
        mach = vt / torch.sqrt(1.4 * 1716.3 * temp)
        qbar = .5 * rho * vt * vt
        ps = 1715.0 * rho * temp
        ps = torch.where(ps == 0, torch.full_like(ps, 1715.0), ps)
        return mach, qbar, ps
