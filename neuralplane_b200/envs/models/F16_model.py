"""F-16 plug-in (reference: envs/models/F16_model.py:10-182) over SoA device buffers.

`s` ([n,12]) and `u` ([n,5]) are transposed VIEWS of the field-major buffers the step kernel reads and writes, so
reference-style indexing (`model.s[:, 2]`, `model.s[mask, :] = ...`) works and hits the live state.  The dynamics
(`update`, `get_extended_state` and every getter built on it) run in libnplane.so.
"""
import torch

from .. import _soa
from ... import _native as nv
from .model_base import BaseModel


class F16Model(BaseModel):
    model_id = nv.MODEL_IDS["F16"]

    def __init__(self, config, n, device, random_seed, aero=None, ld=None):
        super().__init__(config, n, device, random_seed)
        self.num_states = getattr(self.config, 'num_states', 12)
        self.num_controls = getattr(self.config, 'num_controls', 5)
        self.dt = getattr(self.config, 'dt', 0.02)
        self.solver = getattr(self.config, 'solver', 'euler')
        self.airspeed = getattr(self.config, 'airspeed', 0)
        if self.solver != 'euler' or self.num_states != 12 or self.num_controls != 5:
            raise NotImplementedError("the native F-16 step implements solver='euler', 12 states, 5 controls")
        self.max_altitude = getattr(self.config, 'max_altitude', 20000)
        self.min_altitude = getattr(self.config, 'min_altitude', 19000)
        self.max_vt = getattr(self.config, 'max_vt', 1200)
        self.min_vt = getattr(self.config, 'min_vt', 1000)
        self.init_state = getattr(self.config, 'init_state', {'init_T': getattr(self.config, 'init_T', 2000)})

        self.aero = aero if aero is not None else self._load_aero(device)
        self.ld = ld if ld is not None else _soa.pitch(n)
        self._s = torch.zeros((12, self.ld), device=device)     # field-major state rows
        self._u = torch.zeros((5, self.ld), device=device)      # T el ail rud lef
        self._xdot = torch.zeros((17, self.ld), device=device)  # rows 12..16 stay 0 (F16_dynamics.py:60)
        self.s = self._s.t()[:n]
        self.u = self._u.t()[:n]
        # recent_s / recent_u: the state / controls the last update() started from (F16_model.py:58,63).  The fused env step
        # does not spend HBM traffic on them (nothing on the control-task path reads them); the stand-alone update() does.
        self._recent_s = torch.zeros((12, self.ld), device=device)
        self._recent_u = torch.zeros((5, self.ld), device=device)
        self.recent_s = self._recent_s.t()[:n]
        self.recent_u = self._recent_u.t()[:n]

    def _load_aero(self, device):
        from ...aero import get_aero
        return get_aero(device)

    # -- dynamics (native) ---------------------------------------------------------------------------------
    def reset(self, env):
        env.reset()

    def _action_ptr(self, action):
        if not torch.is_tensor(action):
            action = torch.as_tensor(action, dtype=torch.float32, device=self.device)
        if action.dim() != 2 or action.shape[0] != self.n or action.shape[1] < 4:
            raise ValueError(f"action must have shape [{self.n}, >=4], got {tuple(action.shape)}")
        if action.shape[1] != 4 or action.dtype != torch.float32 or not action.is_contiguous() or action.device != self._s.device:
            action = action[:, :4].to(device=self._s.device, dtype=torch.float32).contiguous()
        return action

    def update(self, action):
        """F16Model.update (F16_model.py:51-67): clamp, control low-pass, one explicit Euler step of nlplant -- one native
        launch (np_f16_update), bit-identical to what env.step() does to the same (s, u, action)."""
        a = self._action_ptr(action)
        st = nv.lib().np_f16_update(self.aero.handle, self._s.data_ptr(), self._u.data_ptr(), self._recent_s.data_ptr(),
                                    self._recent_u.data_ptr(), a.data_ptr(), self.n, self.ld, float(self.dt),
                                    torch.cuda.current_stream(self._s.device).cuda_stream)
        nv.check(st, "np_f16_update")

    def get_extended_state(self):
        """xdot of F16Dynamics.nlplant at the current (s, u): [n,17] view, columns 12..16 zero."""
        st = nv.lib().np_f16_nlplant(self.aero.handle, self._s.data_ptr(), self._u.data_ptr(), self._xdot.data_ptr(),
                                     self.n, self.ld, torch.cuda.current_stream(self.device).cuda_stream)
        nv.check(st, "np_f16_nlplant")
        return self._xdot.t()[:self.n]

    # -- getters (F16_model.py:69-182) -----------------------------------------------------------------------
    def get_state(self):
        return self.s

    def get_control(self):
        return self.u

    def get_position(self):
        return self.s[:, 0], self.s[:, 1], self.s[:, 2]

    def get_ground_speed(self):
        es = self.get_extended_state()
        return es[:, 0], es[:, 1]

    def get_climb_rate(self):
        return self.get_extended_state()[:, 2]

    def get_posture(self):
        return self.s[:, 3], self.s[:, 4], self.s[:, 5]

    def get_euler_angular_velocity(self):
        es = self.get_extended_state()
        return es[:, 3], es[:, 4], es[:, 5]

    def get_vt(self):
        return self.s[:, 6]

    def get_TAS(self):
        return self.s[:, 6] + self.airspeed * torch.ones_like(self.s[:, 6])

    def get_EAS(self):
        return self.get_TAS() / self.get_EAS2TAS()

    def get_AOA(self):
        return self.s[:, 7]

    def get_AOS(self):
        return self.s[:, 8]

    def get_angular_velocity(self):
        return self.s[:, 9], self.s[:, 10], self.s[:, 11]

    def get_thrust(self):
        return self.u[:, 0]

    def get_control_surface(self):
        return self.u[:, 1], self.u[:, 2], self.u[:, 3], self.u[:, 4]

    def _wind_axes(self):
        s = self.s
        return torch.sin(s[:, 7]), torch.cos(s[:, 7]), torch.sin(s[:, 8]), torch.cos(s[:, 8])

    def get_velocity(self):
        sina, cosa, sinb, cosb = self._wind_axes()
        vt = self.s[:, 6]
        return vt * cosb * cosa, vt * sinb, vt * cosb * sina

    def get_acceleration(self):
        xdot = self.get_extended_state()
        s = self.s
        sina, cosa, sinb, cosb = self._wind_axes()
        vel_u, vel_v, vel_w = self.get_velocity()
        u_dot = cosb * cosa * xdot[:, 6] - s[:, 6] * sinb * cosa * xdot[:, 8] - s[:, 6] * cosb * sina * xdot[:, 7]
        v_dot = sinb * xdot[:, 6] + s[:, 6] * cosb * xdot[:, 8]
        w_dot = cosb * sina * xdot[:, 6] - s[:, 6] * sinb * sina * xdot[:, 8] + s[:, 6] * cosb * cosa * xdot[:, 7]
        ax = u_dot + s[:, 10] * vel_w - s[:, 11] * vel_v
        ay = v_dot + s[:, 11] * vel_u - s[:, 9] * vel_w
        az = w_dot + s[:, 9] * vel_v - s[:, 10] * vel_u
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
        tfac = 1 - .703e-5 * self.s[:, 2]
        return torch.sqrt(1 / torch.pow(tfac, 4.14))

    def get_atmos(self):
        alt, vt = self.s[:, 2], self.s[:, 6]
        tfac = 1 - .703e-5 * alt
        temp = torch.where(alt >= 35000.0, torch.full_like(alt, 390.0), 519.0 * tfac)
        rho = 2.377e-3 * torch.pow(tfac, 4.14)
        mach = vt / torch.sqrt(1.4 * 1716.3 * temp)
        qbar = .5 * rho * vt * vt
        ps = 1715.0 * rho * temp
        ps = torch.where(ps == 0, torch.full_like(ps, 1715.0), ps)
        return mach, qbar, ps


class F16TablesModel(F16Model):
    """The same F-16 plug-in with the TABLE aero back-end (SURVEY f-3): every aero coefficient is a multilinear
    interpolation of the NASA tables (example/data/*.dat, example/train_model/hifi_F16_AeroData.py:406-483) instead of
    the MLP surrogate fitted to them.  Selected with ControlEnv(model='F16_tables')."""
    aero_backend = "tables"

    def _load_aero(self, device):
        from ...aero_tables import get_tables
        return get_tables(device)

    def update(self, action):
        a = self._action_ptr(action)
        st = nv.lib().np_f16_table_update(self.aero.handle, self._s.data_ptr(), self._u.data_ptr(), self._recent_s.data_ptr(),
                                          self._recent_u.data_ptr(), a.data_ptr(), self.n, self.ld, float(self.dt),
                                          torch.cuda.current_stream(self._s.device).cuda_stream)
        nv.check(st, "np_f16_table_update")

    def get_extended_state(self):
        st = nv.lib().np_f16_table_nlplant(self.aero.handle, self._s.data_ptr(), self._u.data_ptr(), self._xdot.data_ptr(),
                                           self.n, self.ld, torch.cuda.current_stream(self.device).cuda_stream)
        nv.check(st, "np_f16_table_nlplant")
        return self._xdot.t()[:self.n]
