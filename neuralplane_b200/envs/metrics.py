"""Flight-quality metrics of a recorded rollout (reference: renders/evaluate_result.py:29-43) and a recorder that
samples a few aircraft from the device state each step (SURVEY f-4).  Host-side, outside the hot path: the recorder
moves `max_aircraft` rows per step; the load factor G of every recorded sample is evaluated at the end by ONE native
nlplant call over the stacked samples (F16Model.get_G, F16_model.py:150-154)."""
import numpy as np
import torch

METRIC_NAMES = ("G", "TAS", "RoC", "AOA", "ASM", "SSM", "OSM", "AOASM", "AOSSM")


def flight_metrics(G, vt, pitch, alpha, beta, altitude):
    """The nine scalars of evaluate_result.py:29-43 from recorded series (any shape; ft, ft/s, rad, g):
    manoeuvrability  G (mean |G| / G_limit), TAS (mean Mach), RoC (mean |climb rate| / 100 m/s), AOA (mean |alpha| / 32.5 deg);
    safety margins   ASM altitude, SSM speed, OSM overload, AOASM angle of attack, AOSSM sideslip."""
    G, vt, pitch, alpha, beta, altitude = (np.asarray(x, dtype=np.float64) for x in (G, vt, pitch, alpha, beta, altitude))
    g_lim = 300 / 32.17
    mach = vt * 0.3048 / 340
    return {
        "G": float(np.mean(np.abs(G)) / g_lim),
        "TAS": float(np.mean(vt) * 0.3048 / 340),
        "RoC": float(np.mean(np.abs(vt * np.sin(pitch))) * 0.3048 / 100),
        "AOA": float(np.mean(np.abs(alpha)) * 180 / np.pi / 32.5),
        "ASM": float(np.mean(altitude - 2500) * 0.3048 / 5000),
        "SSM": float(np.mean(1.505 - np.abs(mach - 1.505)) / 1.505),
        "OSM": float(np.mean(g_lim - np.abs(G)) / g_lim),
        "AOASM": float(np.mean(32.5 - np.abs(alpha * 180 / np.pi - 12.5)) / 32.5),
        "AOSSM": float(np.mean(30 - np.abs(beta) * 180 / np.pi) / 30),
    }


class FlightRecorder:
    """Samples the first `max_aircraft` aircraft of an env after every step (what renders/render_ppo.py:98-102,153-186
    does for its single aircraft) and reduces the record to flight_metrics()."""

    def __init__(self, env, max_aircraft=16):
        self.env = env
        self.m = min(int(max_aircraft), env.n)
        self._s, self._u = [], []

    def record(self):
        self._s.append(self.env.model.s[:self.m].clone())
        self._u.append(self.env.model.u[:self.m].clone())

    def __len__(self):
        return len(self._s)

    def series(self):
        """Recorded series as numpy arrays [steps, max_aircraft]; G from one native nlplant call over all samples."""
        if not self._s:
            raise RuntimeError("nothing recorded")
        S, U = torch.cat(self._s, 0), torch.cat(self._u, 0)
        model = type(self.env.model)(self.env.config, S.shape[0], self.env.device, None)
        model.s[:] = S
        model.u[:, :U.shape[1]] = U
        G = model.get_G()
        shape = (len(self._s), self.m)
        out = {"G": G, "vt": S[:, 6], "pitch": S[:, 4], "alpha": S[:, 7], "beta": S[:, 8], "altitude": S[:, 2],
               "npos": S[:, 0], "epos": S[:, 1], "roll": S[:, 3], "yaw": S[:, 5]}
        return {k: v.cpu().numpy().reshape(shape) for k, v in out.items()}

    def metrics(self):
        r = self.series()
        return flight_metrics(r["G"], r["vt"], r["pitch"], r["alpha"], r["beta"], r["altitude"])
