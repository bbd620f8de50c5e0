"""Tacview .acmi track writer (reference: BaseEnv.render, envs/env_base.py:111-151; geodetic helpers
envs/utils/utils.py:34-142).  Host-side file I/O outside the hot path: one device->host copy of the first
`max_aircraft` aircraft's position / attitude per rendered step.

Differences from the reference: every aircraft is written with ITS OWN attitude (the reference indexes `[0]`, so all
objects share aircraft 0's roll / pitch / yaw, env_base.py:138-140) and only a bounded sample of the population is
written (the reference loops over all n aircraft in Python).
"""
import os

import numpy as np

_A = 6378137.0            # WGS-84 semi-major axis, m
_B = 6356752.3142         # semi-minor axis, m (the reference's constant, utils.py:101)
_E2 = 1.0 - (_B / _A) ** 2


def enu_to_geodetic(east, north, up, lat0=0.0, lon0=0.0, h0=0.0):
    """Local east / north / up metres about (lat0, lon0, h0) degrees -> (lat deg, lon deg, height m), vectorised.
    ENU -> ECEF rotation, then Heikkinen's closed-form ECEF -> geodetic."""
    east, north, up = (np.asarray(v, dtype=np.float64) for v in (east, north, up))
    lam, phi = np.radians(lat0), np.radians(lon0)
    sl, cl, sp, cp = np.sin(lam), np.cos(lam), np.sin(phi), np.cos(phi)
    N0 = _A / np.sqrt(1.0 - _E2 * sl * sl)
    x0, y0, z0 = (h0 + N0) * cl * cp, (h0 + N0) * cl * sp, (h0 + (1.0 - _E2) * N0) * sl
    t = cl * up - sl * north
    x = cp * t - sp * east + x0
    y = sp * t + cp * east + y0
    z = sl * up + cl * north + z0
    # Heikkinen (1982)
    r = np.hypot(x, y)
    ep2 = (_A * _A - _B * _B) / (_B * _B)
    F = 54.0 * _B * _B * z * z
    G = r * r + (1.0 - _E2) * z * z - _E2 * (_A * _A - _B * _B)
    c = _E2 * _E2 * F * r * r / (G * G * G)
    s = np.cbrt(1.0 + c + np.sqrt(c * c + 2.0 * c))
    P = F / (3.0 * (s + 1.0 / s + 1.0) ** 2 * G * G)
    Q = np.sqrt(1.0 + 2.0 * _E2 * _E2 * P)
    r0 = -(P * _E2 * r) / (1.0 + Q) + np.sqrt(0.5 * _A * _A * (1.0 + 1.0 / Q) - P * (1.0 - _E2) * z * z / (Q * (1.0 + Q)) - 0.5 * P * r * r)
    tmp = (r - _E2 * r0) ** 2
    U, V = np.sqrt(tmp + z * z), np.sqrt(tmp + (1.0 - _E2) * z * z)
    zo = _B * _B * z / (_A * V)
    height = U * (1.0 - _B * _B / (_A * V))
    lat = np.degrees(np.arctan((z + ep2 * zo) / r))
    lon = np.degrees(np.arctan2(y, x))
    return lat, lon, height


class AcmiWriter:
    """Appends one time frame per call; starts a new file when `count == 0` or after an episode boundary, like the
    reference (env_base.py:121-129,148-151)."""

    def __init__(self, filename='./tracks/F16SimRecording-', max_aircraft=16, name="F16", color="Red"):
        self.prefix, self.max_aircraft, self.name, self.color = filename, int(max_aircraft), name, color
        self.path, self.open_file = None, False

    def _begin(self, count):
        self.path = f"{self.prefix}{count}.txt.acmi"
        os.makedirs(os.path.dirname(os.path.abspath(self.path)), exist_ok=True)
        with open(self.path, "w", encoding="utf-8") as f:
            f.write("FileType=text/acmi/tacview\nFileVersion=2.0\n0,ReferenceTime=2023-04-01T00:00:00Z\n")
        self.open_file = True

    def write(self, count, timestamp, npos_ft, epos_ft, alt_ft, roll, pitch, yaw, episode_ended):
        if count == 0 or not self.open_file:
            self._begin(count)
        m = min(self.max_aircraft, len(npos_ft))
        lat, lon, h = enu_to_geodetic(np.asarray(epos_ft[:m]) * 0.3048, np.asarray(npos_ft[:m]) * 0.3048, np.asarray(alt_ft[:m]) * 0.3048)
        with open(self.path, "a", encoding="utf-8") as f:
            f.write(f"#{timestamp:.2f}\n")
            for i in range(m):
                f.write(f"{100 + i},T={lon[i]}|{lat[i]}|{h[i]}|{np.degrees(roll[i])}|{np.degrees(pitch[i])}|{np.degrees(yaw[i])},"
                        f"Name={self.name},Color={self.color}\n")
        if episode_ended:
            self.open_file = False
        return self.path
