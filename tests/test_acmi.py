"""Tacview writer host logic (SURVEY f-4): ENU -> geodetic against values produced by the reference's
envs/utils/utils.py enu_to_geodetic (tests/golden/geodetic_golden.npz), and the .acmi record format of
envs/env_base.py:121-147."""
import os

import numpy as np

from neuralplane_b200.envs.acmi import AcmiWriter, enu_to_geodetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_enu_to_geodetic_matches_reference_values():
    g = np.load(os.path.join(GOLDEN, "geodetic_golden.npz"))
    lat, lon, h = enu_to_geodetic(g["enu"][:, 0], g["enu"][:, 1], g["enu"][:, 2])
    assert np.abs(lat - g["llh"][:, 0]).max() < 1e-9 and np.abs(lon - g["llh"][:, 1]).max() < 1e-9
    assert np.abs(h - g["llh"][:, 2]).max() < 1e-3          # metres


def test_acmi_file_format(tmp_path):
    w = AcmiWriter(str(tmp_path / "trk-"), max_aircraft=2)
    z = np.zeros(3)
    p = w.write(0, 0.02, np.array([0.0, 100.0, 5.0]), z, np.array([20000.0, 19000.0, 1.0]), z, z + 0.1, z, False)
    w.write(1, 0.04, np.array([10.0, 110.0, 5.0]), z, np.array([20000.0, 19000.0, 1.0]), z, z, z, True)
    p2 = w.write(2, 0.06, z, z, z + 100, z, z, z, False)      # episode ended -> a new file
    lines = open(p).read().splitlines()
    assert lines[:3] == ["FileType=text/acmi/tacview", "FileVersion=2.0", "0,ReferenceTime=2023-04-01T00:00:00Z"]
    assert lines[3] == "#0.02" and lines[4].startswith("100,T=") and lines[5].startswith("101,T=") and lines[6] == "#0.04"
    assert lines[4].endswith("Name=F16,Color=Red") and len(lines) == 9 and len(lines[4].split("|")) == 6
    assert abs(float(lines[4].split("|")[2]) - 20000 * 0.3048) < 1e-3 and abs(float(lines[4].split("|")[4]) - np.degrees(0.1)) < 1e-9
    assert p2 != p and os.path.exists(p2)
