"""GPU parity of the table aero back-end (K6, np_f16_table_coeffs) against the reference's golden vectors
envs/models/F16/model/coefs.csv and the float64 tables oracle.  fp32 multilinear interpolation of O(1) values:
tolerance 2e-6 of each coefficient's range."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_table_coefficients_vs_reference_golden_vectors():
    from neuralplane_b200.aero_tables import COEF_NAMES, F16AeroTables
    from oracle.f16_tables_oracle import COEF_NAMES as ORACLE_NAMES
    assert tuple(COEF_NAMES) == tuple(ORACLE_NAMES)
    g = np.load(os.path.join(GOLDEN, "f16_table_coefs.npz"))
    a, b, e = (torch.from_numpy(x.astype(np.float32)).cuda() for x in g["inputs"])
    got = F16AeroTables("cuda:0").coefficients(a, b, e).cpu().numpy().T          # [44, 630]
    in_alpha2 = g["inputs"][0] <= 45.0
    for k, name in enumerate(COEF_NAMES):
        cols = in_alpha2 if name.endswith("_lef") else slice(None)
        ref = g["coefs"][k][cols]
        scale = max(np.abs(g["coefs"][k]).max(), 1e-3)
        assert np.abs(got[k][cols] - ref).max() <= 2e-6 * scale + 1e-7, (name, np.abs(got[k][cols] - ref).max(), scale)


def test_table_coefficients_vs_oracle_random_and_clamped():
    from neuralplane_b200.aero_tables import F16AeroTables
    from oracle.f16_tables_oracle import F16Tables
    rng = np.random.RandomState(3)
    n = 100_001
    a = rng.uniform(-30, 100, n).astype(np.float32); b = rng.uniform(-40, 40, n).astype(np.float32); e = rng.uniform(-30, 30, n).astype(np.float32)
    a[:50] = np.array([-20, 45, 90, 0, 5] * 10, np.float32); b[:50] = 0.0; e[:50] = np.array([-25, 0, 25, 10, -10] * 10, np.float32)   # on grid lines
    got = F16AeroTables("cuda:0").coefficients(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(e).cuda()).cpu().numpy().T
    ref = F16Tables().coefficients(a.astype(np.float64), b.astype(np.float64), e.astype(np.float64))
    scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-3)
    assert (np.abs(got - ref) / scale).max() <= 3e-6
