"""The table aero oracle (oracle/f16_tables_oracle.py) against the reference's golden vectors
envs/models/F16/model/coefs.csv (snapshot: tests/golden/f16_table_coefs.npz, made by tools/pack_f16_tables.py):
630 (alpha, beta, el) points x 44 coefficients.  Rows built on the ALPHA2 grid (-20..45 deg: the lef / damping_lef /
a20_lef groups) are only defined for alpha <= 45, which is how the reference's own comparison uses them
(test_model.py:163-258,282-320: first 400 columns)."""
import os

import numpy as np

from oracle.f16_tables_oracle import COEF_NAMES, F16Tables

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALPHA2_ROWS = [k for k, nm in enumerate(COEF_NAMES) if nm.endswith("_lef")]


def test_table_coefficients_match_the_references_golden_vectors():
    g = np.load(os.path.join(GOLDEN, "f16_table_coefs.npz"))
    a, b, e = g["inputs"]
    out = F16Tables().coefficients(a, b, e)
    assert out.shape == g["coefs"].shape == (44, 630) and len(COEF_NAMES) == 44
    in_alpha2 = a <= 45.0
    assert in_alpha2.sum() >= 400
    for k, name in enumerate(COEF_NAMES):
        cols = in_alpha2 if k in ALPHA2_ROWS else slice(None)
        assert np.abs(out[k][cols] - g["coefs"][k][cols]).max() <= 1e-12, name


def test_table_aero_adapter_feeds_the_env_oracle():
    """TableAero plugs the tables into F16EnvOracle (the oracle of ControlEnv(model='F16_tables')): every coefficient
    nlplant consumes is present, equals F16Tables.coefficients, and fp32 / fp64 env steps agree to fp32 rounding."""
    import torch
    from oracle import tapes
    from oracle.f16_oracle import F16EnvOracle
    from oracle.f16_tables_oracle import COEF_NAMES, F16Tables, TableAero
    a, b, e = torch.tensor([5.0, 20.0, -5.0]), torch.tensor([0.0, 5.0, -10.0]), torch.tensor([-2.0, 5.0, 10.0])
    c = TableAero(dtype=torch.float64).coeffs(a, b, e)
    rows = F16Tables().coefficients(a.numpy().astype(np.float64), b.numpy().astype(np.float64), e.numpy().astype(np.float64))
    assert set(c) == set(COEF_NAMES)
    for k, name in enumerate(COEF_NAMES):
        assert np.array_equal(c[name].numpy(), rows[k]), name
    n = 64
    o32 = F16EnvOracle(n, "heading", aero=TableAero(dtype=torch.float32))
    o64 = F16EnvOracle(n, "heading", aero=TableAero(dtype=torch.float64), dtype=torch.float64)
    d0 = tapes.reset_draw_tape(3, 0, n)
    o32.reset(torch.from_numpy(d0)); o64.reset(torch.from_numpy(d0).double())
    for k in range(1, 11):
        act, d = tapes.action_tape(3, k, n, 0.3), tapes.reset_draw_tape(3, k, n)
        o32.step(torch.from_numpy(act), torch.from_numpy(d)); o64.step(torch.from_numpy(act).double(), torch.from_numpy(d).double())
    assert np.allclose(o32.s.numpy(), o64.s.numpy(), rtol=2e-5, atol=1e-5)
    assert torch.isfinite(o64.s).all() and float(o64.s[:, 6].min()) > 500
