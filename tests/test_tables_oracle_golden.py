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
