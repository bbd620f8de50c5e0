"""CPU oracle of the TABLE aero back-end (SURVEY f-3).  TEST INFRASTRUCTURE ONLY.

The reference trains its 43 MLP surrogates on the NASA F-16 tables under example/data/*.dat, evaluated by multilinear
interpolation (example/train_model/mexndinterp.py:84-110: hyper-cube lookup + successive linear interpolation,
alpha fastest) and combined into 44 coefficients by the group functions of example/train_model/hifi_F16_AeroData.py:
406-483.  This module restates that in float64 numpy.  Pinned by the reference's golden vectors
envs/models/F16/model/coefs.csv (630 points x 44 coefficients from the authors' MATLAB model, test_model.py:61-75;
snapshot in tests/golden/f16_table_coefs.npz): agreement to 1e-12 (tests/test_tables_oracle_golden.py).

Out-of-grid inputs: the reference prints 'Point lies out data grid' and returns garbage for the whole batch
(mexndinterp.py:20-21); here they are clamped to the grid (documented difference; the golden vectors never leave it
on the columns the reference's own test uses).
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
TABLES_NPZ = os.path.join(_HERE, "..", "neuralplane_b200", "data", "f16_tables.npz")

# the 44 coefficients in coefs.csv row order (rows 3..46), test_model.py:69-338
COEF_NAMES = ("Cx", "Cz", "Cm", "Cy", "Cn", "Cl",
              "Cxq", "Cyr", "Cyp", "Czq", "Clr", "Clp", "Cmq", "Cnr", "Cnp",
              "delta_Cx_lef", "delta_Cz_lef", "delta_Cm_lef", "delta_Cy_lef", "delta_Cn_lef", "delta_Cl_lef",
              "delta_Cxq_lef", "delta_Cyr_lef", "delta_Cyp_lef", "delta_Czq_lef", "delta_Clr_lef", "delta_Clp_lef",
              "delta_Cmq_lef", "delta_Cnr_lef", "delta_Cnp_lef",
              "delta_Cy_r30", "delta_Cn_r30", "delta_Cl_r30",
              "delta_Cy_a20", "delta_Cn_a20", "delta_Cl_a20", "delta_Cy_a20_lef", "delta_Cn_a20_lef", "delta_Cl_a20_lef",
              "delta_Cnbeta", "delta_Clbeta", "delta_Cm", "eta_el", "delta_Cm_ds")


class F16Tables:
    def __init__(self, path=TABLES_NPZ):
        d = np.load(path)
        self.bp = {k[3:]: d[k].astype(np.float64) for k in d.files if k.startswith("bp_")}
        self.names = [str(x) for x in d["names"]]
        self.axes = [str(x) for x in d["axes"]]
        self.off = d["offsets"]
        self.values = d["values64"]

    @staticmethod
    def _axes(code):
        out, i = [], 0
        while i < len(code):
            if code[i] == "D":
                out.append(code[i:i + 2]); i += 2
            else:
                out.append(code[i]); i += 1
        return out

    def interp(self, name, *x):
        """Multilinear interpolation of one table (mexndinterp.py interpn), inputs clamped to the grid."""
        k = self.names.index(name)
        ax = self._axes(self.axes[k])
        dims = [self.bp[a].size for a in ax]
        tab = self.values[self.off[k]: self.off[k] + int(np.prod(dims))].reshape(dims, order="F")
        lo, lam = [], []
        for a, v in zip(ax, x):
            g = self.bp[a]
            v = np.clip(np.asarray(v, dtype=np.float64), g[0], g[-1])
            i = np.clip(np.searchsorted(g, v, side="right") - 1, 0, g.size - 2)
            lo.append(i)
            lam.append((v - g[i]) / (g[i + 1] - g[i]))
        out = np.zeros_like(lam[0])
        for corner in range(1 << len(ax)):
            w = np.ones_like(lam[0])
            idx = []
            for j in range(len(ax)):
                bit = (corner >> j) & 1
                w = w * (lam[j] if bit else 1 - lam[j])
                idx.append(lo[j] + bit)
            out = out + w * tab[tuple(idx)]
        return out

    def coefficients(self, alpha, beta, el):
        """The 44 coefficients of hifi_C / hifi_damping / hifi_C_lef / hifi_damping_lef / hifi_rudder / hifi_ailerons /
        hifi_other_coeffs (hifi_F16_AeroData.py:406-483) in coefs.csv row order: [44, n]."""
        t, z = self.interp, np.zeros_like(np.asarray(alpha, dtype=np.float64))
        Cx, Cz, Cm = t("Cx", alpha, beta, el), t("Cz", alpha, beta, el), t("Cm", alpha, beta, el)
        Cy, Cn, Cl = t("Cy", alpha, beta), t("Cn", alpha, beta, el), t("Cl", alpha, beta, el)
        Cx0, Cz0, Cm0 = t("Cx", alpha, beta, z), t("Cz", alpha, beta, z), t("Cm", alpha, beta, z)
        Cn0, Cl0 = t("Cn", alpha, beta, z), t("Cl", alpha, beta, z)
        Cy_lef, Cn_lef, Cl_lef = t("Cy_lef", alpha, beta), t("Cn_lef", alpha, beta), t("Cl_lef", alpha, beta)
        dCy_a20, dCn_a20, dCl_a20 = t("Cy_a20", alpha, beta) - Cy, t("Cn_a20", alpha, beta) - Cn0, t("Cl_a20", alpha, beta) - Cl0
        rows = [Cx, Cz, Cm, Cy, Cn, Cl,
                t("CXq", alpha), t("CYr", alpha), t("CYp", alpha), t("CZq", alpha), t("CLr", alpha), t("CLp", alpha),
                t("CMq", alpha), t("CNr", alpha), t("CNp", alpha),
                t("Cx_lef", alpha, beta) - Cx0, t("Cz_lef", alpha, beta) - Cz0, t("Cm_lef", alpha, beta) - Cm0,
                Cy_lef - Cy, Cn_lef - Cn0, Cl_lef - Cl0,
                t("delta_CXq_lef", alpha), t("delta_CYr_lef", alpha), t("delta_CYp_lef", alpha), t("delta_CZq_lef", alpha),
                t("delta_CLr_lef", alpha), t("delta_CLp_lef", alpha), t("delta_CMq_lef", alpha), t("delta_CNr_lef", alpha),
                t("delta_CNp_lef", alpha),
                t("Cy_r30", alpha, beta) - Cy, t("Cn_r30", alpha, beta) - Cn0, t("Cl_r30", alpha, beta) - Cl0,
                dCy_a20, dCn_a20, dCl_a20,                       # coefs.csv groups the three a20 rows, then the three a20_lef
                t("Cy_a20_lef", alpha, beta) - Cy_lef - dCy_a20, t("Cn_a20_lef", alpha, beta) - Cn_lef - dCn_a20,
                t("Cl_a20_lef", alpha, beta) - Cl_lef - dCl_a20,
                t("delta_CNbeta", alpha), t("delta_CLbeta", alpha), t("delta_Cm", alpha), t("eta_el", el), z]
        return np.stack(rows, 0)


class TableAero:
    """Drop-in for f16_oracle.AeroNets: `coeffs(alpha_deg, beta_deg, el_deg)` -> {name: tensor} from the tables, so that
    F16EnvOracle(aero=TableAero(...)) is the oracle of a table-backed env (ControlEnv(model='F16_tables')).  Both halves
    are pinned on their own: the env logic bit-exactly to the reference env (tests/test_oracle_golden.py), the
    coefficients to coefs.csv at 1e-12 (tests/test_tables_oracle_golden.py).  The reference itself never flies the
    tables through its env (its hifi_F16 holds the MLP surrogates), so there is no end-to-end reference run to pin the
    composition to.  Interpolation runs in float64; the result is cast to `dtype`."""

    def __init__(self, dtype=None, path=TABLES_NPZ):
        import torch
        self.torch = torch
        self.dtype = dtype or torch.float32
        self.tables = F16Tables(path)

    def coeffs(self, alpha_deg, beta_deg, el_deg):
        t = self.torch
        a, b, e = (x.detach().to(t.float64).numpy() for x in (alpha_deg, beta_deg, el_deg))
        rows = self.tables.coefficients(a, b, e)
        return {name: t.from_numpy(np.ascontiguousarray(rows[k])).to(self.dtype) for k, name in enumerate(COEF_NAMES)}
