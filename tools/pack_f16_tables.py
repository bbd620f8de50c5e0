#!/usr/bin/env python
"""Pack the NASA F-16 aerodynamic tables the reference ships under example/data/*.dat (the data its MLP surrogates
were trained on; loader: example/train_model/hifi_F16_AeroData.py:10-76) into neuralplane_b200/data/f16_tables.npz,
and snapshot the reference's golden vectors envs/models/F16/model/coefs.csv (630 points x 44 table-interpolated
coefficients, written by the authors' MATLAB model; test_model.py:61-75) into tests/golden/f16_table_coefs.npz.

    python tools/pack_f16_tables.py [/root/reference]

Data only: flat arrays are Fortran-ordered (alpha fastest), index = ia + Na * ib + Na * Nb * id
(example/train_model/mexndinterp.py:38-47).
"""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# table name -> (file, axes) in the order of hifi_F16.__init__ (hifi_F16_AeroData.py:16-64)
TABLES = [
    ("Cx", "CX0120_ALPHA1_BETA1_DH1_201", "ABD1"), ("Cz", "CZ0120_ALPHA1_BETA1_DH1_301", "ABD1"),
    ("Cm", "CM0120_ALPHA1_BETA1_DH1_101", "ABD1"), ("Cy", "CY0320_ALPHA1_BETA1_401", "AB"),
    ("Cn", "CN0120_ALPHA1_BETA1_DH2_501", "ABD2"), ("Cl", "CL0120_ALPHA1_BETA1_DH2_601", "ABD2"),
    ("Cx_lef", "CX0820_ALPHA2_BETA1_202", "aB"), ("Cz_lef", "CZ0820_ALPHA2_BETA1_302", "aB"),
    ("Cm_lef", "CM0820_ALPHA2_BETA1_102", "aB"), ("Cy_lef", "CY0820_ALPHA2_BETA1_402", "aB"),
    ("Cn_lef", "CN0820_ALPHA2_BETA1_502", "aB"), ("Cl_lef", "CL0820_ALPHA2_BETA1_602", "aB"),
    ("CXq", "CX1120_ALPHA1_204", "A"), ("CZq", "CZ1120_ALPHA1_304", "A"), ("CMq", "CM1120_ALPHA1_104", "A"),
    ("CYp", "CY1220_ALPHA1_408", "A"), ("CYr", "CY1320_ALPHA1_406", "A"), ("CNr", "CN1320_ALPHA1_506", "A"),
    ("CNp", "CN1220_ALPHA1_508", "A"), ("CLp", "CL1220_ALPHA1_608", "A"), ("CLr", "CL1320_ALPHA1_606", "A"),
    ("delta_CXq_lef", "CX1420_ALPHA2_205", "a"), ("delta_CYr_lef", "CY1620_ALPHA2_407", "a"),
    ("delta_CYp_lef", "CY1520_ALPHA2_409", "a"), ("delta_CZq_lef", "CZ1420_ALPHA2_305", "a"),
    ("delta_CLr_lef", "CL1620_ALPHA2_607", "a"), ("delta_CLp_lef", "CL1520_ALPHA2_609", "a"),
    ("delta_CMq_lef", "CM1420_ALPHA2_105", "a"), ("delta_CNr_lef", "CN1620_ALPHA2_507", "a"),
    ("delta_CNp_lef", "CN1520_ALPHA2_509", "a"),
    ("Cy_r30", "CY0720_ALPHA1_BETA1_405", "AB"), ("Cn_r30", "CN0720_ALPHA1_BETA1_503", "AB"),
    ("Cl_r30", "CL0720_ALPHA1_BETA1_603", "AB"), ("Cy_a20", "CY0620_ALPHA1_BETA1_403", "AB"),
    ("Cy_a20_lef", "CY0920_ALPHA2_BETA1_404", "aB"), ("Cn_a20", "CN0620_ALPHA1_BETA1_504", "AB"),
    ("Cn_a20_lef", "CN0920_ALPHA2_BETA1_505", "aB"), ("Cl_a20", "CL0620_ALPHA1_BETA1_604", "AB"),
    ("Cl_a20_lef", "CL0920_ALPHA2_BETA1_605", "aB"),
    ("delta_CNbeta", "CN9999_ALPHA1_brett", "A"), ("delta_CLbeta", "CL9999_ALPHA1_brett", "A"),
    ("delta_Cm", "CM9999_ALPHA1_brett", "A"), ("eta_el", "ETA_DH1_brett", "D1"),
]
AXES = {"A": "ALPHA1", "a": "ALPHA2", "B": "BETA1", "D1": "DH1", "D2": "DH2"}


def read(name):
    return np.array([float(x) for x in open(os.path.join(REF, "example", "data", name + ".dat")).read().split()], dtype=np.float64)


def axes_of(code):
    out, i = [], 0
    while i < len(code):
        if code[i] == "D":
            out.append("D" + code[i + 1]); i += 2
        else:
            out.append(code[i]); i += 1
    return out


def main():
    bp = {k: read(v) for k, v in AXES.items()}
    names, offs, codes, flat = [], [], [], []
    off = 0
    for name, fname, code in TABLES:
        t = read(fname)
        assert t.size == int(np.prod([bp[a].size for a in axes_of(code)])), (name, t.size)
        names.append(name); offs.append(off); codes.append(code); flat.append(t); off += t.size
    flat = np.concatenate(flat)
    out = os.path.join(ROOT, "neuralplane_b200", "data", "f16_tables.npz")
    np.savez_compressed(out, names=np.array(names), offsets=np.array(offs, dtype=np.int32), axes=np.array(codes),
                        values=flat.astype(np.float32), values64=flat,
                        **{"bp_" + k: v.astype(np.float32) for k, v in bp.items()})
    print(out, flat.size, "table values,", len(names), "tables")
    coefs = np.loadtxt(os.path.join(REF, "envs", "models", "F16", "model", "coefs.csv"), delimiter=",")
    gout = os.path.join(ROOT, "tests", "golden", "f16_table_coefs.npz")
    np.savez_compressed(gout, inputs=coefs[:3], coefs=coefs[3:])
    print(gout, coefs.shape)


if __name__ == "__main__":
    main()
