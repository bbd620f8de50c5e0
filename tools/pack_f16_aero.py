#!/usr/bin/env python
"""Pack the reference's 43 aero-coefficient MLPs into one data file.

Reads (build container only, where /root/reference exists):
  envs/models/F16/model/<name>.pth    state_dicts, keys layers.{0,2,4[,6]}.{weight,bias}
                                      (architectures: hifi_F16_AeroData.py:44-129)
  envs/models/F16/model/mean_std.csv  per-net input z-score and output de-normalisation
                                      (used at hifi_F16_AeroData.py:149-166 and siblings)
Writes neuralplane_b200/data/f16_aero.npz:
  names   [43]       net names in OUR canonical order (grouped by architecture/inputs)
  desc    [43, 12]   int32: n_in, sel0, sel1, sel2 (0=alpha_deg 1=beta_deg 2=el_deg, -1 none),
                     n_layers, d0..d4 (layer widths incl. input and the 1-wide output, 0 padded),
                     w_off (float offset of this net in `blob`), used (0 for delta_Czq_lef)
  norm    [43, 8]    float64: in_mean[3], in_std[3] (per selected input, 0/1 padded), out_mean, out_std
  blob    [13563]    float32: per net, per layer: W[out, in] row-major, then b[out]

The weights are model DATA (the "aero tables" of this FDM); no reference source is copied.
"""
import os
import sys

import numpy as np
import pandas as pd
import torch

REF = os.environ.get("NPLANE_REFERENCE", "/root/reference")
MODEL_DIR = os.path.join(REF, "envs/models/F16/model")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "neuralplane_b200", "data", "f16_aero.npz")

A, B, E = 0, 1, 2
# canonical order: el-dependent nets first, then (alpha,beta) nets, then alpha-only nets.
NETS = [
    # (name, inputs)                      el-dependent
    ("Cx", (A, B, E)), ("Cz", (A, B, E)), ("Cm", (A, B, E)), ("Cn", (A, B, E)), ("Cl", (A, B, E)),
    ("eta_el", (E,)),
    # (alpha, beta), hidden [20,10]: two with the "r30/a20" input normalisation, two with the "lef" one
    ("Cy", (A, B)), ("delta_Cl_a20", (A, B)), ("delta_Cx_lef", (A, B)), ("delta_Cl_lef", (A, B)),
    # (alpha, beta), hidden [20,10,5]
    ("delta_Cz_lef", (A, B)), ("delta_Cm_lef", (A, B)), ("delta_Cy_lef", (A, B)), ("delta_Cn_lef", (A, B)),
    ("delta_Cy_r30", (A, B)), ("delta_Cn_r30", (A, B)), ("delta_Cl_r30", (A, B)), ("delta_Cn_a20", (A, B)),
    # (alpha, beta), hidden [20,10,10]
    ("delta_Cy_a20", (A, B)),
    # (alpha, beta), hidden [20,20,10]
    ("delta_Cy_a20_lef", (A, B)), ("delta_Cn_a20_lef", (A, B)), ("delta_Cl_a20_lef", (A, B)),
    # alpha only, hidden [20,10]  (ALPHA1 normalisation)
    ("Cxq", (A,)), ("Czq", (A,)), ("Cmq", (A,)), ("Cyp", (A,)), ("Cyr", (A,)), ("Cnr", (A,)),
    ("Cnp", (A,)), ("Clp", (A,)), ("Clr", (A,)),
    ("delta_Cnbeta", (A,)), ("delta_Clbeta", (A,)), ("delta_Cm", (A,)),
    # alpha only, hidden [20,10]  (lef normalisation)
    ("delta_Cxq_lef", (A,)), ("delta_Cyr_lef", (A,)), ("delta_Clr_lef", (A,)), ("delta_Clp_lef", (A,)),
    ("delta_Cmq_lef", (A,)), ("delta_Cnr_lef", (A,)), ("delta_Cnp_lef", (A,)),
    # alpha only, hidden [20,10,5]
    ("delta_Cyp_lef", (A,)),
    # evaluated by the reference (hifi_F16_AeroData.py:786) but never used (F16_dynamics.py:167-175)
    ("delta_Czq_lef", (A,)),
]
UNUSED = {"delta_Czq_lef"}


def main():
    csv = pd.read_csv(os.path.join(MODEL_DIR, "mean_std.csv")).set_index("name")
    names, desc, norm, chunks = [], [], [], []
    off = 0
    for name, sel in NETS:
        sd = torch.load(os.path.join(MODEL_DIR, name + ".pth"), map_location="cpu")
        keys = sorted({int(k.split(".")[1]) for k in sd})
        dims = [sd[f"layers.{keys[0]}.weight"].shape[1]]
        w_off = off
        for k in keys:
            W = sd[f"layers.{k}.weight"].to(torch.float32).numpy()
            b = sd[f"layers.{k}.bias"].to(torch.float32).numpy()
            assert W.shape[1] == dims[-1]
            dims.append(W.shape[0])
            chunks += [W.reshape(-1), b.reshape(-1)]
            off += W.size + b.size
        assert dims[0] == len(sel) and dims[-1] == 1, (name, dims, sel)
        row = csv.loc[name]
        cols = {A: ("alpha_mean", "alpha_std"), B: ("beta_mean", "beta_std"), E: ("el_mean", "el_std")}
        im = [float(row[cols[s][0]]) for s in sel] + [0.0] * (3 - len(sel))
        isd = [float(row[cols[s][1]]) for s in sel] + [1.0] * (3 - len(sel))
        names.append(name)
        desc.append([len(sel)] + list(sel) + [-1] * (3 - len(sel)) + [len(keys)] + dims + [0] * (5 - len(dims))
                    + [w_off, 0 if name in UNUSED else 1])
        norm.append(im + isd + [float(row["mean"]), float(row["std"])])
    blob = np.concatenate(chunks).astype(np.float32)
    assert len(names) == 43 and blob.size == 13563, (len(names), blob.size)
    np.savez(OUT, names=np.array(names), desc=np.array(desc, dtype=np.int32),
             norm=np.array(norm, dtype=np.float64), blob=blob)
    print("wrote", os.path.normpath(OUT), "nets", len(names), "floats", blob.size)


if __name__ == "__main__":
    sys.exit(main())
