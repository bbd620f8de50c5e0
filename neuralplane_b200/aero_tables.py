"""Table aero back-end (SURVEY f-3): the NASA F-16 tables (reference: example/data/*.dat, loaded by
example/train_model/hifi_F16_AeroData.py:10-76) on the device, evaluated by np_f16_table_coeffs (K6)."""
import ctypes as C
import os

import numpy as np
import torch

from . import _native as nv

TABLES_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "f16_tables.npz")
COEF_NAMES = ("Cx", "Cz", "Cm", "Cy", "Cn", "Cl", "Cxq", "Cyr", "Cyp", "Czq", "Clr", "Clp", "Cmq", "Cnr", "Cnp",
              "delta_Cx_lef", "delta_Cz_lef", "delta_Cm_lef", "delta_Cy_lef", "delta_Cn_lef", "delta_Cl_lef",
              "delta_Cxq_lef", "delta_Cyr_lef", "delta_Cyp_lef", "delta_Czq_lef", "delta_Clr_lef", "delta_Clp_lef",
              "delta_Cmq_lef", "delta_Cnr_lef", "delta_Cnp_lef", "delta_Cy_r30", "delta_Cn_r30", "delta_Cl_r30",
              "delta_Cy_a20", "delta_Cn_a20", "delta_Cl_a20", "delta_Cy_a20_lef", "delta_Cn_a20_lef", "delta_Cl_a20_lef",
              "delta_Cnbeta", "delta_Clbeta", "delta_Cm", "eta_el", "delta_Cm_ds")


class F16AeroTables:
    def __init__(self, device, path=TABLES_NPZ):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("neuralplane_b200 runs on CUDA devices only; there is no CPU fallback")
        d = np.load(path)
        bp = np.concatenate([d["bp_" + k] for k in ("A", "a", "B", "D1", "D2")]).astype(np.float32)
        sizes = np.array([d["bp_" + k].size for k in ("A", "a", "B", "D1", "D2")], dtype=np.int32)
        values = np.ascontiguousarray(d["values"], dtype=np.float32)
        offsets = np.ascontiguousarray(d["offsets"], dtype=np.int32)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            st = nv.lib().np_tables_create(bp.ctypes.data, sizes.ctypes.data, values.ctypes.data, offsets.ctypes.data, offsets.size,
                                           values.size, C.byref(self.handle))
        nv.check(st, "np_tables_create")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                nv.lib().np_tables_destroy(self.handle)
        except Exception:
            pass

    def coefficients(self, alpha_deg, beta_deg, el_deg):
        """[n, 44] coefficients (columns = COEF_NAMES, the reference's coefs.csv row order) for float32 CUDA inputs [n]."""
        a, b, e = (x.to(device=self.device, dtype=torch.float32).contiguous() for x in (alpha_deg, beta_deg, el_deg))
        n = a.numel()
        out = torch.empty((44, n), dtype=torch.float32, device=self.device)
        st = nv.lib().np_f16_table_coeffs(self.handle, a.data_ptr(), b.data_ptr(), e.data_ptr(), out.data_ptr(), n, n,
                                          torch.cuda.current_stream(self.device).cuda_stream)
        nv.check(st, "np_f16_table_coeffs")
        return out.t()


_cache = {}


def get_tables(device):
    """The per-device table image (shared by every table-backed env on that device)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"neuralplane_b200 runs on CUDA devices only (got device={device}); there is no CPU fallback")
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _cache:
        _cache[key] = F16AeroTables(torch.device("cuda", key))
    return _cache[key]
