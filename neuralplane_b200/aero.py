"""Device-resident copy of the 43 aero-coefficient MLPs (reference: hifi_F16.__init__, hifi_F16_AeroData.py:41-129,
which does 43 torch.load + a pandas read per construction; here one packed file is uploaded once per device)."""
import ctypes as C
import os

import numpy as np
import torch

from . import _native as nv

AERO_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "f16_aero.npz")
_cache = {}


class F16Aero:
    def __init__(self, device, path=AERO_NPZ):
        self.device = torch.device(device)
        d = np.load(path)
        self.names = [str(x) for x in d["names"]]
        desc = np.ascontiguousarray(d["desc"], dtype=np.int32)
        assert desc.shape == (nv.NUM_NETS, 12) and C.sizeof(nv.NetDesc) == 48
        norm = np.ascontiguousarray(d["norm"], dtype=np.float64)
        blob = np.ascontiguousarray(d["blob"], dtype=np.float32)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            st = nv.lib().np_aero_create(blob.ctypes.data, blob.size, C.cast(desc.ctypes.data, C.POINTER(nv.NetDesc)),
                                         norm.ctypes.data, nv.NUM_NETS, C.byref(self.handle))
        nv.check(st, "np_aero_create")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                nv.lib().np_aero_destroy(self.handle)
        except Exception:
            pass

    def index(self, name):
        return self.names.index(name)


def get_aero(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"neuralplane_b200 runs on CUDA devices only (got device={device}); there is no CPU fallback")
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _cache:
        _cache[key] = F16Aero(torch.device("cuda", key))
    return _cache[key]
