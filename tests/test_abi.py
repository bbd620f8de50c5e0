"""The C-ABI boundary without a GPU: the shared library loads, exports every function include/nplane.h declares (and
nothing the Python binding expects is missing), reports errors through status codes + np_last_error, and the ctypes
mirrors of the structs have the header's layout.  No compute is launched here."""
import ctypes as C
import os
import re

import numpy as np

from neuralplane_b200 import _native as nv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "nplane.h")).read()


def _declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(np_[a-z0-9_]+)\s*\(", body)) - {"np_env", "np_aero"})


def test_library_exports_every_declared_symbol():
    names = _declared_functions()
    assert len(names) >= 20
    L = nv.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/nplane.h but not exported by libnplane.so"
    assert sorted(nv.SYMBOLS) == names, (set(nv.SYMBOLS) ^ set(names))


def test_abi_version_and_struct_layout():
    L = nv.lib()
    assert L.np_version() == int(re.search(r"#define NP_ABI_VERSION (\d+)", HEADER).group(1))
    # np_env_cfg: 4 x i32, 2 x u64, 16 floats, 2 x i32, 5 floats, 2 x i32, 8 floats, 2 x i32, 1 float = 176 bytes
    assert C.sizeof(nv.EnvCfg) == 176 and nv.EnvCfg.index_stride.offset == 164 and nv.EnvCfg.combat_reward_scale.offset == 172
    assert nv.EnvCfg.seed.offset == 16 and nv.EnvCfg.dt.offset == 32 and nv.EnvCfg.model.offset == 124
    assert C.sizeof(nv.NetDesc) == 48 and C.sizeof(nv.Buffers) == 8 * 8
    fields = [f for f, _ in nv.EnvCfg._fields_]
    struct = re.search(r"typedef struct np_env_cfg \{(.*?)\} np_env_cfg;", HEADER, flags=re.S).group(1)
    struct = re.sub(r"/\*.*?\*/", "", struct, flags=re.S)
    declared = []
    for decl in struct.split(";"):
        decl = decl.strip()
        if decl:
            declared += [x.strip() for x in decl.split(None, 1)[1].split(",")]
    assert fields == declared


def _err():
    buf = C.create_string_buffer(512)
    nv.lib().np_last_error(buf, 512)
    return buf.value.decode()


def test_error_behaviour_without_a_device():
    L = nv.lib()
    out = C.c_void_p()
    assert L.np_env_create(None, None, C.byref(out)) == 1 and "null" in _err()              # NP_EINVAL
    cfg = nv.EnvCfg()
    cfg.n, cfg.ld, cfg.task, cfg.model = 10, 8, 0, 0
    assert L.np_env_create(C.byref(cfg), None, C.byref(out)) == 1 and "np_aero" in _err()   # F16 needs the nets
    cfg.model = 7
    assert L.np_env_create(C.byref(cfg), None, C.byref(out)) == 1 and "model" in _err()
    assert L.np_env_step(None, None, None, None, None) == 3 and "not bound" in _err()       # NP_ESTATE
    assert L.np_env_plan_step(None, None, 50, None, None, None) == 3
    assert L.np_env_combat_step(None, None, 5, None, None) == 3
    assert L.np_f16_nlplant(None, None, None, None, 0, 0, None) == 1
    assert L.np_combat_relgeo(None, None, None, None, 0, None) == 1
    nw = C.c_size_t()
    assert L.np_aero_pack_host(None, 0, None, None, 43, None, 0, C.byref(nw)) == 1
    d = np.load(os.path.join(ROOT, "neuralplane_b200", "data", "f16_aero.npz"))
    desc = np.ascontiguousarray(d["desc"], dtype=np.int32)
    norm = np.ascontiguousarray(d["norm"], dtype=np.float64)
    blob = np.ascontiguousarray(d["blob"], dtype=np.float32)
    bad = desc.copy(); bad[3, 5 + 1] = 21                                                    # wrong hidden width
    args = (blob.ctypes.data, blob.size, C.cast(bad.ctypes.data, C.POINTER(nv.NetDesc)), norm.ctypes.data, 43)
    assert L.np_aero_pack_host(*args, None, 0, C.byref(nw)) == 1 and "architecture" in _err()
    args = (blob.ctypes.data, 1000, C.cast(desc.ctypes.data, C.POINTER(nv.NetDesc)), norm.ctypes.data, 43)
    assert L.np_aero_pack_host(*args, None, 0, C.byref(nw)) == 1 and "too short" in _err()
    assert L.np_env_workspace_bytes(None) == 0
    cfg.ld = 64
    assert L.np_env_workspace_bytes(C.byref(cfg)) >= (18 + 12 + 1) * 64 * 4 + 64


def test_product_has_no_cpu_fallback():
    import pytest
    import torch
    from neuralplane_b200 import ControlEnv
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        ControlEnv(num_envs=4, config="heading", model="F16", device="cpu")
    if not torch.cuda.is_available():
        src = open(os.path.join(ROOT, "neuralplane_b200", "_native.py")).read()
        assert "no CPU fallback" in src
    for dirpath, _, files in os.walk(os.path.join(ROOT, "neuralplane_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} must not use the oracle"
