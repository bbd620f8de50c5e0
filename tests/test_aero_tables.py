"""Host logic of the aero image (np_aero_pack_host): the 21 one-input ReLU MLPs are converted into exact
piecewise-linear tables (neuralplane_b200/csrc/aero_pack.h).  Checked here on the CPU against the oracle's MLP
evaluation (hifi_F16_AeroData.py:12-37,149-166): the fp32 table evaluation must be at least as close to the float64
evaluation of the MLP as the reference's own fp32 MLP arithmetic is, and within 3e-6 of it (relative to the
coefficient's range)."""
import ctypes as C

import numpy as np
import torch

from neuralplane_b200 import _native as nv
from oracle.f16_oracle import AERO_NPZ, AeroNets

HDR = dict(words=0, c0=1, levels_a=2, bp_a=3, segmap=4, nseg_a=5, ent_a=6, levels_e=7, bp_e=8, ent_e=9, taboff=12)
FIRST_A1, NUM_A1, ETA_EL, ROW = 22, 21, 5, 24


def _image():
    d = np.load(AERO_NPZ)
    desc = np.ascontiguousarray(d["desc"], dtype=np.int32)
    norm = np.ascontiguousarray(d["norm"], dtype=np.float64)
    blob = np.ascontiguousarray(d["blob"], dtype=np.float32)
    L = nv.lib()
    nw = C.c_size_t()
    args = (blob.ctypes.data, blob.size, C.cast(desc.ctypes.data, C.POINTER(nv.NetDesc)), norm.ctypes.data, 43)
    nv.check(L.np_aero_pack_host(*args, None, 0, C.byref(nw)), "np_aero_pack_host")
    out = np.zeros(nw.value, dtype=np.uint32)
    nv.check(L.np_aero_pack_host(*args, out.ctypes.data, out.size, C.byref(nw)), "np_aero_pack_host")
    return out


def _search(bp, levels, x):
    """The device's branchless binary search (f16_device.cuh pwl_search), vectorised."""
    pos = np.zeros(x.shape, dtype=np.int64)
    step = 1 << (levels - 1)
    while step:
        pos += np.where(bp[pos + step - 1] <= x, step, 0)
        step >>= 1
    return pos


def _eval(ent, idx, x):
    e = ent[idx]
    return (e[:, 1].astype(np.float64) + e[:, 2].astype(np.float64) * (x - e[:, 0]).astype(np.float64)).astype(np.float32)


def test_image_header_and_alignment():
    im = _image()
    h = im.view(np.int32)
    assert h[HDR["words"]] == im.size and im.size * 4 <= 72 * 1024 and im.size % 4 == 0
    for key in ("c0", "bp_a", "segmap", "ent_a", "bp_e", "ent_e"):
        assert h[HDR[key]] % 4 == 0 and 0 < h[HDR[key]] < im.size, key
    bp = im.view(np.float32)[h[HDR["bp_a"]]: h[HDR["bp_a"]] + (1 << h[HDR["levels_a"]]) - 1]
    assert np.all(np.diff(bp[np.isfinite(bp)]) > 0) and np.isfinite(bp).sum() == h[HDR["nseg_a"]] - 1


def test_one_input_nets_as_tables_match_the_mlps():
    im = _image()
    h, f = im.view(np.int32), im.view(np.float32)
    a32, a64 = AeroNets(), AeroNets(dtype=torch.float64)
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.uniform(-40, 100, 150000), rng.uniform(-400, 400, 2000), [-1e5, 1e5, 0.0]]).astype(np.float32)
    xt = torch.from_numpy(x)
    # alpha nets: merged search -> segment map -> per-net entry
    la, bp = h[HDR["levels_a"]], f[h[HDR["bp_a"]]:]
    m = _search(bp, la, x)
    assert m.max() <= h[HDR["nseg_a"]] - 1
    segmap = im[h[HDR["segmap"]]: h[HDR["segmap"]] + h[HDR["nseg_a"]] * (ROW // 4)].view(np.uint8).reshape(-1, ROW)
    ent = f[h[HDR["ent_a"]]: h[HDR["bp_e"]]]
    ent = ent[: (ent.size // 4) * 4].reshape(-1, 4)
    worst = 0.0
    for k in range(FIRST_A1, FIRST_A1 + NUM_A1):
        idx = h[HDR["taboff"] + k - FIRST_A1] + segmap[m, k - FIRST_A1].astype(np.int64)
        y = _eval(ent, idx, x)
        r64 = a64.eval_net(k, xt.double(), xt.double(), xt.double()).numpy()
        r32 = a32.eval_net(k, xt, xt, xt).numpy()
        sc = np.abs(r64[:150000]).max()
        e_tab, e_ref = np.abs(y - r64) / sc, np.abs(r32 - r64) / sc
        core = slice(0, 150000)
        assert e_tab[core].max() <= max(1.5 * e_ref[core].max(), 5e-7), (a32.names[k], e_tab[core].max(), e_ref[core].max())
        assert (np.abs(y - r32) / sc)[core].max() < 3e-6, a32.names[k]
        # far outside the envelope the outermost linear pieces must still be the net
        assert np.all(np.abs(y - r64) <= 2e-6 * (np.abs(r64) + sc)), a32.names[k]
        worst = max(worst, e_tab[core].max())
    # eta_el(el)
    le, bpe = h[HDR["levels_e"]], f[h[HDR["bp_e"]]:]
    ente = f[h[HDR["ent_e"]]: h[HDR["words"]]].reshape(-1, 4)
    y = _eval(ente, _search(bpe, le, x), x)
    r64 = a64.eval_net(ETA_EL, xt.double(), xt.double(), xt.double()).numpy()
    r32 = a32.eval_net(ETA_EL, xt, xt, xt).numpy()
    sc = np.abs(r64[:150000]).max()
    assert (np.abs(y - r64) / sc)[:150000].max() <= max(1.5 * (np.abs(r32 - r64) / sc)[:150000].max(), 5e-7)
    print("worst table-vs-fp64 error over the alpha nets: %.2e of range" % worst)
