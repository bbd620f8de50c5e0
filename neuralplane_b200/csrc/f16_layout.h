// Net table and device-image layout shared by the host packer (aero_pack.h) and the kernels (f16_device.cuh).
// Canonical net order = neuralplane_b200/data/f16_aero.npz (tools/pack_f16_aero.py).
#pragma once
#include <stdint.h>

namespace npl {

struct NetArch {
  int nin, h1, h2, h3;  // h3 == 0: two hidden layers
};

constexpr int kNumNets = 43;
constexpr int kNumUsed = 42;  // net 42 (delta_Czq_lef) is evaluated by the reference but never consumed (F16_dynamics.py:167-175)

// coefficient slots == net indices
enum Coef : int {
  kCx = 0, kCz, kCm, kCn, kCl, kEtaEl,                                  // el-dependent
  kCy, kdCl_a20, kdCx_lef, kdCl_lef,                                     // (a,b) [20,10]
  kdCz_lef, kdCm_lef, kdCy_lef, kdCn_lef,                                // (a,b) [20,10,5] lef
  kdCy_r30, kdCn_r30, kdCl_r30, kdCn_a20,                                // (a,b) [20,10,5] r
  kdCy_a20,                                                              // (a,b) [20,10,10]
  kdCy_a20_lef, kdCn_a20_lef, kdCl_a20_lef,                              // (a,b) [20,20,10]
  kCxq, kCzq, kCmq, kCyp, kCyr, kCnr, kCnp, kClp, kClr,                  // (a) ALPHA1
  kdCnbeta, kdClbeta, kdCm,
  kdCxq_lef, kdCyr_lef, kdClr_lef, kdClp_lef, kdCmq_lef, kdCnr_lef, kdCnp_lef,  // (a) lef
  kdCyp_lef,                                                             // (a) [20,10,5] lef
  kdCzq_lef                                                              // unused
};
constexpr int kNumSlots = 22;                 // coefficient slots kept in shared memory: nets 0..21 (MLPs + eta_el)
constexpr int kFirstAB2 = kCy;                // the 16 two-input (alpha, beta) MLPs: nets [6, 22)
constexpr int kNumAB2 = kCxq - kCy;           // 16
constexpr int kFirstA1 = kCxq;                // the 21 alpha-only nets: [22, 43) -> piecewise-linear tables
constexpr int kNumA1 = kNumNets - kFirstA1;   // 21

constexpr NetArch arch_of(int k) {
  return k <= kCl ? NetArch{3, 20, 10, 0}
       : k == kEtaEl ? NetArch{1, 20, 10, 0}
       : k <= kdCl_lef ? NetArch{2, 20, 10, 0}
       : k <= kdCn_a20 ? NetArch{2, 20, 10, 5}
       : k == kdCy_a20 ? NetArch{2, 20, 10, 10}
       : k <= kdCl_a20_lef ? NetArch{2, 20, 20, 10}
       : k <= kdCnp_lef ? NetArch{1, 20, 10, 0}
       : k == kdCyp_lef ? NetArch{1, 20, 10, 5}
       : NetArch{1, 20, 10, 0};
}

// Input normalisation groups (mean_std.csv): which (mean, std) pair z-scores each input of a net.
enum ZId : int { kZaC = 0, kZbC, kZeC, kZeEta, kZaR, kZbR, kZaLef2, kZaA1, kZaLef1, kNumZ };
// (kZbR is also the beta normalisation of the lef-2D nets; kZeEta / kZaA1 / kZaLef1 are folded into the tables.)
struct ZSel { int a, b, e; };
constexpr ZSel zsel_of(int k) {
  return k <= kCl ? ZSel{kZaC, kZbC, kZeC}
       : k == kEtaEl ? ZSel{-1, -1, kZeEta}
       : (k == kCy || k == kdCl_a20 || (k >= kdCy_r30 && k <= kdCy_a20)) ? ZSel{kZaR, kZbR, -1}
       : (k <= kdCl_a20_lef) ? ZSel{kZaLef2, kZbR, -1}
       : (k <= kdCm) ? ZSel{kZaA1, -1, -1}
       : ZSel{kZaLef1, -1, -1};
}

// ---- image layout (32-bit words) ------------------------------------------------------------------
constexpr int pad4(int x) { return (x + 3) & ~3; }
constexpr int layer_floats(int in, int out) { return pad4(out + in * out); }
constexpr int net_floats(NetArch a) {
  return layer_floats(a.nin, a.h1) + layer_floats(a.h1, a.h2) +
         (a.h3 ? layer_floats(a.h2, a.h3) + layer_floats(a.h3, 1) : layer_floats(a.h2, 1));
}
// header words (ints)
enum Hdr : int {
  kHdrWordsTotal = 0, kHdrC0, kHdrLevelsA, kHdrBpA, kHdrSegmap, kHdrNumSegA, kHdrEntA, kHdrLevelsE, kHdrBpE, kHdrEntE,
  kHdrTabOff = 12,                    // [21] first float4 entry of each alpha net, relative to kHdrEntA
  kHdrWords = 36
};
static_assert(kHdrTabOff + kNumA1 <= kHdrWords, "header too small");
constexpr int kZnormOff = kHdrWords;
constexpr int kOnormOff = kZnormOff + pad4(2 * kNumZ);
constexpr int kWeightOff = kOnormOff + pad4(2 * kNumNets);
constexpr bool is_mlp(int k) { return k < kFirstA1 && k != kEtaEl; }
constexpr int mlp_offset(int k) {      // word offset of MLP net k's weights (k in 0..21, k != eta_el)
  int off = kWeightOff;
  for (int i = 0; i < k; ++i)
    if (is_mlp(i)) off += net_floats(arch_of(i));
  return off;
}
constexpr int kWeightFloats = mlp_offset(kFirstA1) - kWeightOff;
constexpr int kLevelsA = 10;           // merged alpha breakpoint list: 2^10 - 1 slots (+inf padded)
constexpr int kLevelsE = 6;            // eta_el breakpoint list: 2^6 - 1 slots
constexpr int kSegmapRowBytes = 24;    // 21 alpha nets, padded to 6 words
constexpr int kMaxAeroBytes = 72 * 1024;
static_assert(kWeightOff % 4 == 0 && kWeightFloats % 4 == 0, "16-byte alignment of the weight block");

}  // namespace npl
