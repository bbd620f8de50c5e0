// Device-side building blocks of the F-16 step: the aero-coefficient surrogates (22 MLPs on packed FFMA2 for two
// aircraft per thread, 21 one-input nets as exact piecewise-linear tables), the 6-DoF equations of motion,
// atmosphere, observation helpers and the counter-based RNG.
//
// Numerics contract: this translation unit is compiled with -fmad=false, so every `a*b+c` written
// below rounds twice exactly like the reference's eager PyTorch ops do; the MLP inner products use
// explicit FMAs (the reference's nn.Linear goes through an FMA-based sgemm with unspecified
// summation order, so no order is "the" reference order there).  Division and sqrt are IEEE
// (nvcc defaults -prec-div/-prec-sqrt=true); sinf/cosf/tanf/powf are the accurate CUDA libm versions.
//
// Reference formulas: envs/models/F16/F16_dynamics.py:22-35 (atmos), :37-229 (nlplant),
// envs/models/F16/hifi_F16_AeroData.py:12-37,149-166,748-819 (MLPs), envs/models/F16_model.py:132-162
// (accelerations, EAS2TAS), envs/utils/utils.py:144-154 (wrap_PI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "f16_layout.h"

namespace npl {

// ------------------------------------------------------------------------------------------------
// Shared-memory access helpers.  Weights / tables are read with explicit ld.shared (warp-broadcast for the
// weights).  Besides pinning the access width, the non-volatile asm keeps several thousand weight loads out of
// the compiler's alias analysis; ordering against the one-time TMA staging is carried by the data dependence on
// the base address, which callers obtain from aero_base_after_staging() after the staging barrier.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 q;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(addr));
  return q;
}
__device__ __forceinline__ float lds32f(uint32_t addr) {
  float q;
  asm("ld.shared.f32 %0, [%1];" : "=f"(q) : "r"(addr));
  return q;
}
__device__ __forceinline__ uint32_t lds32u(uint32_t addr) {
  uint32_t q;
  asm("ld.shared.u32 %0, [%1];" : "=r"(q) : "r"(addr));
  return q;
}
// Re-materialise an address through a volatile mov: loads keyed on the result can be neither hoisted out of the
// enclosing loop nor merged with an earlier evaluation's loads (which would pin hundreds of registers).
__device__ __forceinline__ uint32_t opaque_u32(uint32_t x) {
  uint32_t y;
  asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t aero_base_after_staging(const void* blob_smem) {
  uint32_t a;
  asm volatile("mov.u32 %0, %1;" : "=r"(a) : "r"((uint32_t)__cvta_generic_to_shared(blob_smem)) : "memory");
  return a;
}

// ------------------------------------------------------------------------------------------------
// MLP evaluation: TWO aircraft per thread packed in the halves of a 64-bit register pair; every weight is one
// warp-broadcast LDS.128 lane feeding an FFMA2 (fma.rn.f32x2 with a scalar-broadcast multiplicand), i.e. one
// issue slot per two multiply-adds.
// ------------------------------------------------------------------------------------------------
template <int IN, int OUT, bool RELU>
__device__ __forceinline__ void dense2(uint32_t w, const float2 (&x)[IN], float2 (&y)[OUT]) {
  constexpr int NF = OUT + IN * OUT;
  constexpr int NV = (NF + 3) / 4;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 q = lds128(w + 16 * v);  // constant offsets fold into the LDS immediate
    const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int f = 4 * v + k;
      if (f < OUT) {
        y[f] = make_float2(e[k], e[k]);
      } else if (f < NF) {
        const int g = f - OUT;
        y[g % OUT] = __ffma2_rn(make_float2(e[k], e[k]), x[g / OUT], y[g % OUT]);
      }
    }
  }
  if (RELU) {
#pragma unroll
    for (int j = 0; j < OUT; ++j) y[j] = make_float2(fmaxf(y[j].x, 0.0f), fmaxf(y[j].y, 0.0f));
  }
}

template <int NIN, int H1, int H2, int H3>
__device__ __forceinline__ float2 mlp2(uint32_t w, float2 z0, float2 z1, float2 z2) {
  float2 x[NIN];
  x[0] = z0;
  if (NIN > 1) x[NIN > 1 ? 1 : 0] = z1;
  if (NIN > 2) x[NIN > 2 ? 2 : 0] = z2;
  float2 a[H1];
  dense2<NIN, H1, true>(w, x, a);
  w += 4 * layer_floats(NIN, H1);
  float2 b[H2];
  dense2<H1, H2, true>(w, a, b);
  w += 4 * layer_floats(H1, H2);
  float2 y[1];
  if constexpr (H3 > 0) {
    float2 c[H3];
    dense2<H2, H3, true>(w, b, c);
    w += 4 * layer_floats(H2, H3);
    dense2<H3, 1, false>(w, c, y);
  } else {
    dense2<H2, 1, false>(w, b, y);
  }
  return y[0];
}

// z-scores of (alpha_deg, beta_deg, el_deg) for the normalisation groups of the MLP nets: (x - mean) / std with a
// true divide (hifi_F16_AeroData.py:32-33).  Lanes .x / .y are the thread's two aircraft.
struct ZIn2 {
  float2 z[kNumZ];
};
__device__ __forceinline__ float2 zscore2(const float* __restrict__ blob, int zid, float2 x) {
  const float2 ms = reinterpret_cast<const float2*>(blob + kZnormOff)[zid];
  return make_float2((x.x - ms.x) / ms.y, (x.y - ms.x) / ms.y);
}
__device__ __forceinline__ void zscores_ab2(const float* __restrict__ blob, float2 alpha_deg, float2 beta_deg, ZIn2& o) {
  o.z[kZaC] = zscore2(blob, kZaC, alpha_deg);
  o.z[kZbC] = zscore2(blob, kZbC, beta_deg);
  o.z[kZaR] = zscore2(blob, kZaR, alpha_deg);
  o.z[kZbR] = zscore2(blob, kZbR, beta_deg);
  o.z[kZaLef2] = zscore2(blob, kZaLef2, alpha_deg);
}
__device__ __forceinline__ void zscores_el2(const float* __restrict__ blob, float2 el_deg, ZIn2& o) {
  o.z[kZeC] = zscore2(blob, kZeC, el_deg);
}

// y * std + mean, two roundings (hifi_F16_AeroData.py:36-37).
__device__ __forceinline__ float2 denorm2(const float* __restrict__ blob, int k, float2 y) {
  const float2 ms = reinterpret_cast<const float2*>(blob + kOnormOff)[k];
  return make_float2(y.x * ms.y + ms.x, y.y * ms.y + ms.x);
}

// Evaluate MLP nets [K0, K0 + count) of ONE ARCHITECTURE in a rolled loop; result k goes to out[k * stride] (a float2: both
// aircraft of this thread).  The alpha z-score is picked per net at run time (warp-uniform), so nets of the same shape but
// different input normalisation (the "r30 / a20" and the "lef" alpha grids) CAN share one body (-DNPLANE_MERGE_GROUPS: two
// bodies and 0.55 k SASS instructions fewer) -- measured 10 % slower end to end than one loop per (architecture,
// normalisation) pair (profiles/r02_variants.txt), which therefore stays the default.
// `first` / `count` select nets [K0 + first, K0 + first + count) of the group (the cooperative small-population kernel splits a
// group over the warps of a CTA).
template <int K0>
__device__ __forceinline__ void eval_group2(const float* __restrict__ blob, uint32_t wbase, const ZIn2& zi,
                                            float2* __restrict__ out, int stride, int count, int first = 0) {
  constexpr NetArch A = arch_of(K0);
  constexpr int NF = net_floats(A);
  static_assert(A.nin >= 2, "one-input nets are table-driven");
  uint32_t w = wbase + 4 * mlp_offset(K0) + 4 * NF * first;
#pragma unroll 1
  for (int k = K0 + first; k < K0 + first + count; ++k, w += 4 * NF) {
    float2 z0, z1, z2 = make_float2(0.f, 0.f);
    if constexpr (A.nin == 3) {
      z0 = zi.z[kZaC]; z1 = zi.z[kZbC]; z2 = zi.z[kZeC];
    } else {                                   // (alpha, beta) nets: beta is always the "r30" normalisation
      const bool r_grid = k == kCy || k == kdCl_a20 || (k >= kdCy_r30 && k <= kdCy_a20);
      z0 = r_grid ? zi.z[kZaR] : zi.z[kZaLef2];
      z1 = zi.z[kZbR];
    }
    const float2 y = mlp2<A.nin, A.h1, A.h2, A.h3>(w, z0, z1, z2);
    out[k * stride] = denorm2(blob, k, y);
  }
}
// every net the loops above may visit has the architecture and the input selection assumed there
constexpr bool ab2_groups_ok() {
  for (int k = kCy; k <= kdCl_a20_lef; ++k) {
    const ZSel z = zsel_of(k);
    const bool r_grid = k == kCy || k == kdCl_a20 || (k >= kdCy_r30 && k <= kdCy_a20);
    if (z.b != kZbR || z.a != (r_grid ? kZaR : kZaLef2)) return false;
  }
  for (int k = kCx; k <= kCl; ++k)
    if (zsel_of(k).a != kZaC || zsel_of(k).b != kZbC || zsel_of(k).e != kZeC) return false;
  const NetArch a = arch_of(kCy), b = arch_of(kdCz_lef);
  for (int k = kCy; k <= kdCl_lef; ++k)
    if (arch_of(k).h2 != a.h2 || arch_of(k).h3 != a.h3) return false;
  for (int k = kdCz_lef; k <= kdCn_a20; ++k)
    if (arch_of(k).h2 != b.h2 || arch_of(k).h3 != b.h3) return false;
  return true;
}
static_assert(ab2_groups_ok(), "net table changed: revisit eval_group2 / eval_ab2_nets");

// The 16 two-input (alpha, beta) nets: slots [kFirstAB2, kFirstA1), four architectures.
__device__ __forceinline__ void eval_ab2_nets(const float* __restrict__ blob, uint32_t wbase, const ZIn2& zi,
                                              float2* __restrict__ out, int stride) {
#ifdef NPLANE_MERGE_GROUPS
  eval_group2<kCy>(blob, wbase, zi, out, stride, 4);           // [20,10]:    Cy, delta_Cl_a20, delta_Cx_lef, delta_Cl_lef
  eval_group2<kdCz_lef>(blob, wbase, zi, out, stride, 8);      // [20,10,5]:  4 x lef, 3 x r30, delta_Cn_a20
#else
  eval_group2<kCy>(blob, wbase, zi, out, stride, 2);           // [20,10], r30 / a20 alpha grid: Cy, delta_Cl_a20
  eval_group2<kdCx_lef>(blob, wbase, zi, out, stride, 2);      // [20,10], lef alpha grid:       delta_Cx_lef, delta_Cl_lef
  eval_group2<kdCz_lef>(blob, wbase, zi, out, stride, 4);      // [20,10,5], lef
  eval_group2<kdCy_r30>(blob, wbase, zi, out, stride, 4);      // [20,10,5], r30 / a20
#endif
  eval_group2<kdCy_a20>(blob, wbase, zi, out, stride, 1);      // [20,10,10]
  eval_group2<kdCy_a20_lef>(blob, wbase, zi, out, stride, 3);  // [20,20,10]
}
// The three-input nets Cx Cz Cm Cn Cl (alpha, beta, el): `count` = 5 for a full nlplant, 2 (Cx, Cz) when only the
// force equations are needed (the Overload check).
__device__ __forceinline__ void eval_el3_nets(const float* __restrict__ blob, uint32_t wbase, const ZIn2& zi,
                                              float2* __restrict__ out, int stride, int count, int first = 0) {
  eval_group2<kCx>(blob, wbase, zi, out, stride, count, first);
}

// ------------------------------------------------------------------------------------------------
// One-input nets as exact piecewise-linear tables (aero_pack.h): binary search of the merged breakpoint list,
// one byte per net from the segment map, one LDS.128 (a0, y0, slope) + one FMA per net.
// ------------------------------------------------------------------------------------------------
struct AeroTabs {
  uint32_t bp_a, segmap, ent_a, bp_e, ent_e;  // shared-memory byte addresses
};
__device__ __forceinline__ AeroTabs aero_tabs(const void* blob_smem, uint32_t base) {
  const int32_t* h = reinterpret_cast<const int32_t*>(blob_smem);
  AeroTabs t;
  t.bp_a = base + 4u * (uint32_t)h[kHdrBpA];
  t.segmap = base + 4u * (uint32_t)h[kHdrSegmap];
  t.ent_a = base + 4u * (uint32_t)h[kHdrEntA];
  t.bp_e = base + 4u * (uint32_t)h[kHdrBpE];
  t.ent_e = base + 4u * (uint32_t)h[kHdrEntE];
  return t;
}
// Number of breakpoints <= x in a sorted, +inf padded list of 2^LEVELS - 1 floats (NaN -> 0), for the thread's two
// aircraft at once: two independent dependent-load chains in flight instead of one.
template <int LEVELS>
__device__ __forceinline__ void pwl_search2(uint32_t bp, float x0, float x1, uint32_t& p0, uint32_t& p1) {
  p0 = 0; p1 = 0;
#pragma unroll 1   // (a fully unrolled search measured 7 % slower end to end: profiles/r01_variants.txt)
  for (uint32_t step = 1u << (LEVELS - 1); step > 0; step >>= 1) {
    const float b0 = lds32f(bp + 4u * (p0 + step - 1u));
    const float b1 = lds32f(bp + 4u * (p1 + step - 1u));
    p0 += (b0 <= x0) ? step : 0u;
    p1 += (b1 <= x1) ? step : 0u;
  }
}
__device__ __forceinline__ float pwl_entry(uint32_t ent, uint32_t idx, float x) {
  const float4 e = lds128(ent + 16u * idx);
  return fmaf(e.z, x - e.x, e.y);
}
__device__ __forceinline__ float2 eta_el2(const AeroTabs& t, float2 el_deg) {
  uint32_t p0, p1;
  pwl_search2<kLevelsE>(t.bp_e, el_deg.x, el_deg.y, p0, p1);
  return make_float2(pwl_entry(t.ent_e, p0, el_deg.x), pwl_entry(t.ent_e, p1, el_deg.y));
}
// The alpha-only coefficients of one aircraft from its merged-segment index m: net k -> out[k - kFirstA1],
// k - kFirstA1 < COUNT.
template <int COUNT>
__device__ __forceinline__ void alpha_coefs(const void* blob_smem, const AeroTabs& t, uint32_t m, float alpha_deg,
                                            float* __restrict__ out) {
  const uint32_t row = t.segmap + (uint32_t)kSegmapRowBytes * m;
  const int32_t* taboff = reinterpret_cast<const int32_t*>(blob_smem) + kHdrTabOff;
#pragma unroll
  for (int w = 0; w < (COUNT + 3) / 4; ++w) {
    const uint32_t word = lds32u(row + 4u * w);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int k = 4 * w + b;
      if (k < COUNT) out[k] = pwl_entry(t.ent_a, (uint32_t)taboff[k] + ((word >> (8 * b)) & 0xFFu), alpha_deg);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Equations of motion
// ------------------------------------------------------------------------------------------------
// sincosf / powf / Philox bodies.  Sharing ONE copy of each per kernel (__noinline__: -DNPLANE_SHARED_LIBM) takes 1.5 k
// of the step kernel's SASS instructions away (ten inlined sincosf with their Payne-Hanek slow paths alone are 1.1 k) but
// MEASURED 10 % slower end to end (0.500 vs 0.451 ms per 10^6-aircraft step, same box, profiles/r02_variants.txt): a call
// serialises the two aircraft a thread interleaves.  Inlined is the default; the bits are the same either way.
#ifdef NPLANE_SHARED_LIBM
#define NP_LIBM_INLINE __noinline__
#else
#define NP_LIBM_INLINE __forceinline__
#endif
__device__ NP_LIBM_INLINE float2 sincos_shared(float x) {
  float2 r;
  sincosf(x, &r.x, &r.y);
  return r;
}
#ifdef NPLANE_SHARED_POW
__device__ __noinline__ float pow_shared(float x, float y) { return powf(x, y); }
#else
__device__ NP_LIBM_INLINE float pow_shared(float x, float y) { return powf(x, y); }
#endif

struct Trig {
  float sa, ca, sb, cb, st, ct, tt, sphi, cphi, spsi, cpsi;
};
__device__ __forceinline__ Trig make_trig(const float* s) {
  Trig t;
  float2 r;
  r = sincos_shared(s[7]); t.sa = r.x; t.ca = r.y;
  r = sincos_shared(s[8]); t.sb = r.x; t.cb = r.y;
  r = sincos_shared(s[4]); t.st = r.x; t.ct = r.y;
  t.tt = t.st / t.ct;  // tan(theta): one IEEE divide of the two values already needed (<= 3 ulp, like tanf's 4 ulp bound)
  r = sincos_shared(s[3]); t.sphi = r.x; t.cphi = r.y;
  r = sincos_shared(s[5]); t.spsi = r.x; t.cpsi = r.y;
  return t;
}

constexpr float kR2D = 57.29577951308232f;   // 180.0 / pi rounded once to f32
constexpr float kPi = 3.141592653589793f;
constexpr float kTwoPi = 6.283185307179586f;

// x / d for a compile-time constant d in three instructions instead of the ~10 (FCHK, MUFU.RCP, 4 FFMA, slow-path branch)
// of div.rn.f32:  q = RN(x * RN(1/d)) is within one ulp, the FMA residual e = x - q*d is exact, and RN(q + e * RN(1/d)) is
// the correctly rounded quotient (Markstein) -- the very bits of the reference's `tensor / constant`.  Verified by
// enumeration over EVERY float with 2^-100 <= |x| < 2^100 for every divisor used (oracle/divc_check.c; +-0 -> +0).
// Written `x / DC(340.0f)` so operator precedence at the call sites stays that of the reference expression.
struct DC {
  float d;
  __host__ __device__ constexpr explicit DC(float v) : d(v) {}
};
__device__ __forceinline__ float operator/(float x, DC c) {
  const float r = 1.0f / c.d;  // folded at compile time
  const float q = x * r;
  return fmaf(fmaf(-q, c.d, x), r, q);
}

__device__ __forceinline__ float tfac_pow(float alt) {  // tfac ** 4.14 (F16_dynamics.py:25,28)
  const float tfac = 1.0f - .703e-5f * alt;
  return pow_shared(tfac, 4.14f);
}
__device__ __forceinline__ float qbar_of(float tp, float vt) {  // .5 * rho0 * tfac^4.14 * vt^2 (:28,30)
  const float rho = 2.377e-3f * tp;
  return .5f * rho * (vt * vt);
}
__device__ __forceinline__ float eas2tas_of(float tp) {  // F16_model.py:156-162
  return sqrtf(1.0f / tp);
}

// fmodf(a, 2 pi) in ~10 instructions instead of ~70 (and 13 inlined copies of them): q = trunc(a * RN(1 / 2pi)) is the true
// truncated quotient or off by one (|a / 2pi| < 2^20: the product's error is < 0.25); r = fma(-q, 2pi, a) is EXACT when q is
// right (the fmod result is always representable) and has the wrong sign / magnitude >= 2pi when q is off by one, in which
// case q is corrected and r recomputed from scratch.  Bit-identical to fmodf: enumerated over every float of that range in
// oracle/fmod_check.c.  Larger arguments (no flying aircraft has them) take the library routine.
__device__ __noinline__ float fmod_twopi_slow(float a) { return fmodf(a, kTwoPi); }
__device__ __forceinline__ float fmod_twopi(float a) {
  if (!(fabsf(a) < 6.0e6f)) return fmod_twopi_slow(a);   // also NaN / inf
  float q = truncf(a * (float)(1.0 / 6.283185307179586));
  float r = fmaf(-q, kTwoPi, a);
  const float sgn = copysignf(1.0f, a);
  const float rs = r * sgn;                // remainder folded to a's sign: must lie in [0, 2pi)
  if (rs < 0.0f) q -= sgn;
  else if (rs >= kTwoPi) q += sgn;
  else return r == 0.0f ? copysignf(0.0f, a) : r;     // fmod keeps the dividend's sign on an exact zero
  r = fmaf(-q, kTwoPi, a);
  return r == 0.0f ? copysignf(0.0f, a) : r;
}
// torch `%` (Python-style remainder) then the two wraps of utils.py:144-154.
__device__ __forceinline__ float wrap_pi(float a) {
  float r = fmod_twopi(a);
  if (r != 0.0f && r < 0.0f) r += kTwoPi;
  if (r < 0.0f) r += kTwoPi;
  if (r > kPi) r -= kTwoPi;
  return r;
}

// Force (body-axis velocity derivative) part of nlplant: returns Vt_dot, alpha_dot, beta_dot
// (F16_dynamics.py:197-207,215-220).  `c` indexes MLP coefficient k (< kNumSlots) at c[k * cs]; `a1` holds the
// alpha-only (table-driven) coefficients, net k at a1[k - kFirstA1].
struct ForceOut {
  float vt_dot, alpha_dot, beta_dot;
};
struct AeroTotals {
  float Cx, Cy, Cz;
};

__device__ __forceinline__ AeroTotals force_totals(const float* __restrict__ c, int cs, const float* __restrict__ a1, float vt, float P, float Q,
                                                   float R, float dail, float drud, float dlef) {
  constexpr float cbar = 11.32f, B = 30.0f;
  const float k2 = cbar / (2.0f * vt);
  const float b2 = B / (2.0f * vt);
  AeroTotals t;
  const float dXdQ = k2 * (a1[kCxq - kFirstA1] + a1[kdCxq_lef - kFirstA1] * dlef);
  t.Cx = c[kCx * cs] + c[kdCx_lef * cs] * dlef + dXdQ * Q;
  const float dZdQ = k2 * (a1[kCzq - kFirstA1] + c[kdCz_lef * cs] * dlef);  // sic (F16_dynamics.py:199)
  t.Cz = c[kCz * cs] + c[kdCz_lef * cs] * dlef + dZdQ * Q;
  const float dYdail = c[kdCy_a20 * cs] + c[kdCy_a20_lef * cs] * dlef;
  const float dYdR = b2 * (a1[kCyr - kFirstA1] + a1[kdCyr_lef - kFirstA1] * dlef);
  const float dYdP = b2 * (a1[kCyp - kFirstA1] + a1[kdCyp_lef - kFirstA1] * dlef);
  t.Cy = c[kCy * cs] + c[kdCy_lef * cs] * dlef + dYdail * dail + c[kdCy_r30 * cs] * drud + dYdR * R + dYdP * P;
  return t;
}

struct BodyVel {
  float U, V, W;
};
__device__ __forceinline__ BodyVel body_vel(float vt, const Trig& g) {
  BodyVel b;
  b.U = vt * g.ca * g.cb;
  b.V = vt * g.sb;
  b.W = vt * g.sa * g.cb;
  return b;
}

__device__ __forceinline__ ForceOut force_eqs(const AeroTotals& t, const BodyVel& b, const Trig& g, float vt, float P,
                                              float Q, float R, float qbar, float T) {
  constexpr float grav = 32.17f, m = 636.94f, S = 300.0f;
  const float Udot = R * b.V - Q * b.W - grav * g.st + qbar * S * t.Cx / m + T / m;
  const float Vdot = P * b.W - R * b.U + grav * g.ct * g.sphi + qbar * S * t.Cy / m;
  const float Wdot = Q * b.U - P * b.V + grav * g.ct * g.cphi + qbar * S * t.Cz / m;
  ForceOut o;
  o.vt_dot = (b.U * Udot + b.V * Vdot + b.W * Wdot) / vt;
  o.alpha_dot = (b.U * Wdot - b.W * Udot) / (b.U * b.U + b.W * b.W);
  o.beta_dot = (Vdot * vt - b.V * o.vt_dot) / (vt * vt * g.cb);
  return o;
}

// The part of nlplant beyond the force equations: navigation + Euler-angle kinematics (xdot[0..5], :133-138) and the
// moment equations (xdot[9..11], :208-214,221-227), given the force totals `t` of the same (s, u).
__device__ __forceinline__ void nlplant_kin_moments(const float* s, float ail, float rud, float lef, const Trig& g,
                                                    float qbar, float vt, const BodyVel& b, const AeroTotals& t,
                                                    const float* __restrict__ c, int cs, const float* __restrict__ a1,
                                                    float* xdot) {
  constexpr float B = 30.0f, S = 300.0f, cbar = 11.32f;
  constexpr float Jy = 55814.0f, Jxz = 982.0f, Jz = 63100.0f, Jx = 9496.0f;
  constexpr float xcg_arm = (float)(0.35 - 0.30);          // (xcgr - xcg) evaluated in double, then f32
  constexpr float cbar_over_B = (float)(11.32 / 30.0);
  const float P = s[9], Q = s[10], R = s[11];
  const float beta_deg = s[8] * kR2D;
  const float dail = ail / 21.5f;
  const float drud = rud / 30.0f;
  const float dlef = 1.0f - lef / 25.0f;

  // navigation + Euler-angle kinematics (:133-138)
  xdot[0] = b.U * (g.ct * g.cpsi) + b.V * (g.sphi * g.cpsi * g.st - g.cphi * g.spsi) +
            b.W * (g.cphi * g.st * g.cpsi + g.sphi * g.spsi);
  xdot[1] = b.U * (g.ct * g.spsi) + b.V * (g.sphi * g.spsi * g.st + g.cphi * g.cpsi) +
            b.W * (g.cphi * g.st * g.spsi - g.sphi * g.cpsi);
  xdot[2] = b.U * g.st - b.V * (g.sphi * g.ct) - b.W * (g.cphi * g.ct);
  xdot[3] = P + g.tt * (Q * g.sphi + R * g.cphi);
  xdot[4] = Q * g.cphi - R * g.sphi;
  xdot[5] = (Q * g.sphi + R * g.cphi) / g.ct;

  // aero build-up (:208-214)
  const float k2 = cbar / (2.0f * vt);
  const float b2 = B / (2.0f * vt);
  const float dMdQ = k2 * (a1[kCmq - kFirstA1] + a1[kdCmq_lef - kFirstA1] * dlef);
  const float Cm_tot = c[kCm * cs] * c[kEtaEl * cs] + t.Cz * xcg_arm + c[kdCm_lef * cs] * dlef + dMdQ * Q +
                       a1[kdCm - kFirstA1];  // + delta_Cm_ds (== 0, hifi_F16_AeroData.py:811-818)
  const float dNdail = c[kdCn_a20 * cs] + c[kdCn_a20_lef * cs] * dlef;
  const float dNdR = b2 * (a1[kCnr - kFirstA1] + a1[kdCnr_lef - kFirstA1] * dlef);
  const float dNdP = b2 * (a1[kCnp - kFirstA1] + a1[kdCnp_lef - kFirstA1] * dlef);
  const float Cn_tot = c[kCn * cs] + c[kdCn_lef * cs] * dlef - t.Cy * xcg_arm * cbar_over_B + dNdail * dail +
                       c[kdCn_r30 * cs] * drud + dNdR * R + dNdP * P + a1[kdCnbeta - kFirstA1] * beta_deg;
  const float dLdail = c[kdCl_a20 * cs] + c[kdCl_a20_lef * cs] * dlef;
  const float dLdR = b2 * (a1[kClr - kFirstA1] + a1[kdClr_lef - kFirstA1] * dlef);
  const float dLdP = b2 * (a1[kClp - kFirstA1] + a1[kdClp_lef - kFirstA1] * dlef);
  const float Cl_tot = c[kCl * cs] + c[kdCl_lef * cs] * dlef + dLdail * dail + c[kdCl_r30 * cs] * drud + dLdR * R +
                       dLdP * P + a1[kdClbeta - kFirstA1] * beta_deg;

  // moments (:221-227); the Heng (= 0) terms add exact zeros and are omitted.
  const float L_tot = Cl_tot * qbar * S * B;
  const float M_tot = Cm_tot * qbar * S * cbar;
  const float N_tot = Cn_tot * qbar * S * B;
  constexpr float denom = (float)(9496.0 * 63100.0 - 982.0 * 982.0);
  constexpr float kQR = (float)(63100.0 * (63100.0 - 55814.0) + 982.0 * 982.0);
  constexpr float kPQ = (float)(982.0 * (9496.0 - 55814.0 + 63100.0));
  constexpr float kPQ2 = (float)(9496.0 * (9496.0 - 55814.0) + 982.0 * 982.0);
  constexpr float kJzJx = (float)(63100.0 - 9496.0);
  xdot[9] = (Jz * L_tot + Jxz * N_tot - kQR * Q * R + kPQ * P * Q) / denom;
  xdot[10] = (M_tot + kJzJx * P * R - Jxz * (P * P - R * R)) / Jy;
  xdot[11] = (Jx * N_tot + Jxz * L_tot + kPQ2 * P * Q - kPQ * Q * R) / denom;
}

// The force part shared by the Euler derivative and the Overload check: clamps vt (:104), builds the force
// totals and solves the force equations.  Outputs the pieces nlplant_kin_moments needs.
struct ForcePart {
  float vt, qbar;
  BodyVel b;
  AeroTotals t;
  ForceOut f;
};
__device__ __forceinline__ ForcePart force_part(const float* s, float T, float ail, float rud, float lef, const Trig& g,
                                                float tp, const float* __restrict__ c, int cs,
                                                const float* __restrict__ a1) {
  ForcePart o;
  o.vt = s[6] <= 0.01f ? 0.01f : s[6];           // :104
  o.qbar = qbar_of(tp, o.vt);
  o.b = body_vel(o.vt, g);
  o.t = force_totals(c, cs, a1, o.vt, s[9], s[10], s[11], ail / 21.5f, rud / 30.0f, 1.0f - lef / 25.0f);
  o.f = force_eqs(o.t, o.b, g, o.vt, s[9], s[10], s[11], o.qbar, T);
  return o;
}

// Full nlplant: xdot[0..11] from s[0..11], controls (T, el, ail, rud, lef) and the coefficients, which must already
// be evaluated at (alpha_deg, beta_deg, el) of this (s, u).
__device__ __forceinline__ void nlplant_from_coefs(const float* s, float T, float ail, float rud, float lef,
                                                   const Trig& g, float tp, const float* __restrict__ c, int cs,
                                                   const float* __restrict__ a1, float* xdot) {
  const ForcePart fp = force_part(s, T, ail, rud, lef, g, tp, c, cs, a1);
  nlplant_kin_moments(s, ail, rud, lef, g, fp.qbar, fp.vt, fp.b, fp.t, c, cs, a1, xdot);
  xdot[6] = fp.f.vt_dot;
  xdot[7] = fp.f.alpha_dot;
  xdot[8] = fp.f.beta_dot;
}

// Body-axis accelerations of F16Model.get_acceleration (F16_model.py:132-148) from Vt_dot/alpha_dot/beta_dot.
__device__ __forceinline__ void body_accel(const float* s, const Trig& g, const ForceOut& f, float& ax, float& ay,
                                           float& az) {
  const float vt = s[6];  // raw vt (the getter does not clamp)
  const float vel_u = vt * g.cb * g.ca;
  const float vel_v = vt * g.sb;
  const float vel_w = vt * g.cb * g.sa;
  const float u_dot = g.cb * g.ca * f.vt_dot - vt * g.sb * g.ca * f.beta_dot - vt * g.cb * g.sa * f.alpha_dot;
  const float v_dot = g.sb * f.vt_dot + vt * g.cb * f.beta_dot;
  const float w_dot = g.cb * g.sa * f.vt_dot - vt * g.sb * g.sa * f.beta_dot + vt * g.cb * g.ca * f.alpha_dot;
  ax = u_dot + s[10] * vel_w - s[11] * vel_v;
  ay = v_dot + s[11] * vel_u - s[9] * vel_w;
  az = w_dot + s[9] * vel_v - s[10] * vel_u;
}

// ------------------------------------------------------------------------------------------------
// Counter-based RNG: Philox4x32-10 keyed by the env seed; counter = (global aircraft index, step, stream).
// ------------------------------------------------------------------------------------------------
#ifdef NPLANE_SHARED_PHILOX
__device__ __noinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#else
__device__ NP_LIBM_INLINE uint4 philox4x32_10(uint4 ctr, uint2 key) {
#endif
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }  // [0,1)
// Two normals of standard deviation `sc` from ONE 32-bit word (observation noise only): 16-bit radius and angle
// uniforms taken at bin midpoints, MUFU log2 / sqrt / sin / cos.  |n| <= 4.8 sc; mean 0, variance sc^2 to < 1e-4.
// r = sc * sqrt(-2 ln u1) is returned with the two direction cosines so the caller adds the noise with one FMA each.
__device__ __forceinline__ void box_muller16(uint32_t w, float sc, float& r, float& cs, float& sn) {
  const float u1 = fmaf((float)(w >> 16), 1.52587890625e-5f, 0.5f * 1.52587890625e-5f);     // (0,1)
  const float th = fmaf((float)(w & 0xFFFFu), kTwoPi * 1.52587890625e-5f, 0.5f * kTwoPi * 1.52587890625e-5f);
  float l2, rt;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));                          // u1 >= 2^-17: never denormal
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(l2 * -1.3862943611198906f));  // -2 ln 2 * log2 u1 = -2 ln u1
  r = rt * sc;
  __sincosf(th, &sn, &cs);
}

}  // namespace npl
