// Device-side building blocks of the F-16 step: the 43 MLP aero surrogates, the 6-DoF equations of
// motion, atmosphere, observation helpers and the counter-based RNG.
//
// Numerics contract: this translation unit is compiled with -fmad=false, so every `a*b+c` written
// below rounds twice exactly like the reference's eager PyTorch ops do; the MLP inner products use
// explicit fmaf() (the reference's nn.Linear goes through an FMA-based sgemm with unspecified
// summation order, so no order is "the" reference order there).  Division and sqrt are IEEE
// (nvcc defaults -prec-div/-prec-sqrt=true); sinf/cosf/tanf/powf are the accurate CUDA libm versions.
//
// Reference formulas: envs/models/F16/F16_dynamics.py:22-35 (atmos), :37-229 (nlplant),
// envs/models/F16/hifi_F16_AeroData.py:12-37,149-166,748-819 (MLPs), envs/models/F16_model.py:132-162
// (accelerations, EAS2TAS), envs/utils/utils.py:144-154 (wrap_PI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace npl {

// ------------------------------------------------------------------------------------------------
// Net table: canonical order of neuralplane_b200/data/f16_aero.npz (tools/pack_f16_aero.py).
// ------------------------------------------------------------------------------------------------
struct NetArch {
  int nin, h1, h2, h3;  // h3 == 0: two hidden layers
};

constexpr int kNumNets = 43;
constexpr int kNumUsed = 42;  // net 42 (delta_Czq_lef) is never consumed (F16_dynamics.py:167-175)

// coefficient slots == net indices
enum Coef : int {
  kCx = 0, kCz, kCm, kCn, kCl, kEtaEl,                                  // el-dependent (G1, G2)
  kCy, kdCl_a20, kdCx_lef, kdCl_lef,                                     // (a,b) [20,10]       (G3, G4)
  kdCz_lef, kdCm_lef, kdCy_lef, kdCn_lef,                                // (a,b) [20,10,5] lef (G5)
  kdCy_r30, kdCn_r30, kdCl_r30, kdCn_a20,                                // (a,b) [20,10,5] r   (G6)
  kdCy_a20,                                                              // (a,b) [20,10,10]    (G7)
  kdCy_a20_lef, kdCn_a20_lef, kdCl_a20_lef,                              // (a,b) [20,20,10]    (G8)
  kCxq, kCzq, kCmq, kCyp, kCyr, kCnr, kCnp, kClp, kClr,                  // (a) ALPHA1          (G9)
  kdCnbeta, kdClbeta, kdCm,
  kdCxq_lef, kdCyr_lef, kdClr_lef, kdClp_lef, kdCmq_lef, kdCnr_lef, kdCnp_lef,  // (a) lef   (G10)
  kdCyp_lef,                                                             // (a) [20,10,5] lef   (G11)
  kdCzq_lef                                                              // unused
};
constexpr int kFirstAB = kCy;               // nets [kFirstAB, kNumUsed) depend on (alpha, beta) only
constexpr int kNumAB = kNumUsed - kFirstAB;  // 36

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
// (kZbR is also the beta normalisation of the lef-2D nets.)
struct ZSel { int a, b, e; };
constexpr ZSel zsel_of(int k) {
  return k <= kCl ? ZSel{kZaC, kZbC, kZeC}
       : k == kEtaEl ? ZSel{-1, -1, kZeEta}
       : (k == kCy || k == kdCl_a20 || (k >= kdCy_r30 && k <= kdCy_a20)) ? ZSel{kZaR, kZbR, -1}
       : (k <= kdCl_a20_lef) ? ZSel{kZaLef2, kZbR, -1}
       : (k <= kdCm) ? ZSel{kZaA1, -1, -1}
       : ZSel{kZaLef1, -1, -1};
}

// Device blob layout (floats): [znorm: kNumZ x {mean, std}] [onorm: 43 x {mean, std}] [weights...]
// per net, per layer: bias[out] then W^T[in][out] (input-major), the layer padded to a multiple of 4 floats
// so every layer starts 16-byte aligned for LDS.128.
constexpr int pad4(int x) { return (x + 3) & ~3; }
constexpr int layer_floats(int in, int out) { return pad4(out + in * out); }
constexpr int net_floats(NetArch a) {
  return layer_floats(a.nin, a.h1) + layer_floats(a.h1, a.h2) +
         (a.h3 ? layer_floats(a.h2, a.h3) + layer_floats(a.h3, 1) : layer_floats(a.h2, 1));
}
constexpr int kZnormOff = 0;
constexpr int kOnormOff = pad4(2 * kNumZ);
constexpr int kWeightOff = kOnormOff + pad4(2 * kNumNets);
constexpr int net_offset(int k) {
  int off = kWeightOff;
  for (int i = 0; i < k; ++i) off += net_floats(arch_of(i));
  return off;
}
constexpr int kBlobFloats = net_offset(kNumNets);
constexpr int kBlobBytes = kBlobFloats * 4;
static_assert(kBlobBytes % 16 == 0, "blob must be a whole number of 16-byte chunks for cp.async.bulk");

// ------------------------------------------------------------------------------------------------
// MLP evaluation: one aircraft per thread, weights broadcast from shared memory with LDS.128.
// ------------------------------------------------------------------------------------------------
// Weights are read with explicit ld.shared.v4 (LDS.128, warp-broadcast).  Besides pinning the access width,
// the non-volatile asm keeps several thousand weight loads out of the compiler's alias analysis (which otherwise
// dominates compile time); ordering against the one-time TMA staging is carried by the data dependence on
// `wbase`, which callers obtain from aero_base_after_staging() after the staging barrier.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 q;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(addr));
  return q;
}
// Re-materialise an address through a volatile mov: loads keyed on the result can be neither hoisted out of the
// enclosing loop nor merged with an earlier evaluation's loads (which would pin hundreds of registers).
__device__ __forceinline__ uint32_t opaque_u32(uint32_t x) {
  uint32_t y;
  asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t aero_base_after_staging(const float* blob_smem) {
  uint32_t a;
  asm volatile("mov.u32 %0, %1;" : "=r"(a) : "r"((uint32_t)__cvta_generic_to_shared(blob_smem)) : "memory");
  return a;
}

template <int IN, int OUT, bool RELU>
__device__ __forceinline__ void dense(uint32_t w, const float (&x)[IN], float (&y)[OUT]) {
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
        y[f] = e[k];
      } else if (f < NF) {
        const int g = f - OUT;
        y[g % OUT] = fmaf(x[g / OUT], e[k], y[g % OUT]);
      }
    }
  }
  if (RELU) {
#pragma unroll
    for (int j = 0; j < OUT; ++j) y[j] = fmaxf(y[j], 0.0f);
  }
}

template <int NIN, int H1, int H2, int H3>
__device__ __forceinline__ float mlp(uint32_t w, float z0, float z1, float z2) {
  float x[NIN];
  x[0] = z0;
  if (NIN > 1) x[NIN > 1 ? 1 : 0] = z1;
  if (NIN > 2) x[NIN > 2 ? 2 : 0] = z2;
  float a[H1];
  dense<NIN, H1, true>(w, x, a);
  w += 4 * layer_floats(NIN, H1);
  float b[H2];
  dense<H1, H2, true>(w, a, b);
  w += 4 * layer_floats(H1, H2);
  float y[1];
  if constexpr (H3 > 0) {
    float c[H3];
    dense<H2, H3, true>(w, b, c);
    w += 4 * layer_floats(H2, H3);
    dense<H3, 1, false>(w, c, y);
  } else {
    dense<H2, 1, false>(w, b, y);
  }
  return y[0];
}

// z-scores of (alpha_deg, beta_deg, el_deg) for every normalisation group: (x - mean) / std with a true
// divide (hifi_F16_AeroData.py:32-33).
struct ZIn {
  float z[kNumZ];
};
__device__ __forceinline__ void zscores_ab(const float* __restrict__ blob, float alpha_deg, float beta_deg, ZIn& o) {
  const float* zn = blob + kZnormOff;
  o.z[kZaC] = (alpha_deg - zn[2 * kZaC]) / zn[2 * kZaC + 1];
  o.z[kZbC] = (beta_deg - zn[2 * kZbC]) / zn[2 * kZbC + 1];
  o.z[kZaR] = (alpha_deg - zn[2 * kZaR]) / zn[2 * kZaR + 1];
  o.z[kZbR] = (beta_deg - zn[2 * kZbR]) / zn[2 * kZbR + 1];
  o.z[kZaLef2] = (alpha_deg - zn[2 * kZaLef2]) / zn[2 * kZaLef2 + 1];
  o.z[kZaA1] = (alpha_deg - zn[2 * kZaA1]) / zn[2 * kZaA1 + 1];
  o.z[kZaLef1] = (alpha_deg - zn[2 * kZaLef1]) / zn[2 * kZaLef1 + 1];
}
__device__ __forceinline__ void zscores_el(const float* __restrict__ blob, float el_deg, ZIn& o) {
  const float* zn = blob + kZnormOff;
  o.z[kZeC] = (el_deg - zn[2 * kZeC]) / zn[2 * kZeC + 1];
  o.z[kZeEta] = (el_deg - zn[2 * kZeEta]) / zn[2 * kZeEta + 1];
}

// y * std + mean, two roundings (hifi_F16_AeroData.py:36-37).
__device__ __forceinline__ float denorm(const float* __restrict__ blob, int k, float y) {
  const float2 ms = reinterpret_cast<const float2*>(blob + kOnormOff)[k];
  return y * ms.y + ms.x;
}

// Evaluate nets [K0, K1) of one architecture/normalisation group in a rolled loop; result k goes to
// out[k * stride].  The loop keeps the code footprint at one body per group.
template <int K0, int K1>
__device__ __forceinline__ void eval_group(const float* __restrict__ blob, uint32_t wbase, const ZIn& zi,
                                           float* __restrict__ out, int stride) {
  constexpr NetArch A = arch_of(K0);
  constexpr ZSel Z = zsel_of(K0);
  constexpr int NF = net_floats(A);
  float z0, z1 = 0.f, z2 = 0.f;
  if constexpr (A.nin == 3) {
    z0 = zi.z[Z.a]; z1 = zi.z[Z.b]; z2 = zi.z[Z.e];
  } else if constexpr (A.nin == 2) {
    z0 = zi.z[Z.a]; z1 = zi.z[Z.b];
  } else {
    z0 = Z.a >= 0 ? zi.z[Z.a >= 0 ? Z.a : 0] : zi.z[Z.e >= 0 ? Z.e : 0];
  }
  uint32_t w = wbase + 4 * net_offset(K0);
#pragma unroll 1
  for (int k = K0; k < K1; ++k, w += 4 * NF) {
    const float y = mlp<A.nin, A.h1, A.h2, A.h3>(w, z0, z1, z2);
    out[k * stride] = denorm(blob, k, y);
  }
}

// The (alpha, beta)-only nets: slots [kFirstAB, kNumUsed).
__device__ __forceinline__ void eval_ab_nets(const float* __restrict__ blob, uint32_t wbase, const ZIn& zi, float* __restrict__ out,
                                             int stride) {
  eval_group<kCy, kdCl_a20 + 1>(blob, wbase, zi, out, stride);
  eval_group<kdCx_lef, kdCl_lef + 1>(blob, wbase, zi, out, stride);
  eval_group<kdCz_lef, kdCn_lef + 1>(blob, wbase, zi, out, stride);
  eval_group<kdCy_r30, kdCn_a20 + 1>(blob, wbase, zi, out, stride);
  eval_group<kdCy_a20, kdCy_a20 + 1>(blob, wbase, zi, out, stride);
  eval_group<kdCy_a20_lef, kdCl_a20_lef + 1>(blob, wbase, zi, out, stride);
  eval_group<kCxq, kdCm + 1>(blob, wbase, zi, out, stride);
  eval_group<kdCxq_lef, kdCnp_lef + 1>(blob, wbase, zi, out, stride);
  eval_group<kdCyp_lef, kdCyp_lef + 1>(blob, wbase, zi, out, stride);
}
// The elevator-dependent nets: Cx Cz Cm Cn Cl (alpha, beta, el) and eta_el (el).
__device__ __forceinline__ void eval_el_nets(const float* __restrict__ blob, uint32_t wbase, const ZIn& zi, float* __restrict__ out,
                                             int stride) {
  eval_group<kCx, kCl + 1>(blob, wbase, zi, out, stride);
  eval_group<kEtaEl, kEtaEl + 1>(blob, wbase, zi, out, stride);
}
// Only Cx and Cz: all the force equations (and so the Overload check) need from the el-dependent nets.
__device__ __forceinline__ void eval_el_force_nets(const float* __restrict__ blob, uint32_t wbase, const ZIn& zi,
                                                   float* __restrict__ out, int stride) {
  eval_group<kCx, kCz + 1>(blob, wbase, zi, out, stride);
}

// ------------------------------------------------------------------------------------------------
// Equations of motion
// ------------------------------------------------------------------------------------------------
struct Trig {
  float sa, ca, sb, cb, st, ct, tt, sphi, cphi, spsi, cpsi;
};
__device__ __forceinline__ Trig make_trig(const float* s) {
  Trig t;
  sincosf(s[7], &t.sa, &t.ca);
  sincosf(s[8], &t.sb, &t.cb);
  sincosf(s[4], &t.st, &t.ct);
  t.tt = tanf(s[4]);
  sincosf(s[3], &t.sphi, &t.cphi);
  sincosf(s[5], &t.spsi, &t.cpsi);
  return t;
}

constexpr float kR2D = 57.29577951308232f;   // 180.0 / pi rounded once to f32
constexpr float kPi = 3.141592653589793f;
constexpr float kTwoPi = 6.283185307179586f;

__device__ __forceinline__ float tfac_pow(float alt) {  // tfac ** 4.14 (F16_dynamics.py:25,28)
  const float tfac = 1.0f - .703e-5f * alt;
  return powf(tfac, 4.14f);
}
__device__ __forceinline__ float qbar_of(float tp, float vt) {  // .5 * rho0 * tfac^4.14 * vt^2 (:28,30)
  const float rho = 2.377e-3f * tp;
  return .5f * rho * (vt * vt);
}
__device__ __forceinline__ float eas2tas_of(float tp) {  // F16_model.py:156-162
  return sqrtf(1.0f / tp);
}

// torch `%` (Python-style remainder) then the two wraps of utils.py:144-154.
__device__ __forceinline__ float wrap_pi(float a) {
  float r = fmodf(a, kTwoPi);
  if (r != 0.0f && r < 0.0f) r += kTwoPi;
  if (r < 0.0f) r += kTwoPi;
  if (r > kPi) r -= kTwoPi;
  return r;
}

// Force (body-axis velocity derivative) part of nlplant: returns Vt_dot, alpha_dot, beta_dot
// (F16_dynamics.py:197-207,215-220).  `c` indexes coefficient k at c[k * cs].
struct ForceOut {
  float vt_dot, alpha_dot, beta_dot;
};
struct AeroTotals {
  float Cx, Cy, Cz;
};

__device__ __forceinline__ AeroTotals force_totals(const float* __restrict__ c, int cs, float vt, float P, float Q,
                                                   float R, float dail, float drud, float dlef) {
  constexpr float cbar = 11.32f, B = 30.0f;
  const float k2 = cbar / (2.0f * vt);
  const float b2 = B / (2.0f * vt);
  AeroTotals t;
  const float dXdQ = k2 * (c[kCxq * cs] + c[kdCxq_lef * cs] * dlef);
  t.Cx = c[kCx * cs] + c[kdCx_lef * cs] * dlef + dXdQ * Q;
  const float dZdQ = k2 * (c[kCzq * cs] + c[kdCz_lef * cs] * dlef);  // sic (F16_dynamics.py:199)
  t.Cz = c[kCz * cs] + c[kdCz_lef * cs] * dlef + dZdQ * Q;
  const float dYdail = c[kdCy_a20 * cs] + c[kdCy_a20_lef * cs] * dlef;
  const float dYdR = b2 * (c[kCyr * cs] + c[kdCyr_lef * cs] * dlef);
  const float dYdP = b2 * (c[kCyp * cs] + c[kdCyp_lef * cs] * dlef);
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

// Full nlplant: xdot[0..11] from s[0..11], controls (T, el, ail, rud, lef) and the 42 coefficients.
// Coefficients must already be evaluated at (alpha_deg, beta_deg, el) of this (s, u).
__device__ __forceinline__ void nlplant_from_coefs(const float* s, float T, float ail, float rud, float lef,
                                                   const Trig& g, float tp, const float* __restrict__ c, int cs,
                                                   float* xdot) {
  constexpr float B = 30.0f, S = 300.0f, cbar = 11.32f;
  constexpr float Jy = 55814.0f, Jxz = 982.0f, Jz = 63100.0f, Jx = 9496.0f;
  constexpr float xcg_arm = (float)(0.35 - 0.30);          // (xcgr - xcg) evaluated in double, then f32
  constexpr float cbar_over_B = (float)(11.32 / 30.0);
  const float P = s[9], Q = s[10], R = s[11];
  const float beta_deg = s[8] * kR2D;
  const float vt = s[6] <= 0.01f ? 0.01f : s[6];           // :104
  const float dail = ail / 21.5f;
  const float drud = rud / 30.0f;
  const float dlef = 1.0f - lef / 25.0f;
  const float qbar = qbar_of(tp, vt);
  const BodyVel b = body_vel(vt, g);

  // navigation + Euler-angle kinematics (:133-138)
  xdot[0] = b.U * (g.ct * g.cpsi) + b.V * (g.sphi * g.cpsi * g.st - g.cphi * g.spsi) +
            b.W * (g.cphi * g.st * g.cpsi + g.sphi * g.spsi);
  xdot[1] = b.U * (g.ct * g.spsi) + b.V * (g.sphi * g.spsi * g.st + g.cphi * g.cpsi) +
            b.W * (g.cphi * g.st * g.spsi - g.sphi * g.cpsi);
  xdot[2] = b.U * g.st - b.V * (g.sphi * g.ct) - b.W * (g.cphi * g.ct);
  xdot[3] = P + g.tt * (Q * g.sphi + R * g.cphi);
  xdot[4] = Q * g.cphi - R * g.sphi;
  xdot[5] = (Q * g.sphi + R * g.cphi) / g.ct;

  // aero build-up (:197-214)
  const AeroTotals t = force_totals(c, cs, vt, P, Q, R, dail, drud, dlef);
  const float k2 = cbar / (2.0f * vt);
  const float b2 = B / (2.0f * vt);
  const float dMdQ = k2 * (c[kCmq * cs] + c[kdCmq_lef * cs] * dlef);
  const float Cm_tot = c[kCm * cs] * c[kEtaEl * cs] + t.Cz * xcg_arm + c[kdCm_lef * cs] * dlef + dMdQ * Q +
                       c[kdCm * cs];  // + delta_Cm_ds (== 0, hifi_F16_AeroData.py:811-818)
  const float dNdail = c[kdCn_a20 * cs] + c[kdCn_a20_lef * cs] * dlef;
  const float dNdR = b2 * (c[kCnr * cs] + c[kdCnr_lef * cs] * dlef);
  const float dNdP = b2 * (c[kCnp * cs] + c[kdCnp_lef * cs] * dlef);
  const float Cn_tot = c[kCn * cs] + c[kdCn_lef * cs] * dlef - t.Cy * xcg_arm * cbar_over_B + dNdail * dail +
                       c[kdCn_r30 * cs] * drud + dNdR * R + dNdP * P + c[kdCnbeta * cs] * beta_deg;
  const float dLdail = c[kdCl_a20 * cs] + c[kdCl_a20_lef * cs] * dlef;
  const float dLdR = b2 * (c[kClr * cs] + c[kdClr_lef * cs] * dlef);
  const float dLdP = b2 * (c[kClp * cs] + c[kdClp_lef * cs] * dlef);
  const float Cl_tot = c[kCl * cs] + c[kdCl_lef * cs] * dlef + dLdail * dail + c[kdCl_r30 * cs] * drud + dLdR * R +
                       dLdP * P + c[kdClbeta * cs] * beta_deg;

  const ForceOut f = force_eqs(t, b, g, vt, P, Q, R, qbar, T);
  xdot[6] = f.vt_dot;
  xdot[7] = f.alpha_dot;
  xdot[8] = f.beta_dot;

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
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
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
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = ((float)(a >> 8) + 1.0f) * 5.9604644775390625e-8f;  // (0,1]
  const float r = sqrtf(-2.0f * __logf(u1));
  float sn, cs;
  __sincosf(kTwoPi * u01(b), &sn, &cs);
  n0 = r * cs;
  n1 = r * sn;
}

}  // namespace npl
