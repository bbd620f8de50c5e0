// Kernels off the step's hot path: the device-resident rollout buffer (K7), the table coefficient kernels (K6), the stand-alone
// nlplant / coefficient / model.update() kernels behind the plug-in getters, and the reset-constants kernel of np_aero_create.
#pragma once
#include "env_device.cuh"
#include "ptx_device.cuh"
#include "tables_device.cuh"

// ------------------------------------------------------------------------------------------------
// K7: device-resident rollout buffer (SURVEY f-2).  ReplayBuffer.compute_returns (algorithms/utils/buffer.py:139-172) as a
// backward scan, one thread per (env, agent) column -- rows are [T(+1)][M] so every load / store is coalesced; and the
// mask derivation of F16SimRunner.insert (runner/F16sim_runner.py:141-157) straight from the env's flag rows.
// Arithmetic order == numpy's fp32 evaluation of the reference expressions (built with -fmad=false): bit-exact.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rollout_returns_kernel(const float* __restrict__ R, float* __restrict__ V,
                                                              const float* __restrict__ Mk, const float* __restrict__ Bm,
                                                              float* __restrict__ Ret, int T, int M, float gamma, float gl,
                                                              int use_gae, int proper) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
    if (use_gae) {
      float gae = 0.0f, vnext = V[(size_t)T * M + j];   // value_preds[-1] = next_value was written by the caller
#pragma unroll 4
      for (int t = T - 1; t >= 0; --t) {
        const size_t i = (size_t)t * M + j, i1 = i + M;
        const float v = V[i], m = Mk[i1];
        const float td = R[i] + gamma * vnext * m - v;                    // buffer.py:151 / :163
        gae = td + gl * m * gae;                                          // :152 / :166  (gl = f32(gamma * gae_lambda))
        if (proper) gae = gae * Bm[i1];                                   // :153
        Ret[i] = gae + v;                                                 // :154 / :167
        vnext = v;
      }
    } else {
      float ret = Ret[(size_t)T * M + j];                // returns[-1] = next_value was written by the caller
#pragma unroll 4
      for (int t = T - 1; t >= 0; --t) {
        const size_t i = (size_t)t * M + j, i1 = i + M;
        const float m = Mk[i1];
        if (proper) {                                                     // :158-159
          const float bm = Bm[i1];
          ret = (ret * gamma * m + R[i]) * bm + (1.0f - bm) * V[i];
        } else {
          ret = ret * gamma * m + R[i];                                   // :171
        }
        Ret[i] = ret;
      }
    }
  }
}

// masks[e, a] = 0 where ANY agent of env e is done, bad_masks likewise for bad_done, reset_env[e] = any flag of any agent
// (F16sim_runner.py:144-155); one thread per env.
__global__ void __launch_bounds__(256) rollout_masks_kernel(const uint8_t* __restrict__ flags, int ld, int num_envs, int agents,
                                                            float* __restrict__ masks, float* __restrict__ bad_masks,
                                                            uint8_t* __restrict__ reset_env) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < num_envs; e += gridDim.x * blockDim.x) {
    unsigned d = 0, b = 0, x = 0;
    for (int a = 0; a < agents; ++a) {
      const size_t i = (size_t)e * agents + a;
      d |= flags[i]; b |= flags[ld + i]; x |= flags[2 * (size_t)ld + i];
    }
    const float mk = d ? 0.0f : 1.0f, bk = b ? 0.0f : 1.0f;
    for (int a = 0; a < agents; ++a) {
      masks[(size_t)e * agents + a] = mk;
      bad_masks[(size_t)e * agents + a] = bk;
    }
    if (reset_env) reset_env[e] = (d | b | x) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// K6: table aero back-end (tables_device.cuh): 44 coefficients per (alpha, beta, el) point from the NASA tables
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) f16_table_coeffs_kernel(const float* __restrict__ image, const float* __restrict__ A,
                                                               const float* __restrict__ Bd, const float* __restrict__ E,
                                                               float* __restrict__ out, int n, int ld) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* T = reinterpret_cast<float*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(T + kTablesFloats);
  stage_aero(T, image, (uint32_t)(kTablesFloats * 4), bar);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    table_coefficients(T, A[i], Bd[i], E[i], out + i, ld);
}

// nlplant with the table back-end (the getters of a table-backed F16 plug-in): one aircraft per thread
__global__ void __launch_bounds__(256) f16_table_nlplant_kernel(const float* __restrict__ image, const float* __restrict__ S,
                                                                const float* __restrict__ U, float* __restrict__ X, int n, int ld) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* T = reinterpret_cast<float*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(T + kTablesFloats);
  stage_aero(T, image, (uint32_t)(kTablesFloats * 4), bar);
  const ZeroCells zc = zero_cells(T);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s[12], u[5], c[kNumSlots], a1[kNumA1], xdot[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = S[(size_t)j * ld + i];
#pragma unroll
    for (int j = 0; j < 5; ++j) u[j] = U[(size_t)j * ld + i];
    table_env_coefs(T, zc, s[7] * kR2D, s[8] * kR2D, u[1], true, c, a1);
    const Trig g = make_trig(s);
    nlplant_from_coefs(s, u[0], u[2], u[3], u[4], g, tfac_pow(s[2]), c, 1, a1, xdot);
#pragma unroll
    for (int j = 0; j < 12; ++j) X[(size_t)j * ld + i] = xdot[j];
  }
}

// ------------------------------------------------------------------------------------------------
// stand-alone nlplant / coefficient kernels (model plug-in getters, parity tests): same device code as K1,
// two points per thread
// ------------------------------------------------------------------------------------------------
constexpr int kAuxBS = 128;
static int aux_smem_bytes(int aero_bytes) { return aero_bytes + kNumSlots * kAuxBS * 8 + 16; }

__global__ void __launch_bounds__(kAuxBS) f16_nlplant_kernel(const uint32_t* __restrict__ aero, int aero_bytes,
                                                             const float* __restrict__ S, const float* __restrict__ U,
                                                             float* __restrict__ X, int n, int ld) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* blob = reinterpret_cast<float*>(smem_raw);
  float2* coef_all = reinterpret_cast<float2*>(smem_raw + aero_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(coef_all + kNumSlots * kAuxBS);
  stage_aero(blob, aero, (uint32_t)aero_bytes, bar);
  const uint32_t wb0 = aero_base_after_staging(blob);
  const AeroTabs tabs = aero_tabs(blob, wb0);
  float2* coef2 = coef_all + threadIdx.x;
  float* cf = reinterpret_cast<float*>(coef2);
  constexpr int CS = 2 * kAuxBS;
  const int npairs = (n + 1) >> 1;
  for (int pbase = blockIdx.x * kAuxBS; pbase < npairs; pbase += gridDim.x * kAuxBS) {
    const int pr = pbase + threadIdx.x;
    const int prl = pr < npairs ? pr : npairs - 1;
    const uint32_t wb = opaque_u32(wb0);
    float s[2][12], u[2][5];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const float2 v = reinterpret_cast<const float2*>(S + (size_t)j * ld)[prl];
      s[0][j] = v.x; s[1][j] = v.y;
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float2 v = reinterpret_cast<const float2*>(U + (size_t)j * ld)[prl];
      u[0][j] = v.x; u[1][j] = v.y;
    }
    const float2 adeg = make_float2(s[0][7] * kR2D, s[1][7] * kR2D);
    const float2 bdeg = make_float2(s[0][8] * kR2D, s[1][8] * kR2D);
    ZIn2 zi;
    zscores_ab2(blob, adeg, bdeg, zi);
    zscores_el2(blob, make_float2(u[0][1], u[1][1]), zi);
    eval_ab2_nets(blob, wb, zi, coef2, kAuxBS);
    eval_el3_nets(blob, wb, zi, coef2, kAuxBS, 5);
    float xdot[2][12];
    uint32_t seg[2];
    pwl_search2<kLevelsA>(tabs.bp_a, adeg.x, adeg.y, seg[0], seg[1]);
    coef2[kEtaEl * kAuxBS] = eta_el2(tabs, make_float2(u[0][1], u[1][1]));
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float a1[kNumA1];
      alpha_coefs<kNumUsed - kFirstA1>(blob, tabs, seg[q], q == 0 ? adeg.x : adeg.y, a1);
      const Trig g = make_trig(s[q]);
      nlplant_from_coefs(s[q], u[q][0], u[q][2], u[q][3], u[q][4], g, tfac_pow(s[q][2]), cf + q, CS, a1, xdot[q]);
    }
    if (pr < npairs) {  // rows are ld >= n + (n & 1) floats long
#pragma unroll
      for (int j = 0; j < 12; ++j)
        reinterpret_cast<float2*>(X + (size_t)j * ld)[pr] = make_float2(xdot[0][j], xdot[1][j]);
    }
  }
}

__global__ void __launch_bounds__(kAuxBS) f16_coeffs_kernel(const uint32_t* __restrict__ aero, int aero_bytes,
                                                            const float* __restrict__ A, const float* __restrict__ Bd,
                                                            const float* __restrict__ E, float* __restrict__ out, int n,
                                                            int ld) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* blob = reinterpret_cast<float*>(smem_raw);
  float2* coef_all = reinterpret_cast<float2*>(smem_raw + aero_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(coef_all + kNumSlots * kAuxBS);
  stage_aero(blob, aero, (uint32_t)aero_bytes, bar);
  const uint32_t wb0 = aero_base_after_staging(blob);
  const AeroTabs tabs = aero_tabs(blob, wb0);
  float2* coef2 = coef_all + threadIdx.x;
  float* cf = reinterpret_cast<float*>(coef2);
  constexpr int CS = 2 * kAuxBS;
  const int npairs = (n + 1) >> 1;
  for (int pbase = blockIdx.x * kAuxBS; pbase < npairs; pbase += gridDim.x * kAuxBS) {
    const int pr = pbase + threadIdx.x;
    const int prl = pr < npairs ? pr : npairs - 1;
    const int i0 = min(2 * prl, n - 1), i1 = min(2 * prl + 1, n - 1);
    const uint32_t wb = opaque_u32(wb0);
    const float2 adeg = make_float2(A[i0], A[i1]), bdeg = make_float2(Bd[i0], Bd[i1]), edeg = make_float2(E[i0], E[i1]);
    ZIn2 zi;
    zscores_ab2(blob, adeg, bdeg, zi);
    zscores_el2(blob, edeg, zi);
    eval_ab2_nets(blob, wb, zi, coef2, kAuxBS);
    eval_el3_nets(blob, wb, zi, coef2, kAuxBS, 5);
    uint32_t seg[2];
    pwl_search2<kLevelsA>(tabs.bp_a, adeg.x, adeg.y, seg[0], seg[1]);
    const float2 eta2 = eta_el2(tabs, edeg);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = 2 * pr + q;
      float a1[kNumA1];
      alpha_coefs<kNumA1>(blob, tabs, seg[q], q == 0 ? adeg.x : adeg.y, a1);
      const float eta = q == 0 ? eta2.x : eta2.y;
      if (i < n) {
        for (int k = 0; k < kNumSlots; ++k) out[(size_t)k * ld + i] = k == kEtaEl ? eta : cf[q + k * CS];
#pragma unroll
        for (int k = 0; k < kNumA1; ++k) out[(size_t)(kFirstA1 + k) * ld + i] = a1[k];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The plug-in's stand-alone update(action) (F16_model.py:51-67, UAV_model.py:51-62): clamp -> control low-pass -> one
// explicit Euler step of nlplant, nothing else (no reset, obs, terminations).  The reference's own PlanningEnv
// (planning_env.py:161) and example/quick_start.ipynb drive the model this way.  recent_s / recent_u receive the state /
// controls the update started from (the reference rebinds self.recent_s = self.s before integrating).
// Same device code and evaluation order as the fused step, so env.step and model.update agree bit for bit.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAuxBS) f16_update_kernel(const uint32_t* __restrict__ aero, int aero_bytes, float* __restrict__ S,
                                                            float* __restrict__ U, float* __restrict__ RS, float* __restrict__ RU,
                                                            const float* __restrict__ action, int n, int ld, float dt) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* blob = reinterpret_cast<float*>(smem_raw);
  float2* coef_all = reinterpret_cast<float2*>(smem_raw + aero_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(coef_all + kNumSlots * kAuxBS);
  stage_aero(blob, aero, (uint32_t)aero_bytes, bar);
  const uint32_t wb0 = aero_base_after_staging(blob);
  const AeroTabs tabs = aero_tabs(blob, wb0);
  float2* coef2 = coef_all + threadIdx.x;
  float* cf = reinterpret_cast<float*>(coef2);
  constexpr int CS = 2 * kAuxBS;
  const int npairs = (n + 1) >> 1;
  for (int pbase = blockIdx.x * kAuxBS; pbase < npairs; pbase += gridDim.x * kAuxBS) {
    const int pr = pbase + threadIdx.x;
    const int prl = pr < npairs ? pr : npairs - 1;
    const bool act0 = pr < npairs, act1 = act0 && 2 * pr + 1 < n;
    const int idx[2] = {min(2 * prl, n - 1), min(2 * prl + 1, n - 1)};
    const uint32_t wb = opaque_u32(wb0);
    float s[2][12], u[2][4];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const float2 v = reinterpret_cast<const float2*>(S + (size_t)j * ld)[prl];
      s[0][j] = v.x; s[1][j] = v.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 v = reinterpret_cast<const float2*>(U + (size_t)j * ld)[prl];
      u[0][j] = v.x; u[1][j] = v.y;
    }
    if (act0 && RS) {
#pragma unroll
      for (int j = 0; j < 12; ++j) store_pair(RS + (size_t)j * ld, pr, make_float2(s[0][j], s[1][j]), act1);
    }
    if (act0 && RU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) store_pair(RU + (size_t)j * ld, pr, make_float2(u[0][j], u[1][j]), act1);
      store_pair(RU + (size_t)4 * ld, pr, make_float2(0.f, 0.f), act1);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float4 av = reinterpret_cast<const float4*>(action)[idx[q]];
      const float a[4] = {av.x, av.y, av.z, av.w};
      lowpass_controls(a, u[q]);
    }
    const float2 adeg = make_float2(s[0][7] * kR2D, s[1][7] * kR2D);
    const float2 bdeg = make_float2(s[0][8] * kR2D, s[1][8] * kR2D);
    ZIn2 zi;
    zscores_ab2(blob, adeg, bdeg, zi);
    zscores_el2(blob, make_float2(u[0][1], u[1][1]), zi);
    eval_ab2_nets(blob, wb, zi, coef2, kAuxBS);
    eval_el3_nets(blob, wb, zi, coef2, kAuxBS, 5);
    uint32_t seg[2];
    pwl_search2<kLevelsA>(tabs.bp_a, adeg.x, adeg.y, seg[0], seg[1]);
    coef2[kEtaEl * kAuxBS] = eta_el2(tabs, make_float2(u[0][1], u[1][1]));
    const float h = dt - 0.0f;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float a1[kNumA1], xdot[12];
      alpha_coefs<kNumUsed - kFirstA1>(blob, tabs, seg[q], q == 0 ? adeg.x : adeg.y, a1);
      const Trig g = make_trig(s[q]);
      nlplant_from_coefs(s[q], u[q][0], u[q][2], u[q][3], 0.0f, g, tfac_pow(s[q][2]), cf + q, CS, a1, xdot);
#pragma unroll
      for (int j = 0; j < 12; ++j) s[q][j] = s[q][j] + h * xdot[j];
    }
    if (act0) {
#pragma unroll
      for (int j = 0; j < 12; ++j) store_pair(S + (size_t)j * ld, pr, make_float2(s[0][j], s[1][j]), act1);
#pragma unroll
      for (int j = 0; j < 4; ++j) store_pair(U + (size_t)j * ld, pr, make_float2(u[0][j], u[1][j]), act1);
      store_pair(U + (size_t)4 * ld, pr, make_float2(0.f, 0.f), act1);   // lef = 0 (F16_model.py:57)
    }
  }
}

__global__ void __launch_bounds__(256) f16_table_update_kernel(const float* __restrict__ image, float* __restrict__ S, float* __restrict__ U,
                                                               float* __restrict__ RS, float* __restrict__ RU,
                                                               const float* __restrict__ action, int n, int ld, float dt) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* T = reinterpret_cast<float*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(T + kTablesFloats);
  stage_aero(T, image, (uint32_t)(kTablesFloats * 4), bar);
  const ZeroCells zc = zero_cells(T);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s[12], u[4], c[kNumSlots], a1[kNumA1], xdot[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = S[(size_t)j * ld + i];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = U[(size_t)j * ld + i];
    if (RS) {
#pragma unroll
      for (int j = 0; j < 12; ++j) RS[(size_t)j * ld + i] = s[j];
    }
    if (RU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) RU[(size_t)j * ld + i] = u[j];
      RU[(size_t)4 * ld + i] = 0.0f;
    }
    const float4 av = reinterpret_cast<const float4*>(action)[i];
    const float a[4] = {av.x, av.y, av.z, av.w};
    lowpass_controls(a, u);
    table_env_coefs(T, zc, s[7] * kR2D, s[8] * kR2D, u[1], true, c, a1);
    const Trig g = make_trig(s);
    nlplant_from_coefs(s, u[0], u[2], u[3], 0.0f, g, tfac_pow(s[2]), c, 1, a1, xdot);
    const float h = dt - 0.0f;
#pragma unroll
    for (int j = 0; j < 12; ++j) S[(size_t)j * ld + i] = s[j] + h * xdot[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) U[(size_t)j * ld + i] = u[j];
    U[(size_t)4 * ld + i] = 0.0f;
  }
}

__global__ void __launch_bounds__(256) uav_update_kernel(float* __restrict__ S, float* __restrict__ U, float* __restrict__ RS,
                                                         const float* __restrict__ action, int n, int ld, float dt) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s[12], F[3], xdot[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = S[(size_t)j * ld + i];
#pragma unroll
    for (int j = 0; j < 3; ++j) F[j] = U[(size_t)j * ld + i];
    if (RS) {
#pragma unroll
      for (int j = 0; j < 12; ++j) RS[(size_t)j * ld + i] = s[j];
    }
    const float4 av = reinterpret_cast<const float4*>(action)[i];
    const float a[3] = {fminf(fmaxf(av.x, -1.0f), 1.0f), fminf(fmaxf(av.y, -1.0f), 1.0f), fminf(fmaxf(av.z, -1.0f), 1.0f)};
#pragma unroll
    for (int j = 0; j < 3; ++j) F[j] = 0.9f * F[j] + 0.1f * a[j] * 27000.0f;
    uav_nlplant(s, F, xdot);
    const float h = dt - 0.0f;
#pragma unroll
    for (int j = 0; j < 12; ++j) S[(size_t)j * ld + i] = s[j] + h * xdot[j];
#pragma unroll
    for (int j = 0; j < 3; ++j) U[(size_t)j * ld + i] = F[j];
  }
}

// (alpha,beta)-MLP outputs at alpha = beta = 0, written into the image at np_aero_create (same device code as K1,
// so a reset lane sees bit-identical values whether it takes the constants or an evaluation).
__global__ void __launch_bounds__(32) f16_c0_kernel(uint32_t* aero, int aero_bytes) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* blob = reinterpret_cast<float*>(smem_raw);
  float2* coef2 = reinterpret_cast<float2*>(smem_raw + aero_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(coef2 + kNumSlots * 32);
  stage_aero(blob, aero, (uint32_t)aero_bytes, bar);
  const uint32_t wb = aero_base_after_staging(blob);
  ZIn2 zi;  // every lane evaluates the same point into its own slots (uniform control flow); lane 0 publishes
  zscores_ab2(blob, make_float2(0.0f * kR2D, 0.0f * kR2D), make_float2(0.0f * kR2D, 0.0f * kR2D), zi);
  eval_ab2_nets(blob, wb, zi, coef2 + threadIdx.x, 32);
  if (threadIdx.x == 0) {
    float* c0 = reinterpret_cast<float*>(aero) + reinterpret_cast<const int32_t*>(blob)[kHdrC0];
    for (int k = 0; k < kNumAB2; ++k) c0[k] = coef2[(kFirstAB2 + k) * 32].x;
  }
}

