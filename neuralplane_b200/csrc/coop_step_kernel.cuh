// K1c: BaseEnv.step for SMALL populations (the reference trains at 3 000 envs, scripts/train_heading.sh:13), where a step
// is latency bound: its duration is one warp's pass through ~24 000 dependent instructions of K1, not a throughput.
// Included by nplane.cu after K1 (uses its StepParams, tile constants and device functions).
//
// The cure is to cut the pass, not to add warps: a CTA of FOUR warps flies 32 aircraft pairs (one pair per lane, as K1), and
// the 21 MLP evaluations of a step -- >90 % of the instructions -- are dealt out over the four warps (each lane of every
// warp holds the same pair, so every warp can evaluate any net for it):
//   warp w: the (alpha, beta) nets {w, w+4, w+8, w+12} (Cy dCz_lef dCy_r30 dCy_a20 | dCl_a20 dCm_lef dCn_r30 dCy_a20_lef | ...),
//           pass 0: three-input nets Cx Cz | Cm | Cn | Cl (+ eta_el on warp 3), pass 1: Cx | Cz | - | -.
// Outputs meet in the CTA's coefficient slots ([kNumSlots][32] float2 of shared memory); after a __syncthreads warp 0 flies
// aircraft 2*pair and warp 1 aircraft 2*pair+1 through the scalar tail (trig, atmosphere, PWL coefficients, forces, moments,
// Euler, observation, terminations, reward), and hands the new (alpha, beta) back through shared memory for pass 1.
// The prologue (loads, episodic reset, control lag, cache-hit test) is recomputed by all four warps: it is short and
// CTA-uniform, which keeps every branch around the barriers uniform.
//
// Every value is produced by the same device function on the same operands as in K1 (the two halves of an FFMA2 are
// independent), so the step is BIT-IDENTICAL to K1's: tests/test_gpu_plugin.py compares them directly.
#pragma once

constexpr int kCoopBS = 128, kCoopPairs = 32;
static int coop_smem_bytes(int aero_bytes) {
  return aero_bytes + kNumSlots * kCoopPairs * 8 + kObsTileFloats * 4 + 4 * kCoopPairs * 4 + 16;
}

template <int TASK>
__global__ void __launch_bounds__(kCoopBS, 2) f16_step_coop_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* blob = reinterpret_cast<float*>(smem_raw);
  float2* coef_all = reinterpret_cast<float2*>(smem_raw + p.aero_bytes);     // [kNumSlots][32] float2
  float* otile = reinterpret_cast<float*>(coef_all + kNumSlots * kCoopPairs);  // the CTA's 64 observation rows
  float* xa = otile + kObsTileFloats;                                        // [2][32]: alpha' of aircraft q of pair `lane`
  float* xb = xa + 2 * kCoopPairs;                                           // [2][32]: beta'
  uint64_t* bar = reinterpret_cast<uint64_t*>(xb + 2 * kCoopPairs);

  stage_aero(blob, p.aero, (uint32_t)p.aero_bytes, bar);
  const uint32_t wb0 = aero_base_after_staging(blob);
  const AeroTabs tabs = aero_tabs(blob, wb0);
  const float* c0 = blob + reinterpret_cast<const int32_t*>(blob)[kHdrC0];

  const np_env_cfg& c = p.cfg;
  const int n = c.n, ld = c.ld;
  const int npairs = (n + 1) >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool owner = warp < 2;   // warps 0 / 1 fly aircraft 0 / 1 of each pair through the scalar tail
  const bool q1 = (warp & 1) != 0;
  float2* coef2 = coef_all + lane;                                           // slot k of this lane's pair: coef2[k * 32]
  const float* cq = reinterpret_cast<const float*>(coef2) + (q1 ? 1 : 0);
  constexpr int CS = 2 * kCoopPairs;
  const bool use_cache = c.use_coef_cache != 0;
  bool first_iter = true;

  const int pend = p.pair_end < npairs ? p.pair_end : npairs;
  for (int pbase = p.pair_begin + blockIdx.x * kCoopPairs; pbase < pend; pbase += gridDim.x * kCoopPairs) {
    const int pr = pbase + lane;
    const int prl = pr < pend ? pr : pend - 1;
    const bool act[2] = {pr < pend && 2 * pr < n, pr < pend && 2 * pr + 1 < n};
    const int idx[2] = {min(2 * prl, n - 1), min(2 * prl + 1, n - 1)};
    const bool staged = __all_sync(0xffffffffu, act[1]) && ((reinterpret_cast<uintptr_t>(p.obs) & 15) == 0);

    // ---- load (every warp: both aircraft of the lane's pair) ---------------------------------------------------
    float s[2][12], u[2][4], tgt[2][3], a[2][4];
    int steps[2];
    bool rst[2];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const float2 v = reinterpret_cast<const float2*>(p.s + (size_t)j * ld)[prl];
      s[0][j] = v.x; s[1][j] = v.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 v = reinterpret_cast<const float2*>(p.u + (size_t)j * ld)[prl];
      u[0][j] = v.x; u[1][j] = v.y;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float2 v = reinterpret_cast<const float2*>(p.tgt + (size_t)j * ld)[prl];
      tgt[0][j] = v.x; tgt[1][j] = v.y;
    }
    {
      const int2 v = reinterpret_cast<const int2*>(p.step_count)[prl];
      steps[0] = v.x; steps[1] = v.y;
      const uchar2 f0 = reinterpret_cast<const uchar2*>(p.flags)[prl];
      const uchar2 f1 = reinterpret_cast<const uchar2*>(p.flags + ld)[prl];
      const uchar2 f2 = reinterpret_cast<const uchar2*>(p.flags + 2 * (size_t)ld)[prl];
      rst[0] = (f0.x | f1.x | f2.x) != 0;
      rst[1] = (f0.y | f1.y | f2.y) != 0;
    }
    // this warp's four cached (alpha, beta)-net outputs go straight into the slots it would otherwise compute
    float2 ka = make_float2(0.f, 0.f), kb = ka;
    if (use_cache) {
      ka = reinterpret_cast<const float2*>(p.cache + (size_t)kNumAB2 * ld)[prl];
      kb = reinterpret_cast<const float2*>(p.cache + (size_t)(kNumAB2 + 1) * ld)[prl];
#pragma unroll
      for (int k = warp; k < kNumAB2; k += 4) coef2[(kFirstAB2 + k) * kCoopPairs] = reinterpret_cast<const float2*>(p.cache + (size_t)k * ld)[prl];
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float4 av = reinterpret_cast<const float4*>(p.action)[idx[q]];
      a[q][0] = av.x; a[q][1] = av.y; a[q][2] = av.z; a[q][3] = av.w;
    }

    // ---- episodic reset (env_base.py:83-97) -----------------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (rst[q]) {
        const Draws r = reset_draws(p, idx[q]);
        reset_aircraft(c, TASK, r, s[q], u[q], tgt[q]);
        steps[q] = 0;
      }
    }
    if (warp == 0) count_cause2(p.counters, 7, rst[0] && act[0], rst[1] && act[1]);

    // ---- cache hit / reset constants / miss: the same decision in all four warps ------------------------------------
    bool hit[2] = {rst[0], rst[1]};
    if (use_cache) {
      hit[0] |= __float_as_uint(ka.x) == __float_as_uint(s[0][7]) && __float_as_uint(kb.x) == __float_as_uint(s[0][8]);
      hit[1] |= __float_as_uint(ka.y) == __float_as_uint(s[1][7]) && __float_as_uint(kb.y) == __float_as_uint(s[1][8]);
    }
    const bool miss = __any_sync(0xffffffffu, !(hit[0] && hit[1])) != 0;
    if (!miss && (rst[0] || rst[1])) {
#pragma unroll
      for (int k = warp; k < kNumAB2; k += 4) {
        float2 v = coef2[(kFirstAB2 + k) * kCoopPairs];
        if (rst[0]) v.x = c0[k];
        if (rst[1]) v.y = c0[k];
        coef2[(kFirstAB2 + k) * kCoopPairs] = v;
      }
    }

    // ---- control lag (F16_model.py:52-57) -----------------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < 2; ++q) {
#pragma unroll
      for (int j = 0; j < 4; ++j) a[q][j] = fminf(fmaxf(a[q][j], -1.0f), 1.0f);
      u[q][0] = 0.9f * u[q][0] + 0.1f * a[q][0] * 0.225f * 76300.0f / DC(0.3048f);
      u[q][1] = 0.9f * u[q][1] + 0.1f * a[q][1] * 45.0f;
      u[q][2] = 0.9f * u[q][2] + 0.1f * a[q][2] * 45.0f;
      u[q][3] = 0.9f * u[q][3] + 0.1f * a[q][3] * 45.0f;
    }

    // ---- the aircraft this warp flies through the tail ----------------------------------------------------------------
    float sq[12], uq[4], tq[3];
#pragma unroll
    for (int j = 0; j < 12; ++j) sq[j] = q1 ? s[1][j] : s[0][j];
#pragma unroll
    for (int j = 0; j < 4; ++j) uq[j] = q1 ? u[1][j] : u[0][j];
#pragma unroll
    for (int j = 0; j < 3; ++j) tq[j] = q1 ? tgt[1][j] : tgt[0][j];
    int stepq = q1 ? steps[1] : steps[0];
    const bool actq = q1 ? act[1] : act[0];
    const int idxq = q1 ? idx[1] : idx[0];
    const int row = 2 * pr + (q1 ? 1 : 0);
    float2 al = make_float2(s[0][7], s[1][7]), be = make_float2(s[0][8], s[1][8]);
    const float2 el = make_float2(u[0][1], u[1][1]);
    bool badq = false, doneq = false, excq = false;
    float rewq = 0.0f;
    int causesq = 0;

    if (!first_iter && threadIdx.x == 0) bulk_wait_read0();   // the tile may still be being read by the previous bulk store
    first_iter = false;

#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const float2 adeg = make_float2(al.x * kR2D, al.y * kR2D);
      const float2 bdeg = make_float2(be.x * kR2D, be.y * kR2D);
      const uint32_t wb = opaque_u32(wb0);
      ZIn2 zi;
      zscores_ab2(blob, adeg, bdeg, zi);
      zscores_el2(blob, el, zi);
      if (pass == 1 || miss) eval_ab2_nets_quarter(blob, wb, zi, coef2, kCoopPairs, warp);
      eval_el3_nets_quarter(blob, wb, zi, coef2, kCoopPairs, warp, pass == 0);
      if (pass == 0 && warp == 3) coef2[kEtaEl * kCoopPairs] = eta_el2(tabs, el);
      __syncthreads();   // all 22 slots of the 32 pairs are in place

      if (pass == 1 && use_cache && act[0]) {   // the next step's Euler derivative needs exactly these
#pragma unroll
        for (int k = warp; k < kNumAB2; k += 4) store_pair(p.cache + (size_t)k * ld, pr, coef2[(kFirstAB2 + k) * kCoopPairs], act[1]);
        if (warp == 2) store_pair(p.cache + (size_t)kNumAB2 * ld, pr, al, act[1]);
        if (warp == 3) store_pair(p.cache + (size_t)(kNumAB2 + 1) * ld, pr, be, act[1]);
      }

      if (owner) {
        const float aq = q1 ? adeg.y : adeg.x;
        uint32_t seg, seg_unused;
        pwl_search2<kLevelsA>(tabs.bp_a, aq, aq, seg, seg_unused);
        const Trig g = make_trig(sq);
        const float tp = tfac_pow(sq[2]);
        float a1[kNumA1];
        alpha_coefs<kNumUsed - kFirstA1>(blob, tabs, seg, aq, a1);
        const ForcePart fp = force_part(sq, uq[0], uq[2], uq[3], 0.0f, g, tp, cq, CS, a1);
        if (pass == 0) {
          float xdot[12];
          nlplant_kin_moments(sq, uq[2], uq[3], 0.0f, g, fp.qbar, fp.vt, fp.b, fp.t, cq, CS, a1, xdot);
          xdot[6] = fp.f.vt_dot; xdot[7] = fp.f.alpha_dot; xdot[8] = fp.f.beta_dot;
          const float h = c.dt - 0.0f;
#pragma unroll
          for (int j = 0; j < 12; ++j) sq[j] = sq[j] + h * xdot[j];
          stepq += 1;
          xa[(q1 ? kCoopPairs : 0) + lane] = sq[7];
          xb[(q1 ? kCoopPairs : 0) + lane] = sq[8];
        } else {
          float o[NP_NUM_OBS];
          make_obs(c, TASK, sq, uq, tq, g, eas2tas_of(tp), o);
          add_obs_noise(p, idxq, o);
          if (staged) {
            float2* orow = reinterpret_cast<float2*>(otile + (2 * lane + (q1 ? 1 : 0)) * NP_NUM_OBS);
#pragma unroll
            for (int j = 0; j < NP_NUM_OBS / 2; ++j) orow[j] = make_float2(o[2 * j], o[2 * j + 1]);
            fence_async_smem();
          } else if (actq) {
            float2* orow = reinterpret_cast<float2*>(p.obs + (size_t)row * NP_NUM_OBS);
#pragma unroll
            for (int j = 0; j < NP_NUM_OBS / 2; ++j) orow[j] = make_float2(o[2 * j], o[2 * j + 1]);
          }
          const Verdict v = judge_state<false, TASK>(c, sq, tq, g, fp.f, stepq);
          excq = v.exc; badq = v.bad; doneq = v.done;
          rewq = v.rw + (float)(-200 * (int)badq + 200 * (int)doneq);
          causesq = actq ? v.causes : 0;
        }
      }
      if (pass == 0) {
        __syncthreads();   // the new (alpha, beta) of both aircraft, for every warp's share of pass 1
        al = make_float2(xa[lane], xa[kCoopPairs + lane]);
        be = make_float2(xb[lane], xb[kCoopPairs + lane]);
      }
    }

    if (owner) {
#pragma unroll
      for (int w = 0; w < 7; ++w) count_cause1(p.counters, w, (causesq >> w) & 1);
      if (actq) {
#pragma unroll
        for (int j = 0; j < 12; ++j) p.s[(size_t)j * ld + row] = sq[j];
#pragma unroll
        for (int j = 0; j < 4; ++j) p.u[(size_t)j * ld + row] = uq[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) p.tgt[(size_t)j * ld + row] = tq[j];
        p.reward[row] = rewq;
        p.step_count[row] = stepq;
        p.flags[row] = doneq ? 1 : 0;
        p.flags[ld + row] = badq ? 1 : 0;
        p.flags[2 * (size_t)ld + row] = excq ? 1 : 0;
        if (p.flags_mirror) {
          const size_t ml = (size_t)p.flags_mirror_ld;
          p.flags_mirror[row] = doneq ? 1 : 0;
          p.flags_mirror[ml + row] = badq ? 1 : 0;
          p.flags_mirror[2 * ml + row] = excq ? 1 : 0;
        }
      }
    }
    __syncthreads();   // tile complete; slots and exchange rows free for the next 32 pairs
    if (staged) {      // tile -> obs[64 rows], 5 632 contiguous bytes
      float* dst = p.obs + (size_t)(2 * pbase) * NP_NUM_OBS;
      if (!p.obs_stg) {
        if (threadIdx.x == 0) {
          bulk_s2g(dst, otile, kObsTileFloats * 4);
          bulk_commit();
        }
      } else {
        const float4* src = reinterpret_cast<const float4*>(otile);
        for (int k = threadIdx.x; k < kObsTileFloats / 4; k += kCoopBS) reinterpret_cast<float4*>(dst)[k] = src[k];
        __syncthreads();
      }
    }
  }
  if (threadIdx.x == 0) bulk_wait0();   // shared memory must outlive the last bulk store
}
