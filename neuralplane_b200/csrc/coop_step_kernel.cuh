// K1c: BaseEnv.step -- and, by MODE, PlanningEnv.step and the pair-sharded combat step -- for SMALL populations (the reference
// trains at 3 000 envs, scripts/train_heading.sh:13, and 10 000 planning envs, train_tracking.sh), where a step is latency bound: its duration is one warp's pass through ~24 000 dependent instructions of K1, not a throughput.
// Included by nplane.cu after K1 (uses its StepParams, tile constants and device functions).
//
// The cure is to cut the pass, not to add aircraft: a CTA of NW = 4 or 8 warps flies 32 aircraft pairs (one pair per lane, as
// K1), and the 21 MLP evaluations of a step -- most of the instructions -- are dealt out over the warps (each lane of every warp
// holds the same pair, so every warp can evaluate any net for it): kCoopShare below.  Outputs meet in the CTA's coefficient
// slots ([kNumSlots][32] float2 of shared memory); after a __syncthreads warp 0 flies aircraft 2*pair and warp 1 aircraft
// 2*pair+1 through the scalar tail (forces, moments, Euler; terminations, reward), and hands the new (alpha, beta) back through
// shared memory for pass 1.  Warps 0 / 1 get a small MLP share and use the slack for everything in the tail that needs no MLP
// output -- trig, atmosphere, the PWL alpha coefficients, and in pass 1 the whole observation row with its noise -- so that
// only forces / moments / Euler / verdict remain behind the barriers.
// The prologue (loads, episodic reset, control lag, cache-hit test) is recomputed by all warps: it is short and CTA-uniform,
// which keeps every branch around the barriers uniform.
//   NW = 8 (one CTA per SM, two warps per scheduler): up to 148 x 64 = 9 472 aircraft in one wave;
//   NW = 4 (two CTAs per SM): up to 18 944.
//
// Every value is produced by the same device function on the same operands as in K1 (the two halves of an FFMA2 are
// independent), so the step is BIT-IDENTICAL to K1's: tests/test_gpu_plugin.py compares them directly.
#pragma once
#include <type_traits>

constexpr int kCoopPairs = 32;
// Net groups of one architecture and z-score selection (f16_device.cuh eval_ab2_nets): first (alpha, beta) net, nets in the group
constexpr int kCoopGroupK0[6] = {kCy, kdCx_lef, kdCz_lef, kdCy_r30, kdCy_a20, kdCy_a20_lef};
constexpr int kCoopGroupN[6] = {2, 2, 4, 4, 1, 3};
// kCoopShare[NW == 8][warp][g] = {first, count}: the nets [first, first + count) of group g this warp evaluates; g = 6: the
// three-input nets Cx Cz Cm Cn Cl in pass 0 (all five), g = 7: in pass 1 (Cx Cz).  MACs per net: 250 250 295 295 350 650 | 270.
//   NW = 4, pass 1:  1 190 | 1 490 | 1 760 | 1 760   (warps 0 / 1 also prepare their aircraft's tail)
//   NW = 8, pass 1:    545 |   295 | 900 | 900 | 900 | 885 | 885 | 890
struct CoopRange { signed char first, count; };
#define NP_COOP_SHARE_TABLE { \
    {   /* NW = 4 */ \
        {{0, 1}, {0, 0}, {0, 1}, {0, 1}, {0, 1}, {0, 0}, {4, 1}, {0, 0}}, \
        {{1, 1}, {0, 0}, {1, 1}, {1, 1}, {0, 0}, {0, 1}, {0, 0}, {0, 0}}, \
        {{0, 0}, {0, 1}, {2, 1}, {2, 1}, {0, 0}, {1, 1}, {0, 2}, {0, 1}}, \
        {{0, 0}, {1, 1}, {3, 1}, {3, 1}, {0, 0}, {2, 1}, {2, 2}, {1, 1}}, \
        {}, {}, {}, {}, \
    }, \
    {   /* NW = 8 */ \
        {{0, 1}, {0, 0}, {0, 1}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}}, \
        {{0, 0}, {0, 0}, {1, 1}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}}, \
        {{1, 1}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 1}, {0, 1}, {0, 0}}, \
        {{0, 0}, {0, 1}, {0, 0}, {0, 0}, {0, 0}, {1, 1}, {1, 1}, {0, 0}}, \
        {{0, 0}, {1, 1}, {0, 0}, {0, 0}, {0, 0}, {2, 1}, {2, 1}, {0, 0}}, \
        {{0, 0}, {0, 0}, {2, 2}, {0, 1}, {0, 0}, {0, 0}, {3, 1}, {0, 0}}, \
        {{0, 0}, {0, 0}, {0, 0}, {1, 3}, {0, 0}, {0, 0}, {4, 1}, {0, 0}}, \
        {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 1}, {0, 0}, {0, 0}, {0, 2}}, \
    }}
constexpr CoopRange kCoopShare[2][8][8] = NP_COOP_SHARE_TABLE;          // compile-time uses (ownership, checks)
__constant__ CoopRange kCoopShareDev[2][8][8] = NP_COOP_SHARE_TABLE;      // the kernel's warp-indexed copy
constexpr int coop_eta_warp(int v) { return v ? 7 : 1; }   // eta_el (a PWL lookup) in pass 0
// which warp owns (alpha, beta) net k (0 ... 15): it loads / patches / evaluates that slot, so no two warps write one slot
constexpr int coop_owner(int v, int k) {
  int g = 0, base = 0;
  while (k >= base + kCoopGroupN[g]) { base += kCoopGroupN[g]; ++g; }
  for (int w = 0; w < 8; ++w)
    if (k - base >= kCoopShare[v][w][g].first && k - base < kCoopShare[v][w][g].first + kCoopShare[v][w][g].count) return w;
  return -1;
}
constexpr bool coop_shares_ok(int v, int nw) {
  for (int k = 0; k < kNumAB2; ++k)
    if (coop_owner(v, k) < 0 || coop_owner(v, k) >= nw) return false;
  for (int g = 0; g < 8; ++g) {   // every net of every group exactly once
    int total = 0;
    for (int w = 0; w < nw; ++w) total += kCoopShare[v][w][g].count;
    if (total != (g < 6 ? kCoopGroupN[g] : g == 6 ? 5 : 2)) return false;
  }
  return true;
}
static_assert(coop_shares_ok(0, 4) && coop_shares_ok(1, 8), "K1c: the MLP shares must cover every net exactly once");
template <int K, int N, class F>
__device__ __forceinline__ void coop_static_for(F&& f) {
  if constexpr (K < N) {
    f(std::integral_constant<int, K>{});
    coop_static_for<K + 1, N>(f);
  }
}
static_assert(kCoopGroupK0[0] - kFirstAB2 == 0 && kCoopGroupK0[5] + 3 - kFirstAB2 == kNumAB2, "K1c: group table vs net enum");
#ifdef NPLANE_COOP_TIMING   // debug build (tools/k1c_phases.py): per-warp clock stamps of CTA 0, left in the first obs rows
#define NP_COOP_STAMP_INIT() __shared__ long long np_stamps[8][10]; long long* stamps = np_stamps[warp]; const long long t_start = clock64()
#define NP_COOP_STAMP(i) do { if (lane == 0) stamps[i] = clock64() - t_start; } while (0)
#define NP_COOP_STAMP_FLUSH() do { NP_COOP_STAMP(9); __syncthreads(); if (blockIdx.x == 0 && lane < 10) p.obs[warp * 10 + lane] = (float)np_stamps[warp][lane]; } while (0)
#else
#define NP_COOP_STAMP_INIT() do { } while (0)
#define NP_COOP_STAMP(i) do { } while (0)
#define NP_COOP_STAMP_FLUSH() do { } while (0)
#endif
static int coop_smem_bytes(int aero_bytes) {
  return aero_bytes + kNumSlots * kCoopPairs * 8 + kObsTileFloats * 4 + 12 * kCoopPairs * 4 + 16;
}

// PLAN = true: PlanningEnv.step (planning_env.py:144-177) -- K1's MODE_PLAN on the same CTA shape.  train_tracking.sh runs it
// at 10 000 envs: 50 FDM sub-steps per env step under the fused PID controller, so the pass that K1c shortens is walked 100
// times per launch.  Warps 0 / 1 run their aircraft's controller and control lag at the top of every sub-step and publish the
// new elevator deflection (the one control the nets see) through shared memory; state, controller state and flags stay in
// their registers for the whole env step; the observation row is produced in the last sub-step only.
// MODE_COMBAT: SingleCombatEnv.step / MultipleCombatEnv.step, pair-sharded (K1's MODE_COMBAT without p.records): the lane's two
// aircraft are the duel.  Warps 0 / 1 run the attitude-demand controller per sub-step, exchange their new positions with the
// (alpha, beta) hand-off for the Crash check, and after the last sub-step warp 1 hands its aircraft to warp 0, which produces
// the pair's 15-D observations, rewards and blood exactly as K1's thread does (combat_outputs).
template <int TASK, int NW, int MODE = MODE_STEP>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 2 : 1) f16_step_coop_kernel(const __grid_constant__ StepParams p) {
  static_assert(NW == 4 || NW == 8, "K1c: four or eight warps");
  constexpr bool PLAN = MODE == MODE_PLAN, COMBAT = MODE == MODE_COMBAT, SUBSTEPS = PLAN || COMBAT;
  constexpr int V = NW == 8 ? 1 : 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* blob = reinterpret_cast<float*>(smem_raw);
  float2* coef_all = reinterpret_cast<float2*>(smem_raw + p.aero_bytes);     // [kNumSlots][32] float2
  float* otile = reinterpret_cast<float*>(coef_all + kNumSlots * kCoopPairs);  // the CTA's 64 observation rows
  float* xa = otile + kObsTileFloats;                                        // [2][32]: alpha' of aircraft q of pair `lane`
  float* xb = xa + 2 * kCoopPairs;                                           // [2][32]: beta'
  float* xe = xb + 2 * kCoopPairs;                                           // [2][32]: elevator after the control lag (PLAN / COMBAT)
  float* xp = xe + 2 * kCoopPairs;                                           // [2][3][32]: position after the Euler step (COMBAT: Crash)
  uint64_t* bar = reinterpret_cast<uint64_t*>(xp + 6 * kCoopPairs);

  stage_aero_issue(blob, p.aero, (uint32_t)p.aero_bytes, bar);   // waited for below, behind the first state loads
  // Launched with programmatic stream serialisation: everything above (CTA start-up, the image copies: immutable data) may
  // overlap the tail of the previous kernel in the stream -- e.g. the previous env step.  Nothing that kernel wrote is read, and
  // nothing is written, before this wait; the next kernel may then start its own prologue behind us.
  grid_dependency_wait();
  grid_launch_dependents();

  const np_env_cfg& c = p.cfg;
  const int n = c.n, ld = c.ld;
  const int npairs = (n + 1) >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool owner = warp < 2;   // warps 0 / 1 fly aircraft 0 / 1 of each pair through the scalar tail
  const bool q1 = (warp & 1) != 0;
  NP_COOP_STAMP_INIT();
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
    const bool staged = !COMBAT && __all_sync(0xffffffffu, act[1]) && ((reinterpret_cast<uintptr_t>(p.obs) & 15) == 0);

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
      const float2 v = COMBAT ? make_float2(0.f, 0.f) : reinterpret_cast<const float2*>(p.tgt + (size_t)j * ld)[prl];
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
      coop_static_for<0, kNumAB2>([&](auto kc) {
        constexpr int k = decltype(kc)::value, ow = coop_owner(V, k);
        if (ow == warp) coef2[(kFirstAB2 + k) * kCoopPairs] = reinterpret_cast<const float2*>(p.cache + (size_t)k * ld)[prl];
      });
    }
    if constexpr (!SUBSTEPS) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 av = reinterpret_cast<const float4*>(p.action)[idx[q]];
        a[q][0] = av.x; a[q][1] = av.y; a[q][2] = av.z; a[q][3] = av.w;
      }
    }

    // the RNG counter of this launch: one more L2 round trip in flight with the state loads (requested AFTER them: its first
    // consumer is an add the compiler places right behind the load, and a warp issues in order)
    const uint32_t rng = rng_step(p);
    // the aero image (64 KB, TMA) has been arriving while the loads above were issued; later iterations pass straight through
    mbar_wait(bar, 0);
    const uint32_t wb0 = aero_base_after_staging(blob);
    const AeroTabs tabs = aero_tabs(blob, wb0);
    const float* c0 = blob + reinterpret_cast<const int32_t*>(blob)[kHdrC0];

    // ---- episodic reset (env_base.py:83-97); COMBAT: env-level (singlecombat_env.py:207-238), either flag re-initialises the pair
    float blood[2] = {0.f, 0.f};
    if constexpr (COMBAT) {
      const float2 bv = reinterpret_cast<const float2*>(p.blood)[prl];
      blood[0] = bv.x; blood[1] = bv.y;
      bool r = rst[0] || rst[1];
      // MultipleCombat (multiplecombat_env.py:207-238): an env is TWO adjacent duels = two adjacent lanes; any flag resets all four
      if (c.combat_pairs_per_env == 2) r |= __shfl_xor_sync(0xffffffffu, (int)r, 1) != 0;
      rst[0] = rst[1] = r;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (rst[q]) {
        const Draws r = reset_draws(p, idx[q], rng);
        if constexpr (COMBAT) {
          combat_reset_aircraft(c, r, s[q], u[q]);
          blood[q] = 100.0f;
        } else {
          reset_aircraft(c, TASK, r, s[q], u[q], tgt[q]);
        }
        steps[q] = 0;
      }
    }
    if (warp == 0) count_cause2(p.counters, 7, rst[0] && act[0], rst[1] && act[1]);

    // ---- cache hit / reset constants / miss: the same decision in every warp ------------------------------------
    bool hit[2] = {rst[0], rst[1]};
    if (use_cache) {
      hit[0] |= __float_as_uint(ka.x) == __float_as_uint(s[0][7]) && __float_as_uint(kb.x) == __float_as_uint(s[0][8]);
      hit[1] |= __float_as_uint(ka.y) == __float_as_uint(s[1][7]) && __float_as_uint(kb.y) == __float_as_uint(s[1][8]);
    }
    const bool miss = __any_sync(0xffffffffu, !(hit[0] && hit[1])) != 0;
    if (!miss && (rst[0] || rst[1])) {
      coop_static_for<0, kNumAB2>([&](auto kc) {
        constexpr int k = decltype(kc)::value, ow = coop_owner(V, k);
        if (ow == warp) {
          float2 v = coef2[(kFirstAB2 + k) * kCoopPairs];
          if (rst[0]) v.x = c0[k];
          if (rst[1]) v.y = c0[k];
          coef2[(kFirstAB2 + k) * kCoopPairs] = v;
        }
      });
    }

    // ---- control lag (F16_model.py:52-57); PLAN / COMBAT: per sub-step, by the warp that flies the aircraft ---------------
    if constexpr (!SUBSTEPS) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int j = 0; j < 4; ++j) a[q][j] = fminf(fmaxf(a[q][j], -1.0f), 1.0f);
        u[q][0] = 0.9f * u[q][0] + 0.1f * a[q][0] * 0.225f * 76300.0f / DC(0.3048f);
        u[q][1] = 0.9f * u[q][1] + 0.1f * a[q][1] * 45.0f;
        u[q][2] = 0.9f * u[q][2] + 0.1f * a[q][2] * 45.0f;
        u[q][3] = 0.9f * u[q][3] + 0.1f * a[q][3] * 45.0f;
      }
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
    float2 el = make_float2(u[0][1], u[1][1]);
    bool badq = false, doneq = false, excq = false;
    float rewq = 0.0f;
    int causesq = 0;
    // planning step: targets from the high-level action (planning_env.py:146-152) and the controller state, owner warps only
    float plan_tgt[3] = {0.f, 0.f, 0.f}, a_cmd[4] = {0.f, 0.f, 0.f, 0.f}, pid[SUBSTEPS ? kPidRows : 1];
    if constexpr (COMBAT) {   // the caller's clamped 4-D action, kept for all sub-steps, and the controller state
      if (owner) {
        const float4 av = reinterpret_cast<const float4*>(p.action)[idxq];
        a_cmd[0] = fminf(fmaxf(av.x, -1.0f), 1.0f); a_cmd[1] = fminf(fmaxf(av.y, -1.0f), 1.0f);
        a_cmd[2] = fminf(fmaxf(av.z, -1.0f), 1.0f); a_cmd[3] = fminf(fmaxf(av.w, -1.0f), 1.0f);
#pragma unroll
        for (int j = 0; j < kPidRows; ++j) pid[j] = p.pid[(size_t)j * ld + idxq];
      }
    }
    if constexpr (PLAN) {
      if (owner) {
        float a3[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) a3[j] = fminf(fmaxf(p.action[(size_t)idxq * 3 + j], -1.0f), 1.0f);
        plan_tgt[0] = sq[4] + a3[0] * 0.3f;
        plan_tgt[1] = sq[5] + a3[1] * 0.3f;
        plan_tgt[2] = sq[6] + a3[2] * 30.0f;
#pragma unroll
        for (int j = 0; j < kPidRows; ++j) pid[j] = p.pid[(size_t)j * ld + idxq];
      }
    }
    const int nsub = SUBSTEPS ? p.n_sub : 1;

    if (!first_iter && threadIdx.x == 0) bulk_wait_read0();   // the tile may still be being read by the previous bulk store
    first_iter = false;

    NP_COOP_STAMP(0);
#pragma unroll 1
    for (int sub = 0; sub < nsub; ++sub) {
    if constexpr (SUBSTEPS) {
      if (owner) {   // the PID stack (ctrl_device.cuh) and the control lag of this FDM sub-step
        float a4[4];
        if constexpr (PLAN) pid_controller(sq, c.airspeed, c.dt, plan_tgt[0], plan_tgt[1], plan_tgt[2], pid, p.pid_first != 0 && sub == 0, a4);
        else combat_controller(sq, c.airspeed, c.dt, a_cmd, pid, p.pid_first != 0 && sub == 0, a4);
#pragma unroll
        for (int j = 0; j < 4; ++j) a4[j] = fminf(fmaxf(a4[j], -1.0f), 1.0f);
        uq[0] = 0.9f * uq[0] + 0.1f * a4[0] * 0.225f * 76300.0f / DC(0.3048f);
        uq[1] = 0.9f * uq[1] + 0.1f * a4[1] * 45.0f;
        uq[2] = 0.9f * uq[2] + 0.1f * a4[2] * 45.0f;
        uq[3] = 0.9f * uq[3] + 0.1f * a4[3] * 45.0f;
        xe[(q1 ? kCoopPairs : 0) + lane] = uq[1];
      }
      __syncthreads();   // the new elevator of both aircraft; the verdict of the last sub-step has read its slots
      el = make_float2(xe[lane], xe[kCoopPairs + lane]);
    }
    for (int pass = 0; pass < 2; ++pass) {
      const float2 adeg = make_float2(al.x * kR2D, al.y * kR2D);
      const float2 bdeg = make_float2(be.x * kR2D, be.y * kR2D);
      const uint32_t wb = opaque_u32(wb0);
      ZIn2 zi;
      zscores_ab2(blob, adeg, bdeg, zi);
      zscores_el2(blob, el, zi);
      const CoopRange* share = kCoopShareDev[V][warp];
      // after the first sub-step the slots already hold the outputs at the current (alpha, beta): pass 1 left them there
      if (pass == 1 || (miss && sub == 0)) {
        eval_group2<kCy>(blob, wb, zi, coef2, kCoopPairs, share[0].count, share[0].first);
        eval_group2<kdCx_lef>(blob, wb, zi, coef2, kCoopPairs, share[1].count, share[1].first);
        eval_group2<kdCz_lef>(blob, wb, zi, coef2, kCoopPairs, share[2].count, share[2].first);
        eval_group2<kdCy_r30>(blob, wb, zi, coef2, kCoopPairs, share[3].count, share[3].first);
        eval_group2<kdCy_a20>(blob, wb, zi, coef2, kCoopPairs, share[4].count, share[4].first);
        eval_group2<kdCy_a20_lef>(blob, wb, zi, coef2, kCoopPairs, share[5].count, share[5].first);
      }
      eval_el3_nets(blob, wb, zi, coef2, kCoopPairs, share[6 + pass].count, share[6 + pass].first);
      if (pass == 0 && warp == coop_eta_warp(V)) coef2[kEtaEl * kCoopPairs] = eta_el2(tabs, el);
      // the part of the tail that needs no MLP output, done by the (lightly loaded) owner warps while the others finish
      const float aq = q1 ? adeg.y : adeg.x;
      Trig g;
      float tp = 0.0f, a1[kNumA1];
      if (owner) {
        uint32_t seg, seg_unused;
        pwl_search2<kLevelsA>(tabs.bp_a, aq, aq, seg, seg_unused);
        g = make_trig(sq);
        tp = tfac_pow(sq[2]);
        alpha_coefs<kNumUsed - kFirstA1>(blob, tabs, seg, aq, a1);
        if (!COMBAT && pass == 1 && sub == nsub - 1) {   // the observation of the new state (env_base.py:103) needs no coefficient at all
          float o[NP_NUM_OBS];
          make_obs(c, TASK, sq, uq, tq, g, eas2tas_of(tp), o);
          add_obs_noise(p, idxq, o, rng);
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
        }
      }
      NP_COOP_STAMP(pass == 0 ? 1 : 5);
      __syncthreads();   // all 22 slots of the 32 pairs are in place
      NP_COOP_STAMP(pass == 0 ? 2 : 5 + 3);

      if (pass == 1 && use_cache && act[0] && sub == nsub - 1) {   // the next step's Euler derivative needs exactly these
#pragma unroll
        for (int k = 0; k < kNumAB2; ++k)   // warps 0 / 1 have the tail to fly: the others store
          if (2 + k % (NW - 2) == warp) store_pair(p.cache + (size_t)k * ld, pr, coef2[(kFirstAB2 + k) * kCoopPairs], act[1]);
        if (warp == 2) store_pair(p.cache + (size_t)kNumAB2 * ld, pr, al, act[1]);
        if (warp == 3) store_pair(p.cache + (size_t)(kNumAB2 + 1) * ld, pr, be, act[1]);
      }

      if (owner) {
        const ForcePart fp = force_part(sq, uq[0], uq[2], uq[3], 0.0f, g, tp, cq, CS, a1);
        if (pass == 0) {
          float xdot[12];
          nlplant_kin_moments(sq, uq[2], uq[3], 0.0f, g, fp.qbar, fp.vt, fp.b, fp.t, cq, CS, a1, xdot);
          xdot[6] = fp.f.vt_dot; xdot[7] = fp.f.alpha_dot; xdot[8] = fp.f.beta_dot;
          const float h = c.dt - 0.0f;
          const bool frozen = PLAN && (badq || doneq);  // planning_env.py:162-166: s <- recent_s (u keeps filtering)
#pragma unroll
          for (int j = 0; j < 12; ++j) sq[j] = frozen ? sq[j] : sq[j] + h * xdot[j];
          stepq += 1;
          xa[(q1 ? kCoopPairs : 0) + lane] = sq[7];
          xb[(q1 ? kCoopPairs : 0) + lane] = sq[8];
          if constexpr (COMBAT) {
#pragma unroll
            for (int j = 0; j < 3; ++j) xp[((q1 ? 3 : 0) + j) * kCoopPairs + lane] = sq[j];
          }
        } else {
          const Verdict v = judge_state<COMBAT, TASK>(c, sq, tq, g, fp.f, stepq);
          excq |= v.exc; badq |= v.bad; doneq |= v.done;                           // OR-accumulated over the sub-steps
          rewq = v.rw + (float)(-200 * (int)badq + 200 * (int)doneq);
          causesq |= actq ? v.causes : 0;
          if constexpr (COMBAT) {   // pair conditions: Crash (crash.py:29-42) and Shutdown (shutdown.py:30-40); ego - enemy, as K1
            const int po = (q1 ? 0 : 3) * kCoopPairs + lane;   // the partner's position after this sub-step
            const float pn = xp[po], pe = xp[po + kCoopPairs], pa = xp[po + 2 * kCoopPairs];
            const float dn0 = q1 ? pn - sq[0] : sq[0] - pn, de0 = q1 ? pe - sq[1] : sq[1] - pe, da0 = q1 ? pa - sq[2] : sq[2] - pa;
            const bool crash = (dn0 * dn0 + de0 * de0 + da0 * da0) <= c.distance_limit * c.distance_limit;
            const bool m1 = blood[0] <= 0.0f, m2 = blood[1] <= 0.0f;
            const bool sd_done = m2 && !m1;
            badq |= crash | m1;
            doneq |= sd_done;
            causesq |= actq ? (((int)(crash | m1) << 5) | ((int)sd_done << 6)) : 0;
          }
        }
      }
      if (pass == 0) {
        NP_COOP_STAMP(3);
        __syncthreads();   // the new (alpha, beta) of both aircraft, for every warp's share of pass 1
        NP_COOP_STAMP(4);
        al = make_float2(xa[lane], xa[kCoopPairs + lane]);
        be = make_float2(xb[lane], xb[kCoopPairs + lane]);
      }
    }
    }  // sub-steps

    if constexpr (COMBAT) {   // warp 1 hands its final state to warp 0 (through the unused observation tile), which finishes the pair
      if (warp == 1) {
#pragma unroll
        for (int j = 0; j < 12; ++j) otile[j * kCoopPairs + lane] = sq[j];
      }
      __syncthreads();
      if (warp == 0) {
        float sp[2][12], rw[2];
#pragma unroll
        for (int j = 0; j < 12; ++j) { sp[0][j] = sq[j]; sp[1][j] = otile[j * kCoopPairs + lane]; }
        combat_outputs(p, sp, blood, rw, pr, act, nsub > 0);   // 15-D rows, rewards, blood (singlecombat_env.py:64-181,263-271)
        if (act[0]) {
          store_pair(p.blood, pr, make_float2(blood[0], blood[1]), act[1]);
          store_pair(p.reward, pr, make_float2(rw[0], rw[1]), act[1]);
        }
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
        if constexpr (!COMBAT) {
#pragma unroll
          for (int j = 0; j < 3; ++j) p.tgt[(size_t)j * ld + row] = tq[j];
          p.reward[row] = rewq;
        }
        if constexpr (SUBSTEPS) {
#pragma unroll
          for (int j = 0; j < kPidRows; ++j) p.pid[(size_t)j * ld + row] = pid[j];
        }
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
    NP_COOP_STAMP(6);
    __syncthreads();   // tile complete; slots and exchange rows free for the next 32 pairs
    NP_COOP_STAMP(7);
    if (staged) {      // tile -> obs[64 rows], 5 632 contiguous bytes
      float* dst = p.obs + (size_t)(2 * pbase) * NP_NUM_OBS;
      if (!p.obs_stg) {
        if (threadIdx.x == 0) {
          bulk_s2g(dst, otile, kObsTileFloats * 4);
          bulk_commit();
        }
      } else {
        const float4* src = reinterpret_cast<const float4*>(otile);
        for (int k = threadIdx.x; k < kObsTileFloats / 4; k += NW * 32) reinterpret_cast<float4*>(dst)[k] = src[k];
        __syncthreads();
      }
    }
  }
  if (threadIdx.x == 0) bulk_wait0();   // shared memory must outlive the last bulk store
  NP_COOP_STAMP_FLUSH();
}
