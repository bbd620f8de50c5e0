// libnplane.so -- kernels and C ABI (include/nplane.h) of the B200 F-16 flight-dynamics step.
//
// One persistent kernel launch per BaseEnv.step() (reference: envs/env_base.py:99-109):
//   masked episodic reset -> control low-pass -> nlplant(s,u') -> explicit Euler -> step_count -> 22-D obs
//   -> nlplant(s',u') force part for the Overload check -> six termination predicates -> reward -> stores.
// Layout: SoA state rows [F][ld]; every thread owns TWO adjacent aircraft, so each field access is one coalesced
// 8-byte-per-lane (256 B per warp) transaction and the MLP arithmetic runs on packed FFMA2 (two aircraft per issue
// slot).  The aero image (22 MLP weight sets + the piecewise-linear tables of the 21 one-input nets, aero_pack.h)
// is TMA-bulk-loaded into shared memory once per persistent CTA; weights are read as warp-broadcast LDS.128.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (see f16_device.cuh for why).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <memory>
#include <string>
#include <vector>

#include "../../include/nplane.h"
#include "aero_pack.h"
#include "f16_device.cuh"
#include "uav_device.cuh"
#include "ctrl_device.cuh"
#include "tables_device.cuh"
#include "ptx_device.cuh"
#include "env_device.cuh"
#include "combat_device.cuh"
#include "uav_kernels.cuh"
#include "aux_kernels.cuh"

using namespace npl;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define NP_CUDA(expr)                                                                           \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail(NP_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));               \
  } while (0)

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
constexpr int kCacheRows = kNumAB2 + 2;  // 16 (alpha,beta)-MLP outputs + alpha key + beta key
// workspace rows of [ld] f32: coefficient cache | controller state (kPidRows, ctrl_device.cuh: 12) | blood | pair_reset (u8, one row)
constexpr int kWorkspaceRows = kCacheRows + 12 + 2;

struct np_aero {
  uint32_t* image_dev = nullptr;
  int bytes = 0;  // multiple of 16
  int device = 0;
};

struct np_tables {
  float* image_dev = nullptr;  // kTablesFloats (tables_device.cuh)
  int device = 0;
};

// Every entry point works on the device its handle (or its pointers) lives on, whatever the caller's current device is
// (the reference accepts any device= transparently): switch for the duration of the call, restore on the way out.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
// device that owns a device pointer (handle-less entry points); -1 = leave the current device alone
static int device_of(const void* p) {
  cudaPointerAttributes a;
  if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}

struct np_env {
  np_env_cfg cfg;
  const np_aero* aero = nullptr;
  const np_tables* tables = nullptr;  // set instead of aero: the TABLE aero back-end
  np_buffers buf;
  bool bound = false;
  uint32_t step_index = 0;
  bool pid_started = false;  // the fused PID controller has run at least once (PID.reset, pid.py:13)
  int device = 0;            // the device the env was created on; every entry point switches to it
  int obs_stg = 0;           // NPLANE_OBS_STORE=stg: per-lane stores of the staged observation tile instead of the TMA bulk store
  uint8_t* mirror = nullptr; // set by np_env_step_mapped around a step: second copy of the new flags in mapped host memory
  int mirror_ld = 0;
  int tab_pairs = 0;         // NPLANE_TAB_KERNEL=pairs: the table back-end on K1's two-aircraft-per-thread kernel (round 1) instead of K1t
  int block = 0;             // 0: chosen per launch (pick_block); else forced by NPLANE_BLOCK
  int coop_pairs = 0;        // ranges of up to this many pairs run on K1c (NPLANE_COOP_PAIRS; 0 = never)
  int coop_warps = 0;        // NPLANE_COOP_WARPS = 4 / 8: force K1c's CTA shape (0: by population)
  int coop_grid = 0;         // NPLANE_COOP_GRID: cap on K1c's grid (experiments: how much of a step is cold start)
  int pdl = 1;               // K1c is launched with programmatic stream serialisation (NPLANE_PDL=0: ordinary launch)
  int tab_block = 384, grid = 0, smem = 0, num_sms = 0, last_block = 384;
  // np_env_step_host: one in-order stream per engine (upload, kernels, download) and the events chaining them
  static constexpr int kMaxHostChunks = 16;
  cudaStream_t hs[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t hev_start = nullptr, hev_done = nullptr, hev_up[kMaxHostChunks] = {}, hev_run[kMaxHostChunks] = {};
};

// ------------------------------------------------------------------------------------------------
// K1: the fused step kernel
// ------------------------------------------------------------------------------------------------
// Observation rows are 88 B: a thread's two rows are staged in a per-warp tile (64 rows = 5 632 B, contiguous in the
// row-major obs array) and leave as eleven fully coalesced 16-byte-per-lane stores (512 B per warp instruction, whole
// 128-B lines).  For HBM this replaces 22 strided 8-byte stores per thread; for MAPPED HOST memory (np_env_step_mapped)
// it is what makes the PCIe writes full-line posted writes.  Block sizes above 384 have no room for the tiles next to the
// aero image and the coefficient slots and store their rows directly.
constexpr int kObsTileFloats = 64 * NP_NUM_OBS;
#ifdef NPLANE_NO_STAGE
constexpr bool stage_obs(int, int) { return false; }
#else
constexpr bool stage_obs(int bs, int mode) { return bs <= 384 && mode != 2 /* MODE_COMBAT: 15-D rows, written by combat_outputs */; }
#endif
static int step_smem_bytes(int aero_bytes, int bs, int mode, bool tab = false) {
  return aero_bytes + (tab ? 0 : kNumSlots * bs * 8) + (stage_obs(bs, mode) ? (bs / 32) * kObsTileFloats * 4 : 0) + 16;
}

// MODE_STEP  : BaseEnv.step (one FDM step driven by the caller's 4-D action).
// MODE_COMBAT: SingleCombatEnv.step (singlecombat_env.py:240-274): the thread's two aircraft ARE the pair (ego = 2e,
//               enemy = 2e + 1), so the relative geometry needs no exchange; env-level reset, 5 FDM sub-steps under the
//               attitude-demand controller, 15-D obs, AO/TA/range reward, blood model, Crash / Shutdown / Timeout.
// MODE_PLAN  : PlanningEnv.step (planning_env.py:144-177): the caller's 3-D action sets pitch / heading / speed targets
//               that the fused PID controller (ctrl_device.cuh) tracks for n_sub FDM sub-steps; aircraft that terminate
//               inside the env step are frozen (s <- recent_s); state stays in registers across the sub-steps.
enum { MODE_STEP = 0, MODE_PLAN = 1, MODE_COMBAT = 2 };
// TAB = true: the TABLE aero back-end (SURVEY f-3, tables_device.cuh).  The staged image is the 54 KB of NASA tables
// instead of the MLP image; every coefficient is a multilinear interpolation into registers, so the shared coefficient
// slots and the (alpha, beta) cache do not exist.  Everything around the coefficients is the same code.

template <int BS, int MINB, int MODE, bool TAB = false, int TASK = NP_TASK_HEADING>
__global__ void __launch_bounds__(BS, MINB) f16_step_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* blob = reinterpret_cast<float*>(smem_raw);
  float2* coef_all = reinterpret_cast<float2*>(smem_raw + p.aero_bytes);     // [kNumSlots][BS] float2 (MLP back-end)
  constexpr bool STAGE = stage_obs(BS, MODE);
  float* otile = reinterpret_cast<float*>(coef_all + (TAB ? 0 : kNumSlots * BS)) + (threadIdx.x >> 5) * kObsTileFloats;  // this warp's tile
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(coef_all + (TAB ? 0 : kNumSlots * BS)) +
                                              (STAGE ? (BS / 32) * kObsTileFloats : 0));

  // the RNG counter of this launch (host counter + device epoch, env_device.cuh rng_step): one L2 read per CTA, parked in
  // shared memory (a read per aircraft measured -2.5 %; a register held through the kernel is what K1 has none of to spare)
  uint32_t& rng_s = reinterpret_cast<uint32_t*>(bar)[2];   // the spare word behind the staging barrier
  if (threadIdx.x == 32) rng_s = rng_step(p);
  stage_aero(blob, p.aero, (uint32_t)p.aero_bytes, bar);
  uint32_t wb0 = 0;
  AeroTabs tabs{};
  const float* c0 = blob;
  ZeroCells zc{};
  if constexpr (TAB) {
    zc = zero_cells(blob);
  } else {
    wb0 = aero_base_after_staging(blob);
    tabs = aero_tabs(blob, wb0);
    c0 = blob + reinterpret_cast<const int32_t*>(blob)[kHdrC0];
  }

  const np_env_cfg& c = p.cfg;
  const int n = c.n, ld = c.ld;
  const int npairs = (n + 1) >> 1;
  float2* coef2 = coef_all + threadIdx.x;                       // slot k of this thread's pair: coef2[k * BS]
  float* cf = reinterpret_cast<float*>(coef2);                  // aircraft a, slot k: cf[a + k * 2 * BS]
  constexpr int CS = TAB ? 1 : 2 * BS;
  constexpr bool PLAN = MODE == MODE_PLAN, COMBAT = MODE == MODE_COMBAT;
  const bool use_cache = !TAB && c.use_coef_cache != 0;

  const int pend = p.pair_end < npairs ? p.pair_end : npairs;
  for (int pbase = p.pair_begin + blockIdx.x * BS; pbase < pend; pbase += gridDim.x * BS) {
    const int pr = pbase + threadIdx.x;
    const int prl = pr < pend ? pr : pend - 1;  // inactive lanes shadow the last pair and never store
    const bool act[2] = {pr < pend && 2 * pr < n, pr < pend && 2 * pr + 1 < n};
    const int idx[2] = {min(2 * prl, n - 1), min(2 * prl + 1, n - 1)};  // row index into [n][*] arrays
    // the warp's 64 observation rows leave through its staging tile when all of them exist and the destination is 16-B aligned
    const bool staged = STAGE && __all_sync(0xffffffffu, act[1]) && ((reinterpret_cast<uintptr_t>(p.obs) & 15) == 0);

    // ---- load ----------------------------------------------------------------------------------
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
    // cached (alpha,beta)-MLP outputs and their key: requested with the state rows (one memory round trip, not three)
    // and parked straight in this thread's coefficient slots; reset lanes / misses overwrite them below
    float2 ka = make_float2(0.f, 0.f), kb = ka;
    if (use_cache) {
      ka = reinterpret_cast<const float2*>(p.cache + (size_t)kNumAB2 * ld)[prl];
      kb = reinterpret_cast<const float2*>(p.cache + (size_t)(kNumAB2 + 1) * ld)[prl];
#pragma unroll
      for (int k = 0; k < kNumAB2; ++k) coef2[(kFirstAB2 + k) * BS] = reinterpret_cast<const float2*>(p.cache + (size_t)k * ld)[prl];
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (PLAN) {
        a[q][0] = p.action[(size_t)idx[q] * 3 + 0]; a[q][1] = p.action[(size_t)idx[q] * 3 + 1];
        a[q][2] = p.action[(size_t)idx[q] * 3 + 2]; a[q][3] = 0.0f;
      } else {
        const float4 av = reinterpret_cast<const float4*>(p.action)[idx[q]];
        a[q][0] = av.x; a[q][1] = av.y; a[q][2] = av.z; a[q][3] = av.w;
      }
    }

    // ---- episodic reset of terminated aircraft (env_base.py:83-97) --------------------------------------------
    float blood[2] = {0.f, 0.f};
    if (COMBAT) {  // env-level reset (singlecombat_env.py:207-238): either flag re-initialises the whole pair
      const float2 bv = reinterpret_cast<const float2*>(p.blood)[prl];
      blood[0] = bv.x; blood[1] = bv.y;
      if (p.records) {   // role-sharded: the two lanes are different envs; their env-level flags were OR-ed by the pair kernel
        const uchar2 pv = reinterpret_cast<const uchar2*>(p.pair_reset)[prl];
        rst[0] = pv.x != 0; rst[1] = pv.y != 0;
      } else {
        bool r = rst[0] || rst[1];
        // MultipleCombat (multiplecombat_env.py:207-238): an env is TWO adjacent duels = two adjacent lanes; any flag resets all four
        if (c.combat_pairs_per_env == 2) r |= __shfl_xor_sync(0xffffffffu, (int)r, 1) != 0;
        rst[0] = rst[1] = r;
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (rst[q]) {
        const Draws r = reset_draws(p, idx[q], rng_s);
        if (COMBAT) {
          combat_reset_aircraft(c, r, s[q], u[q]);
          blood[q] = 100.0f;
        } else {
          reset_aircraft(c, TASK, r, s[q], u[q], tgt[q]);
        }
        steps[q] = 0;
      }
    }
    // planning step: targets from the high-level action (planning_env.py:146-152); controller state of both modes
    float plan_tgt[2][3], pid[2][kPidRows];
    if (COMBAT) {
#pragma unroll
      for (int j = 0; j < kPidRows; ++j) {
        const float2 v = reinterpret_cast<const float2*>(p.pid + (size_t)j * ld)[prl];
        pid[0][j] = v.x; pid[1][j] = v.y;
      }
    }
    float a_cmd[2][4];   // the caller's clamped 4-D action, kept for all sub-steps (App. D.10 fixed)
    if (COMBAT) {
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) a_cmd[q][j] = fminf(fmaxf(a[q][j], -1.0f), 1.0f);
    }
    if (PLAN) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int j = 0; j < 3; ++j) a[q][j] = fminf(fmaxf(a[q][j], -1.0f), 1.0f);
        plan_tgt[q][0] = s[q][4] + a[q][0] * 0.3f;
        plan_tgt[q][1] = s[q][5] + a[q][1] * 0.3f;
        plan_tgt[q][2] = s[q][6] + a[q][2] * 30.0f;
      }
#pragma unroll
      for (int j = 0; j < kPidRows; ++j) {
        const float2 v = reinterpret_cast<const float2*>(p.pid + (size_t)j * ld)[prl];
        pid[0][j] = v.x; pid[1][j] = v.y;
      }
    }
    count_cause2(p.counters, 7, rst[0] && act[0], rst[1] && act[1]);

    // ---- (alpha,beta)-MLP outputs at (s): cache hit (the Overload evaluation of the previous step was at exactly
    //      this alpha, beta), a reset lane (constants for alpha = beta = 0), or a miss -> the warp evaluates ----
    bool miss = false;
    if constexpr (!TAB) {
      bool hit[2] = {rst[0], rst[1]};
      if (use_cache) {
        hit[0] |= __float_as_uint(ka.x) == __float_as_uint(s[0][7]) && __float_as_uint(kb.x) == __float_as_uint(s[0][8]);
        hit[1] |= __float_as_uint(ka.y) == __float_as_uint(s[1][7]) && __float_as_uint(kb.y) == __float_as_uint(s[1][8]);
      }
      miss = __any_sync(0xffffffffu, !(hit[0] && hit[1])) != 0;
      if (!miss && (rst[0] || rst[1])) {  // a reset lane sits at alpha = beta = 0: constants from the aero image
#pragma unroll 4
        for (int k = 0; k < kNumAB2; ++k) {
          float2 v = coef2[(kFirstAB2 + k) * BS];
          if (rst[0]) v.x = c0[k];
          if (rst[1]) v.y = c0[k];
          coef2[(kFirstAB2 + k) * BS] = v;
        }
      }
    }

    // ---- two passes over the same code: pass 0 = Euler derivative at (s, u'), then obs of the new state;
    //      pass 1 = force equations at (s', u') for the Overload check, terminations, reward ------------------
    bool bad[2] = {false, false}, done[2] = {false, false};   // flags were cleared by the reset above (env_base.py:93-95)
    float rew[2];
    int causes[2] = {0, 0};
    const int nsub = (PLAN || COMBAT) ? p.n_sub : 1;
    bool exc[2] = {false, false};
#pragma unroll 1
    for (int sub = 0; sub < nsub; ++sub) {
    // ---- action of this FDM step + control low-pass (F16_model.py:52-57) ------------------------------------
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (PLAN) pid_controller(s[q], c.airspeed, c.dt, plan_tgt[q][0], plan_tgt[q][1], plan_tgt[q][2], pid[q],
                               p.pid_first != 0 && sub == 0, a[q]);
      if (COMBAT) combat_controller(s[q], c.airspeed, c.dt, a_cmd[q], pid[q], p.pid_first != 0 && sub == 0, a[q]);
#pragma unroll
      for (int j = 0; j < 4; ++j) a[q][j] = fminf(fmaxf(a[q][j], -1.0f), 1.0f);
      u[q][0] = 0.9f * u[q][0] + 0.1f * a[q][0] * 0.225f * 76300.0f / DC(0.3048f);
      u[q][1] = 0.9f * u[q][1] + 0.1f * a[q][1] * 45.0f;
      u[q][2] = 0.9f * u[q][2] + 0.1f * a[q][2] * 45.0f;
      u[q][3] = 0.9f * u[q][3] + 0.1f * a[q][3] * 45.0f;
    }
#ifdef NPLANE_UNROLL_PASS
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int pass = 0; pass < 2; ++pass) {
      const float2 adeg = make_float2(s[0][7] * kR2D, s[1][7] * kR2D);
      const float2 bdeg = make_float2(s[0][8] * kR2D, s[1][8] * kR2D);
      uint32_t seg[2] = {0u, 0u};
      if constexpr (!TAB) {
      const uint32_t wb = opaque_u32(wb0);
      ZIn2 zi;
      zscores_ab2(blob, adeg, bdeg, zi);
      zscores_el2(blob, make_float2(u[0][1], u[1][1]), zi);
      // after the first sub-step the slots already hold the outputs at the current (alpha, beta): pass 1 left them there
      if (pass == 1 || (miss && sub == 0)) eval_ab2_nets(blob, wb, zi, coef2, BS);
      eval_el3_nets(blob, wb, zi, coef2, BS, pass == 0 ? 5 : 2);
      if (pass == 1 && use_cache && act[0] && sub == nsub - 1) {  // the next step's Euler derivative needs exactly these
#pragma unroll 4
        for (int k = 0; k < kNumAB2; ++k) store_pair(p.cache + (size_t)k * ld, pr, coef2[(kFirstAB2 + k) * BS], act[1]);
        store_pair(p.cache + (size_t)kNumAB2 * ld, pr, make_float2(s[0][7], s[1][7]), act[1]);
        store_pair(p.cache + (size_t)(kNumAB2 + 1) * ld, pr, make_float2(s[0][8], s[1][8]), act[1]);
      }
      pwl_search2<kLevelsA>(tabs.bp_a, adeg.x, adeg.y, seg[0], seg[1]);
      if (pass == 0) coef2[kEtaEl * BS] = eta_el2(tabs, make_float2(u[0][1], u[1][1]));
      }

#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float* sq = s[q];
        const float* uq = u[q];
        const float* tq = tgt[q];
        const float* cq = cf + q;
        const Trig g = make_trig(sq);
        const float tp = tfac_pow(sq[2]);
        float a1[kNumA1];
        float ctab[TAB ? kNumSlots : 1];
        if constexpr (TAB) {
          table_env_coefs(blob, zc, q == 0 ? adeg.x : adeg.y, q == 0 ? bdeg.x : bdeg.y, uq[1], pass == 0, ctab, a1);
          cq = ctab;
        } else {
          alpha_coefs<kNumUsed - kFirstA1>(blob, tabs, seg[q], q == 0 ? adeg.x : adeg.y, a1);
        }
        const ForcePart fp = force_part(sq, uq[0], uq[2], uq[3], 0.0f, g, tp, cq, CS, a1);

        if (pass == 0) {
          // ---- Euler step (F16_model.py:64-67; torchdiffeq fixed-grid euler on t=[0,dt]) ---------------
          float xdot[12];
          nlplant_kin_moments(sq, uq[2], uq[3], 0.0f, g, fp.qbar, fp.vt, fp.b, fp.t, cq, CS, a1, xdot);
          xdot[6] = fp.f.vt_dot; xdot[7] = fp.f.alpha_dot; xdot[8] = fp.f.beta_dot;
          const float h = c.dt - 0.0f;
          const bool frozen = PLAN && (bad[q] || done[q]);  // planning_env.py:162-166: s <- recent_s (u keeps filtering)
#pragma unroll
          for (int j = 0; j < 12; ++j) sq[j] = frozen ? sq[j] : sq[j] + h * xdot[j];
          steps[q] += 1;  // env_base.py:102
        } else {
          // ---- observation of the new state (env_base.py:103): pass 1's trig / atmosphere terms are exactly the ones
          //      it needs, so it is produced here rather than with a third evaluation at the end of pass 0;
          //      planning / combat return only the last sub-step's observation (planning_env.py:177) -------------
          if (!COMBAT && (!PLAN || sub == nsub - 1)) {
            float o[NP_NUM_OBS];
            make_obs(c, TASK, sq, uq, tq, g, eas2tas_of(tp), o);
            add_obs_noise(p, idx[q], o, rng_s);
            if (staged) {
              if (q == 0 && !p.obs_stg) {   // the tile's previous contents may still be being read by the last bulk store
                if ((threadIdx.x & 31) == 0) bulk_wait_read0();
                __syncwarp();
              }
              float2* orow = reinterpret_cast<float2*>(otile + (2 * (threadIdx.x & 31) + q) * NP_NUM_OBS);
#pragma unroll
              for (int j = 0; j < NP_NUM_OBS / 2; ++j) orow[j] = make_float2(o[2 * j], o[2 * j + 1]);
            } else if (act[q]) {
              float2* orow = reinterpret_cast<float2*>(p.obs + (size_t)(2 * pr + q) * NP_NUM_OBS);  // 88-B rows: 8-B aligned
#pragma unroll
              for (int j = 0; j < NP_NUM_OBS / 2; ++j) orow[j] = make_float2(o[2 * j], o[2 * j + 1]);
            }
          }
          // ---- terminations (task_base.py:75-96) and the task reward on the new state ----------------------
          const Verdict v = judge_state<COMBAT, TASK>(c, sq, tq, g, fp.f, steps[q]);
          exc[q] |= v.exc;
          bad[q] |= v.bad;                                                      // OR-accumulated (env_base.py:70-75)
          done[q] |= v.done;
          rew[q] = v.rw + (float)(-200 * (int)bad[q] + 200 * (int)done[q]);     // event_driven_reward.py:28
          causes[q] |= act[q] ? v.causes : 0;
        }
      }
      if (STAGE && pass == 1 && staged && (!PLAN || sub == nsub - 1)) {   // tile -> obs[64 rows], 5 632 contiguous bytes
        const int lane = threadIdx.x & 31;
        float* dst = p.obs + (size_t)(2 * (pr - lane)) * NP_NUM_OBS;
        if (!p.obs_stg) {            // ONE TMA bulk store per warp: no LSU instructions, large write bursts
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            bulk_s2g(dst, otile, kObsTileFloats * 4);
            bulk_commit();
          }
        } else {                     // 11 x 512 B per warp through the LSU
          __syncwarp();
          const float4* src = reinterpret_cast<const float4*>(otile);
#pragma unroll
          for (int k = 0; k < kObsTileFloats / 128; ++k) reinterpret_cast<float4*>(dst)[lane + 32 * k] = src[lane + 32 * k];
          __syncwarp();
        }
      }
      if (COMBAT && pass == 1 && p.records) {   // role-sharded: the partner is on another rank; the pair kernel checks Crash
        if (sub < nsub - 1 && sub < kCombatMaxSub - 1) {   // against the partner's position after the same sub-step
#pragma unroll
          for (int q = 0; q < 2; ++q)
            if (act[q]) {
              float* row = p.records + (size_t)(2 * pr + q) * kCombatRecFloats + 16 + 3 * sub;
              row[0] = s[q][0]; row[1] = s[q][1]; row[2] = s[q][2];
            }
        }
      } else if (COMBAT && pass == 1) {   // pair conditions: Crash (crash.py:29-42) and Shutdown (shutdown.py:30-40)
        const float dn0 = s[0][0] - s[1][0], de0 = s[0][1] - s[1][1], da0 = s[0][2] - s[1][2];
        const bool crash = (dn0 * dn0 + de0 * de0 + da0 * da0) <= c.distance_limit * c.distance_limit;
        const bool m1 = blood[0] <= 0.0f, m2 = blood[1] <= 0.0f;
        const bool sd_done = m2 && !m1;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          bad[q] |= crash | m1;
          done[q] |= sd_done;
          causes[q] |= act[q] ? (((int)(crash | m1) << 5) | ((int)sd_done << 6)) : 0;
        }
      }
    }

    }  // sub-steps

    if (COMBAT && p.records) {   // role-sharded: publish this aircraft's record; obs / reward / blood / pair flags follow in
      rew[0] = rew[1] = 0.0f;    // combat_pair_kernel once the partner's record is visible
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (act[q])
          combat_rec_store(p.records + (size_t)(2 * pr + q) * kCombatRecFloats, combat_rec(s[q]), blood[q],
                           (done[q] ? 1 : 0) | (bad[q] ? 2 : 0) | (exc[q] ? 4 : 0));
    } else if (COMBAT) {
      combat_outputs(p, s, blood, rew, pr, act, nsub > 0);
    }

    // ---- termination-cause counters (replace the reference's per-condition print(torch.sum(...)) syncs) --------
#pragma unroll
    for (int w = 0; w < 7; ++w) count_cause2(p.counters, w, (causes[0] >> w) & 1, (causes[1] >> w) & 1);

    // ---- store ----------------------------------------------------------------------------------
    if (act[0]) {  // a half-active tail pair (odd n) stores its first aircraft only
#pragma unroll
      for (int j = 0; j < 12; ++j) store_pair(p.s + (size_t)j * ld, pr, make_float2(s[0][j], s[1][j]), act[1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) store_pair(p.u + (size_t)j * ld, pr, make_float2(u[0][j], u[1][j]), act[1]);
      if (!COMBAT) {
#pragma unroll
        for (int j = 0; j < 3; ++j) store_pair(p.tgt + (size_t)j * ld, pr, make_float2(tgt[0][j], tgt[1][j]), act[1]);
      } else {
        store_pair(p.blood, pr, make_float2(blood[0], blood[1]), act[1]);
      }
      if (!(COMBAT && p.records)) store_pair(p.reward, pr, make_float2(rew[0], rew[1]), act[1]);
      if (PLAN || COMBAT) {
#pragma unroll
        for (int j = 0; j < kPidRows; ++j) store_pair(p.pid + (size_t)j * ld, pr, make_float2(pid[0][j], pid[1][j]), act[1]);
      }
      if (act[1]) {
        reinterpret_cast<int2*>(p.step_count)[pr] = make_int2(steps[0], steps[1]);
        reinterpret_cast<uchar2*>(p.flags)[pr] = make_uchar2(done[0] ? 1 : 0, done[1] ? 1 : 0);
        reinterpret_cast<uchar2*>(p.flags + ld)[pr] = make_uchar2(bad[0] ? 1 : 0, bad[1] ? 1 : 0);
        reinterpret_cast<uchar2*>(p.flags + 2 * (size_t)ld)[pr] = make_uchar2(exc[0] ? 1 : 0, exc[1] ? 1 : 0);  // control tasks: Timeout is commented out (heading_task.py:45)
        if (p.flags_mirror) {   // the numpy boundary's copy, written straight into mapped host memory
          const size_t ml = (size_t)p.flags_mirror_ld;
          reinterpret_cast<uchar2*>(p.flags_mirror)[pr] = make_uchar2(done[0] ? 1 : 0, done[1] ? 1 : 0);
          reinterpret_cast<uchar2*>(p.flags_mirror + ml)[pr] = make_uchar2(bad[0] ? 1 : 0, bad[1] ? 1 : 0);
          reinterpret_cast<uchar2*>(p.flags_mirror + 2 * ml)[pr] = make_uchar2(exc[0] ? 1 : 0, exc[1] ? 1 : 0);
        }
      } else {
        p.step_count[2 * pr] = steps[0];
        p.flags[2 * pr] = done[0] ? 1 : 0;
        p.flags[ld + 2 * pr] = bad[0] ? 1 : 0;
        p.flags[2 * (size_t)ld + 2 * pr] = exc[0] ? 1 : 0;
        if (p.flags_mirror) {
          const size_t ml = (size_t)p.flags_mirror_ld;
          p.flags_mirror[2 * pr] = done[0] ? 1 : 0;
          p.flags_mirror[ml + 2 * pr] = bad[0] ? 1 : 0;
          p.flags_mirror[2 * ml + 2 * pr] = exc[0] ? 1 : 0;
        }
      }
    }
  }
  if (STAGE && (threadIdx.x & 31) == 0) bulk_wait0();   // shared memory must outlive the last bulk stores
}

// ------------------------------------------------------------------------------------------------
// K1t: BaseEnv.step with the TABLE aero back-end, ONE aircraft per thread.  The table step has no MLP arithmetic to pack
// two aircraft into (FFMA2), so the pair-per-thread shape of K1 only cost it registers: 168 -> 12 warps per SM, at 50 %
// issue utilisation on dependent shared-memory gathers.  Here: 384 threads, <= 85 registers, 2 CTAs = 24 warps per SM, the
// 54 KB table image staged once per CTA, coalesced 4-byte SoA accesses, the 88-byte observation rows staged per warp
// (32 rows = 2 816 contiguous bytes) and sent with one TMA bulk store.  Same device functions in the same order as K1's
// TAB instantiation: bit-identical outputs.
// ------------------------------------------------------------------------------------------------
#ifndef NPLANE_TAB_BS
#define NPLANE_TAB_BS 384     // x 2 CTAs per SM = 24 warps; 2 x (54 KB tables + 33 KB observation tiles) of shared memory
#define NPLANE_TAB_MINB 2
#endif
#ifdef NPLANE_TAB_NOSTAGE
constexpr bool kTabStage = false;
#else
constexpr bool kTabStage = true;
#endif
constexpr int kTabBS = NPLANE_TAB_BS, kTabMinB = NPLANE_TAB_MINB;
constexpr int kTabTileFloats = 32 * NP_NUM_OBS;
static int table_step_smem_bytes() { return kTablesFloats * 4 + (kTabStage ? (kTabBS / 32) * kTabTileFloats * 4 : 0) + 16; }

__device__ __forceinline__ void count_cause1(unsigned long long* counters, int which, bool p0) {
  const int k = __popc(__ballot_sync(0xffffffffu, p0));
  if (k != 0 && (threadIdx.x & 31) == 0) atomicAdd(&counters[which], (unsigned long long)k);
}

#include "coop_step_kernel.cuh"   // K1c: the small-population (latency) shape of K1

template <int TASK>
__global__ void __launch_bounds__(kTabBS, kTabMinB) f16_table_step_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* T = reinterpret_cast<float*>(smem_raw);
  float* otile = T + kTablesFloats + (threadIdx.x >> 5) * kTabTileFloats;
  uint64_t* bar = reinterpret_cast<uint64_t*>(T + kTablesFloats + (kTabStage ? (kTabBS / 32) * kTabTileFloats : 0));
  uint32_t& rng_s = reinterpret_cast<uint32_t*>(bar)[2];   // as in K1
  if (threadIdx.x == 32) rng_s = rng_step(p);
  stage_aero(T, p.aero, (uint32_t)(kTablesFloats * 4), bar);
  const ZeroCells zc = zero_cells(T);
  const np_env_cfg& c = p.cfg;
  const int n = c.n, ld = c.ld, lane = threadIdx.x & 31;
  const int i_begin = 2 * p.pair_begin, i_end = min(n, 2 * p.pair_end);
  for (int base = i_begin + blockIdx.x * kTabBS; base < i_end; base += gridDim.x * kTabBS) {
    const int i = base + threadIdx.x;
    const bool live = i < i_end;
    const int il = live ? i : i_end - 1;          // idle lanes shadow the last aircraft and never store
    float s[12], u[4], tgt[3], a[4];
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = p.s[(size_t)j * ld + il];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = p.u[(size_t)j * ld + il];
#pragma unroll
    for (int j = 0; j < 3; ++j) tgt[j] = p.tgt[(size_t)j * ld + il];
    int steps = p.step_count[il];
    const bool rst = (p.flags[il] | p.flags[ld + il] | p.flags[2 * (size_t)ld + il]) != 0;
    {
      const float4 av = reinterpret_cast<const float4*>(p.action)[il];
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
    }
    if (rst) {                                     // BaseEnv.reset (env_base.py:83-97)
      reset_aircraft(c, TASK, reset_draws(p, il, rng_s), s, u, tgt);
      steps = 0;
    }
    count_cause1(p.counters, 7, rst && live);
    lowpass_controls(a, u);                        // F16_model.py:52-57
    float rew = 0.0f;
    bool bad = false, done = false;
    int causes = 0;
    const bool staged = kTabStage && __all_sync(0xffffffffu, live) && ((reinterpret_cast<uintptr_t>(p.obs) & 15) == 0);
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      float ctab[kNumSlots], a1[kNumA1];
      table_env_coefs(T, zc, s[7] * kR2D, s[8] * kR2D, u[1], pass == 0, ctab, a1);
      const Trig g = make_trig(s);
      const float tp = tfac_pow(s[2]);
      const ForcePart fp = force_part(s, u[0], u[2], u[3], 0.0f, g, tp, ctab, 1, a1);
      if (pass == 0) {                             // Euler step (F16_model.py:64-67)
        float xdot[12];
        nlplant_kin_moments(s, u[2], u[3], 0.0f, g, fp.qbar, fp.vt, fp.b, fp.t, ctab, 1, a1, xdot);
        xdot[6] = fp.f.vt_dot; xdot[7] = fp.f.alpha_dot; xdot[8] = fp.f.beta_dot;
        const float h = c.dt - 0.0f;
#pragma unroll
        for (int j = 0; j < 12; ++j) s[j] = s[j] + h * xdot[j];
        steps += 1;                                // env_base.py:102
      } else {
        float o[NP_NUM_OBS];
        make_obs(c, TASK, s, u, tgt, g, eas2tas_of(tp), o);
        add_obs_noise(p, il, o, rng_s);
        if (staged) {
          if (lane == 0) bulk_wait_read0();        // the previous slab's bulk store has read the tile
          __syncwarp();
          float2* orow = reinterpret_cast<float2*>(otile + lane * NP_NUM_OBS);
#pragma unroll
          for (int j = 0; j < NP_NUM_OBS / 2; ++j) orow[j] = make_float2(o[2 * j], o[2 * j + 1]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            bulk_s2g(p.obs + (size_t)(i - lane) * NP_NUM_OBS, otile, kTabTileFloats * 4);
            bulk_commit();
          }
        } else if (live) {
          float2* orow = reinterpret_cast<float2*>(p.obs + (size_t)i * NP_NUM_OBS);
#pragma unroll
          for (int j = 0; j < NP_NUM_OBS / 2; ++j) orow[j] = make_float2(o[2 * j], o[2 * j + 1]);
        }
        const Verdict v = judge_state<false, TASK>(c, s, tgt, g, fp.f, steps);
        bad = v.bad; done = v.done;
        rew = v.rw + (float)(-200 * (int)bad + 200 * (int)done);       // event_driven_reward.py:28
        causes = live ? v.causes : 0;
      }
    }
#pragma unroll
    for (int w = 0; w < 7; ++w) count_cause1(p.counters, w, (causes >> w) & 1);
    if (live) {
#pragma unroll
      for (int j = 0; j < 12; ++j) p.s[(size_t)j * ld + i] = s[j];
#pragma unroll
      for (int j = 0; j < 4; ++j) p.u[(size_t)j * ld + i] = u[j];
#pragma unroll
      for (int j = 0; j < 3; ++j) p.tgt[(size_t)j * ld + i] = tgt[j];
      p.step_count[i] = steps;
      p.reward[i] = rew;
      p.flags[i] = done ? 1 : 0;
      p.flags[ld + i] = bad ? 1 : 0;
      p.flags[2 * (size_t)ld + i] = 0;             // control tasks: Timeout is commented out (heading_task.py:45)
      if (p.flags_mirror) {
        const size_t ml = (size_t)p.flags_mirror_ld;
        p.flags_mirror[i] = done ? 1 : 0;
        p.flags_mirror[ml + i] = bad ? 1 : 0;
        p.flags_mirror[2 * ml + i] = 0;
      }
    }
  }
  if (kTabStage && lane == 0) bulk_wait0();        // shared memory must outlive the last bulk stores
}

// ------------------------------------------------------------------------------------------------
// K3: standalone reset (BaseEnv.reset, env_base.py:83-97) -- HBM-bound, one aircraft per thread
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) f16_reset_kernel(const __grid_constant__ StepParams p) {
  const np_env_cfg& c = p.cfg;
  const int n = c.n, ld = c.ld;
  const float* c0 = p.tab ? nullptr : reinterpret_cast<const float*>(p.aero) + reinterpret_cast<const int32_t*>(p.aero)[kHdrC0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s[12], u[4], tgt[3];
    const bool rst = (p.flags[i] | p.flags[ld + i] | p.flags[2 * (size_t)ld + i]) != 0;
    if (rst) {
      const Draws r = reset_draws(p, i);
      reset_aircraft(c, c.task, r, s, u, tgt);
#pragma unroll
      for (int j = 0; j < 12; ++j) p.s[(size_t)j * ld + i] = s[j];
#pragma unroll
      for (int j = 0; j < 4; ++j) p.u[(size_t)j * ld + i] = u[j];
      p.u[(size_t)4 * ld + i] = 0.0f;
#pragma unroll
      for (int j = 0; j < 3; ++j) p.tgt[(size_t)j * ld + i] = tgt[j];
      p.step_count[i] = 0;
      if (c0) {  // the (alpha, beta)-MLP outputs at alpha = beta = 0 (the table back-end keeps no cache)
        for (int k = 0; k < kNumAB2; ++k) p.cache[(size_t)k * ld + i] = c0[k];
        p.cache[(size_t)kNumAB2 * ld + i] = 0.0f;
        p.cache[(size_t)(kNumAB2 + 1) * ld + i] = 0.0f;
      }
      atomicAdd(&p.counters[7], 1ull);
    } else {
#pragma unroll
      for (int j = 0; j < 12; ++j) s[j] = p.s[(size_t)j * ld + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = p.u[(size_t)j * ld + i];
#pragma unroll
      for (int j = 0; j < 3; ++j) tgt[j] = p.tgt[(size_t)j * ld + i];
    }
    p.flags[i] = 0; p.flags[ld + i] = 0; p.flags[2 * (size_t)ld + i] = 0;
    const Trig g = make_trig(s);
    float o[NP_NUM_OBS];
    make_obs(c, c.task, s, u, tgt, g, eas2tas_of(tfac_pow(s[2])), o);
    add_obs_noise(p, i, o);
#pragma unroll
    for (int j = 0; j < NP_NUM_OBS; ++j) p.obs[(size_t)i * NP_NUM_OBS + j] = o[j];
  }
}

// ------------------------------------------------------------------------------------------------
// K5r: the pair half of the ROLE-sharded combat step (singlecombat_env.py:64-181,207-238,263-271).  The local half
// (f16_step_kernel<MODE_COMBAT> with p.records) has flown this rank's aircraft -- all egos or all opponents, one per env --
// through the five sub-steps and published a record each; after a cross-device barrier this kernel pulls the partner's record
// from the peer rank's slab with its own loads over NVLink (PEER) or from an all-gathered array, and produces everything that
// needs both aircraft: Crash at every sub-step, Shutdown, the 15-D observation, the reward, the blood model, the final flags
// and the env-level reset flag of the next step.  One aircraft per thread; 112 B pulled per aircraft.
// ------------------------------------------------------------------------------------------------
struct PairParams {
  const float* own;        // [n][kCombatRecFloats]
  const float* partner;    // [n][kCombatRecFloats] (peer-mapped or gathered)
  float* obs;              // [n][15]
  float* reward;           // [n]
  float* blood;            // [ld]
  uint8_t* flags;          // [3][ld]
  uint8_t* pair_reset;     // [ld]
  unsigned long long* counters;
  int n, ld, role, n_sub;
  float distance_limit, reward_scale;
};
template <bool PEER>
__global__ void __launch_bounds__(256) combat_pair_kernel(const __grid_constant__ PairParams p) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
    const CombatRecFull own = combat_rec_load<false>(p.own + (size_t)i * kCombatRecFloats);
    const CombatRecFull oth = combat_rec_load<PEER>(p.partner + (size_t)i * kCombatRecFloats);
    const CombatRecFull& ego = p.role == 0 ? own : oth;
    const CombatRecFull& enm = p.role == 0 ? oth : own;
    // Crash (crash.py:29-42) after every sub-step, ego - enemy like the pair-sharded kernel
    bool crash = false;
    const float lim2 = p.distance_limit * p.distance_limit;
#pragma unroll
    for (int sub = 0; sub < kCombatMaxSub; ++sub) {     // constant trip count: the records stay in registers
      const bool last = sub == p.n_sub - 1;
      float d[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float pe = (sub < kCombatMaxSub - 1 && !last) ? ego.sub_pos[sub < kCombatMaxSub - 1 ? sub : 0][j] : ego.r.pos[j];
        const float pm = (sub < kCombatMaxSub - 1 && !last) ? enm.sub_pos[sub < kCombatMaxSub - 1 ? sub : 0][j] : enm.r.pos[j];
        d[j] = pe - pm;
      }
      crash |= (sub < p.n_sub) && (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) <= lim2;
    }
    const bool m1 = ego.blood <= 0.0f, m2 = enm.blood <= 0.0f;           // shutdown.py:30-40
    const bool sd_done = m2 && !m1;
    const bool stepped = p.n_sub > 0;
    const int pairbits = stepped ? ((sd_done ? 1 : 0) | ((crash | m1) ? 2 : 0)) : 0;
    const int mine = own.bits | pairbits, theirs = oth.bits | pairbits;
    const CombatGeo g = combat_geo(ego.r, enm.r);
    float o[NP_NUM_OBS_COMBAT];
    combat_obs_row(own.r, oth.r, g, p.role, o);
    float* orow = p.obs + (size_t)i * NP_NUM_OBS_COMBAT;
#pragma unroll
    for (int j = 0; j < NP_NUM_OBS_COMBAT; ++j) orow[j] = o[j];
    p.reward[i] = combat_reward(g, p.role, p.reward_scale);
    p.blood[i] = stepped ? own.blood - combat_damage(g, p.role) : own.blood;
    p.flags[i] = (mine & 1) ? 1 : 0;
    p.flags[p.ld + i] = (mine & 2) ? 1 : 0;
    p.flags[2 * (size_t)p.ld + i] = (mine & 4) ? 1 : 0;
    p.pair_reset[i] = (mine | theirs) ? 1 : 0;
    if (stepped && (crash | m1)) atomicAdd(&p.counters[5], 1ull);
    if (stepped && sd_done) atomicAdd(&p.counters[6], 1ull);
  }
}

// ------------------------------------------------------------------------------------------------
// K5b: relative-geometry records for the ROLE-sharded combat layout (egos and opponents on different ranks).
//   combat_records_kernel : per aircraft, the 8-float record every other rank needs: position (3), inertial velocity
//                           = xdot[0:3] (3), body-axis vx, blood  -> [n][8], written straight into the all-gather send slab
//   combat_relgeo_kernel  : for m (ego, enemy) index pairs into a gathered record array: AO, TA, R (3-D and 2-D), side
//                           flag, delta body-vx, delta altitude -- the pairwise terms of singlecombat_env.py:96-121,
//                           142-177 (get_AO_TA_R / get2d_AO_TA_R, envs/utils/utils.py:156-206).
// Both are HBM-bound (32 B written / 64 B read per aircraft).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) combat_records_kernel(const float* __restrict__ S, const float* __restrict__ blood,
                                                             float* __restrict__ rec, int n, int ld) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = S[(size_t)j * ld + i];
    const Trig t = make_trig(s);
    const float vt = s[6];
    const float vtc = vt <= 0.01f ? 0.01f : vt;
    const BodyVel b = body_vel(vtc, t);
    float4 lo, hi;
    lo.x = s[0]; lo.y = s[1]; lo.z = s[2];
    lo.w = b.U * (t.ct * t.cpsi) + b.V * (t.sphi * t.cpsi * t.st - t.cphi * t.spsi) + b.W * (t.cphi * t.st * t.cpsi + t.sphi * t.spsi);
    hi.x = b.U * (t.ct * t.spsi) + b.V * (t.sphi * t.spsi * t.st + t.cphi * t.cpsi) + b.W * (t.cphi * t.st * t.spsi - t.sphi * t.cpsi);
    hi.y = b.U * t.st - b.V * (t.sphi * t.ct) - b.W * (t.cphi * t.ct);
    hi.z = vt * t.cb * t.ca;
    hi.w = blood[i];
    reinterpret_cast<float4*>(rec)[2 * (size_t)i] = lo;
    reinterpret_cast<float4*>(rec)[2 * (size_t)i + 1] = hi;
  }
}

// pairwise terms of one (ego, enemy) record pair -> out[i]
__device__ __forceinline__ void relgeo_pair(const float4 a0, const float4 a1, const float4 b0, const float4 b1, float* __restrict__ out, int i) {
  const float dp[3] = {b0.x - a0.x, b0.y - a0.y, b0.z - a0.z};
  const float ve[3] = {a0.w, a1.x, a1.y}, vm[3] = {b0.w, b1.x, b1.y};
  float AO, TA, R, AO2, TA2, R2;
  ao_ta_r<3>(dp, ve, vm, AO, TA, R);
  ao_ta_r<2>(dp, ve, vm, AO2, TA2, R2);
  const float cz = ve[0] * dp[1] - ve[1] * dp[0];
  float4 o0, o1;
  o0.x = AO; o0.y = TA; o0.z = R; o0.w = AO2;
  o1.x = TA2; o1.y = R2; o1.z = (cz > 0.0f ? 1.0f : 0.0f) - (cz < 0.0f ? 1.0f : 0.0f);
  o1.w = b1.z - a1.z;   // enemy body-vx minus ego body-vx
  reinterpret_cast<float4*>(out)[2 * (size_t)i] = o0;
  reinterpret_cast<float4*>(out)[2 * (size_t)i + 1] = o1;
}

__global__ void __launch_bounds__(256) combat_relgeo_kernel(const float* __restrict__ rec, const int32_t* __restrict__ ego_idx,
                                                            const int32_t* __restrict__ enm_idx, float* __restrict__ out, int m) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const float4 a0 = reinterpret_cast<const float4*>(rec)[2 * (size_t)ego_idx[i]], a1 = reinterpret_cast<const float4*>(rec)[2 * (size_t)ego_idx[i] + 1];
    const float4 b0 = reinterpret_cast<const float4*>(rec)[2 * (size_t)enm_idx[i]], b1 = reinterpret_cast<const float4*>(rec)[2 * (size_t)enm_idx[i] + 1];
    relgeo_pair(a0, a1, b0, b1, out, i);
  }
}

// The same terms with the exchange FUSED into the kernel: no gathered array, no NCCL call.  slabs[r] is rank r's record slab
// ([n_local][8]) in peer-mapped memory (NVLink P2P: torch symmetric memory); global record g lives at slabs[g / n_local] +
// 8 (g % n_local).  The partner's records are pulled by the kernel's own loads, so the NVLink transfer of one pair
// overlaps the arithmetic of the others, and only the 32 B per pair this rank needs cross the link (an all-gather moves
// world x n_local x 32 B to every rank).
__global__ void __launch_bounds__(256) combat_relgeo_peers_kernel(const float* const* __restrict__ slabs, int n_local,
                                                                  const int32_t* __restrict__ ego_idx, const int32_t* __restrict__ enm_idx,
                                                                  float* __restrict__ out, int m) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const int ge = ego_idx[i], gm = enm_idx[i];
    const float4* pe = reinterpret_cast<const float4*>(slabs[ge / n_local]) + 2 * (size_t)(ge % n_local);
    const float4* pm = reinterpret_cast<const float4*>(slabs[gm / n_local]) + 2 * (size_t)(gm % n_local);
    // peer memory is written by another GPU between launches: volatile-free plain loads are fine (a new launch after the
    // cross-device barrier), but they must not come through the non-coherent read-only path
    const float4 a0 = __ldcg(pe), a1 = __ldcg(pe + 1), b0 = __ldcg(pm), b1 = __ldcg(pm + 1);
    relgeo_pair(a0, a1, b0, b1, out, i);
  }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int np_version(void) { return NP_ABI_VERSION; }

size_t np_last_error(char* buf, size_t cap) {
  if (buf && cap) {
    const size_t k = g_err.size() < cap - 1 ? g_err.size() : cap - 1;
    memcpy(buf, g_err.data(), k);
    buf[k] = 0;
  }
  return g_err.size();
}

int np_aero_pack_host(const float* blob, size_t n_floats, const np_net_desc* descs, const double* norm, int n_nets,
                      uint32_t* out_words, size_t cap_words, size_t* n_words) {
  if (!blob || !descs || !norm || !n_words) return fail(NP_EINVAL, "np_aero_pack_host: null argument");
  std::vector<uint32_t> image;
  const std::string err = pack_aero_image(blob, n_floats, descs, norm, n_nets, &image);
  if (!err.empty()) return fail(NP_EINVAL, "np_aero_pack_host: " + err);
  *n_words = image.size();
  if (out_words) {
    if (cap_words < image.size()) return fail(NP_EINVAL, "np_aero_pack_host: output buffer too small");
    memcpy(out_words, image.data(), image.size() * 4);
  }
  return NP_OK;
}

int np_aero_create(const float* blob, size_t n_floats, const np_net_desc* descs, const double* norm, int n_nets,
                   np_aero** out) {
  if (!blob || !descs || !norm || !out) return fail(NP_EINVAL, "np_aero_create: null argument");
  std::vector<uint32_t> image;
  const std::string err = pack_aero_image(blob, n_floats, descs, norm, n_nets, &image);
  if (!err.empty()) return fail(NP_EINVAL, "np_aero_create: " + err);
  std::unique_ptr<np_aero, int (*)(np_aero*)> a(new np_aero(), np_aero_destroy);   // freed on every error return below
  a->bytes = (int)image.size() * 4;
  NP_CUDA(cudaGetDevice(&a->device));
  NP_CUDA(cudaMalloc(&a->image_dev, a->bytes));
  NP_CUDA(cudaMemcpy(a->image_dev, image.data(), a->bytes, cudaMemcpyHostToDevice));
  const int c0_smem = a->bytes + kNumSlots * 32 * 8 + 16;
  NP_CUDA(cudaFuncSetAttribute(f16_c0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c0_smem));
  f16_c0_kernel<<<1, 32, c0_smem>>>(a->image_dev, a->bytes);
  NP_CUDA(cudaGetLastError());
  NP_CUDA(cudaDeviceSynchronize());
  *out = a.release();
  return NP_OK;
}

int np_aero_destroy(np_aero* aero) {
  if (!aero) return NP_OK;
  DeviceGuard guard(aero->device);
  cudaFree(aero->image_dev);
  delete aero;
  return NP_OK;
}

size_t np_env_workspace_bytes(const np_env_cfg* cfg) {
  if (!cfg) return 0;
  return (((size_t)kWorkspaceRows * (size_t)cfg->ld * sizeof(float) + 127) / 128) * 128 + 256 /* counters */;
}

}  // extern "C"

template <int BS, int MINB, int MODE, bool TAB = false, int TASK = NP_TASK_HEADING>
static int launch_step(np_env* env, const StepParams& p, cudaStream_t st) {
  const int smem = step_smem_bytes(p.aero_bytes, BS, MODE, TAB);
  static int configured[64] = {};  // per device: the attribute lives in the device's context
  auto kern = f16_step_kernel<BS, MINB, MODE, TAB, TASK>;
  const int dev = env->device;
  if (configured[dev & 63] < smem) {
    NP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev & 63] = smem;
  }
  const int npairs = p.pair_end - p.pair_begin;
  if (npairs <= 0) return NP_OK;
  const int want = (npairs + BS - 1) / BS;
  const int grid = want < env->num_sms * MINB ? want : env->num_sms * MINB;
  env->grid = grid;
  env->smem = smem;
  env->last_block = BS;
  kern<<<grid, BS, smem, st>>>(p);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

// BaseEnv.step: one instantiation per task (control_env.py:28-35), chosen here
template <int BS, int MINB, bool TAB = false>
static int launch_env_step(np_env* env, const StepParams& p, cudaStream_t st) {
  switch (env->cfg.task) {
    case NP_TASK_HEADING: return launch_step<BS, MINB, MODE_STEP, TAB, NP_TASK_HEADING>(env, p, st);
    case NP_TASK_CONTROL: return launch_step<BS, MINB, MODE_STEP, TAB, NP_TASK_CONTROL>(env, p, st);
    default: return launch_step<BS, MINB, MODE_STEP, TAB, NP_TASK_TRACKING>(env, p, st);
  }
}

// K1c (coop_step_kernel.cuh): populations of up to one wave of its CTAs, MLP aero back-end, no forced CTA shape
static bool coop_eligible(const np_env* env, const StepParams& p) {
  return env->coop_pairs > 0 && !env->block && !env->tables && p.pair_end - p.pair_begin <= env->coop_pairs;
}
template <int MODE>
static int launch_coop(np_env* env, const StepParams& p, cudaStream_t st) {
  const int npairs = p.pair_end - p.pair_begin;
  if (npairs <= 0) return NP_OK;
  const int smem = coop_smem_bytes(p.aero_bytes);
  const int want = (npairs + kCoopPairs - 1) / kCoopPairs;
  // eight warps while one CTA per SM covers the population (9 472 aircraft), four (two CTAs per SM) up to 18 944
  const int nw = env->coop_warps ? env->coop_warps : (want <= env->num_sms ? 8 : 4);
  const int per_sm = nw == 8 ? 1 : 2;
  env->grid = want < env->num_sms * per_sm ? want : env->num_sms * per_sm;
  if (env->coop_grid > 0 && env->grid > env->coop_grid) env->grid = env->coop_grid;
  env->smem = smem;
  env->last_block = nw * 32;
  static int configured[64][10] = {};
  auto launch = [&](auto kern, int t) -> int {
    if (configured[env->device & 63][t] < smem) {
      NP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured[env->device & 63][t] = smem;
    }
    // programmatic dependent launch: the CTAs may start (and stage the aero image) while the previous kernel of the stream drains
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(env->grid);
    lc.blockDim = dim3(nw * 32);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = env->pdl ? 1 : 0;
    NP_CUDA(cudaLaunchKernelEx(&lc, kern, p));
    return NP_OK;
  };
  if constexpr (MODE == MODE_PLAN) {   // PlanningEnv flies the tracking task (planning_env.py:33)
    return nw == 8 ? launch(f16_step_coop_kernel<NP_TASK_TRACKING, 8, MODE_PLAN>, 6) : launch(f16_step_coop_kernel<NP_TASK_TRACKING, 4, MODE_PLAN>, 7);
  } else if constexpr (MODE == MODE_COMBAT) {
    return nw == 8 ? launch(f16_step_coop_kernel<NP_TASK_HEADING, 8, MODE_COMBAT>, 8) : launch(f16_step_coop_kernel<NP_TASK_HEADING, 4, MODE_COMBAT>, 9);
  } else {
    if (nw == 8) {
      switch (env->cfg.task) {
        case NP_TASK_HEADING: return launch(f16_step_coop_kernel<NP_TASK_HEADING, 8>, 3);
        case NP_TASK_CONTROL: return launch(f16_step_coop_kernel<NP_TASK_CONTROL, 8>, 4);
        default: return launch(f16_step_coop_kernel<NP_TASK_TRACKING, 8>, 5);
      }
    }
    switch (env->cfg.task) {
      case NP_TASK_HEADING: return launch(f16_step_coop_kernel<NP_TASK_HEADING, 4>, 0);
      case NP_TASK_CONTROL: return launch(f16_step_coop_kernel<NP_TASK_CONTROL, 4>, 1);
      default: return launch(f16_step_coop_kernel<NP_TASK_TRACKING, 4>, 2);
    }
  }
}

// Block size of the MLP step for a range of `npairs` aircraft pairs.  384 threads (12 warps, 168 registers, no spills) is the
// fastest shape per SM; 512 (16 warps at 128 registers, a few spills) runs at ~0.93 of its rate but covers a third more aircraft
// per wave.  A persistent grid of 148 CTAs works in whole waves of 148 x BS pairs, so when the population is a small number
// of waves (a strong-scaling shard: 125 k aircraft = 1.1 waves of 384) the wider block that saves a whole wave wins.
static int pick_block(const np_env* env, int npairs) {
  if (env->block) return env->block;
  // (Up to 18 944 aircraft the step does not come here at all: K1c, coop_step_kernel.cuh.)
  // Small populations (what the reference trains at: 3 000 envs, scripts/train_heading.sh:13) are LATENCY bound: a step is as
  // long as one warp's pass through the kernel, and that pass is ~2x shorter when the warp has its scheduler to itself.
  // 128-thread CTAs put one warp on each of an SM's four schedulers and spread the population over 3x as many SMs; up to two of
  // them share an SM (measured, tools/small_n_sweep.py: n = 3 000 ... 37 888: 29 us vs 48 us with 384-thread CTAs; 50 000 / 75 000:
  // 37 / 40 vs 49 us; from 100 000 the 384-thread shape wins again).
  if (npairs <= env->num_sms * 128 * 2) return 128;
  auto cost = [&](int bs, double rate) {
    const long slabs = (npairs + bs - 1) / bs;
    const long waves = (slabs + env->num_sms - 1) / env->num_sms;
    return (double)waves * bs / rate;
  };
  return cost(512, 0.93) < cost(384, 1.0) ? 512 : 384;
}

static StepParams make_params(np_env* env, const float* action, const float* draws, const float* noise) {
  StepParams p;
  p.cfg = env->cfg;
  p.s = env->buf.s_dev;
  p.u = env->buf.u_dev;
  p.tgt = env->buf.tgt_dev;
  p.step_count = env->buf.step_count_dev;
  p.flags = env->buf.flags_dev;
  p.obs = env->buf.obs_dev;
  p.reward = env->buf.reward_dev;
  p.cache = reinterpret_cast<float*>(env->buf.workspace_dev);
  p.pid = p.cache + (size_t)kCacheRows * env->cfg.ld;
  p.blood = p.pid + (size_t)kPidRows * env->cfg.ld;
  uint8_t* pair_reset_row = reinterpret_cast<uint8_t*>(p.blood + env->cfg.ld);
  p.n_sub = 1;
  p.pid_first = 0;
  p.pair_begin = 0;
  p.pair_end = (env->cfg.n + 1) / 2;
  p.counters = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(env->buf.workspace_dev) +
                                                     (((size_t)kWorkspaceRows * env->cfg.ld * 4 + 127) / 128) * 128);
  static_assert(kWorkspaceRows == kCacheRows + kPidRows + 2, "workspace layout");
  p.rng_epoch = reinterpret_cast<const uint32_t*>(p.counters + 16);   // second 128-B line of the 256-B block: not the line the counter atomics hit
  p.aero = env->aero ? env->aero->image_dev : nullptr;
  p.aero_bytes = env->aero ? env->aero->bytes : 0;
  p.tab = 0;
  if (env->tables) {
    p.aero = reinterpret_cast<const uint32_t*>(env->tables->image_dev);
    p.aero_bytes = kTablesFloats * 4;
    p.tab = 1;
  }
  p.action = action;
  p.draws = draws;
  p.noise = noise;
  p.flags_mirror = env->mirror;
  p.flags_mirror_ld = env->mirror_ld;
  p.obs_stg = env->obs_stg;
  p.records = nullptr;
  p.pair_reset = pair_reset_row;
  p.index_stride = env->cfg.index_stride > 0 ? env->cfg.index_stride : 1;
  p.step_index = env->step_index;
  return p;
}

extern "C" {

static int env_create_impl(const np_env_cfg* cfg, const np_aero* aero, const np_tables* tables, np_env** out);

int np_env_create(const np_env_cfg* cfg, const np_aero* aero, np_env** out) {
  if (cfg && cfg->model == NP_MODEL_F16 && !aero) return fail(NP_EINVAL, "np_env_create: the F16 plug-in needs an np_aero");
  return env_create_impl(cfg, aero, nullptr, out);
}

int np_env_create_tables(const np_env_cfg* cfg, const np_tables* tables, np_env** out) {
  if (!tables) return fail(NP_EINVAL, "np_env_create_tables: null np_tables");
  if (cfg && cfg->model != NP_MODEL_F16) return fail(NP_EINVAL, "np_env_create_tables: the tables are the F16 plug-in's aero data");
  return env_create_impl(cfg, nullptr, tables, out);
}

static int env_create_impl(const np_env_cfg* cfg, const np_aero* aero, const np_tables* tables, np_env** out) {
  if (!cfg || !out) return fail(NP_EINVAL, "np_env_create: null argument");
  if (cfg->model != NP_MODEL_F16 && cfg->model != NP_MODEL_UAV) return fail(NP_EINVAL, "np_env_create: unknown aircraft model");
  if (cfg->n <= 0 || cfg->ld < cfg->n + (cfg->n & 1) || cfg->ld % 4)
    return fail(NP_EINVAL, "np_env_create: need n > 0, ld >= n rounded up to even, ld % 4 == 0");
  if (cfg->task < NP_TASK_HEADING || cfg->task > NP_TASK_TRACKING) return fail(NP_EINVAL, "np_env_create: unknown task");
  std::unique_ptr<np_env> e(new np_env());   // freed on every error return below
  e->cfg = *cfg;
  if (e->cfg.combat_pairs_per_env <= 0) e->cfg.combat_pairs_per_env = 1;
  if (e->cfg.combat_reward_scale == 0.0f) e->cfg.combat_reward_scale = 0.01f;   // singlecombat_env.py:176-177
  if (e->cfg.combat_pairs_per_env > 2) return fail(NP_EINVAL, "np_env_create: combat_pairs_per_env must be 1 (1-v-1) or 2 (2-v-2)");
  e->aero = aero;
  e->tables = tables;
  memset(&e->buf, 0, sizeof(e->buf));
  NP_CUDA(cudaGetDevice(&e->device));
  if (aero && aero->device != e->device) return fail(NP_EINVAL, "np_env_create: the np_aero lives on another device");
  if (tables && tables->device != e->device) return fail(NP_EINVAL, "np_env_create_tables: the np_tables live on another device");
  NP_CUDA(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, e->device));
  // 384 threads x 1 CTA/SM: 12 warps = 3 per scheduler (block sizes whose warp count is not a multiple of 4 load the
  // four schedulers unevenly and measured 17 % slower), 168 registers -> no spills; 256 x 2 CTA/SM (128 regs) measured
  // 1.65e9 vs 2.05e9 aircraft-steps/s (profiles/r01_variants.txt).  pick_block() switches to 512 where that saves a wave.
  if (const char* b = getenv("NPLANE_BLOCK")) {
    e->block = atoi(b);
    if (e->block % 32 || e->block < 128 || e->block > 512) return fail(NP_EINVAL, "NPLANE_BLOCK must be a multiple of 32 in [128, 512]");
  }
  e->coop_pairs = e->num_sms * 2 * kCoopPairs;   // one wave of K1c CTAs
  if (const char* b = getenv("NPLANE_COOP_PAIRS")) e->coop_pairs = atoi(b);
  if (const char* b = getenv("NPLANE_PDL")) e->pdl = atoi(b) != 0;
  if (const char* b = getenv("NPLANE_COOP_GRID")) e->coop_grid = atoi(b);
  if (const char* b = getenv("NPLANE_COOP_WARPS")) {
    e->coop_warps = atoi(b);
    if (e->coop_warps != 0 && e->coop_warps != 4 && e->coop_warps != 8) return fail(NP_EINVAL, "NPLANE_COOP_WARPS must be 4 or 8");
  }
  if (const char* b = getenv("NPLANE_TAB_BLOCK")) e->tab_block = atoi(b);
  if (const char* b = getenv("NPLANE_OBS_STORE")) e->obs_stg = strcmp(b, "stg") == 0;
  if (const char* b = getenv("NPLANE_TAB_KERNEL")) e->tab_pairs = strcmp(b, "pairs") == 0;
  *out = e.release();
  return NP_OK;
}

int np_env_bind(np_env* env, const np_buffers* b, void* stream) {
  if (!env || !b) return fail(NP_EINVAL, "np_env_bind: null argument");
  if (!b->s_dev || !b->u_dev || !b->tgt_dev || !b->step_count_dev || !b->flags_dev || !b->obs_dev || !b->reward_dev ||
      !b->workspace_dev)
    return fail(NP_EINVAL, "np_env_bind: null buffer");
  if (((uintptr_t)b->obs_dev & 15) || ((uintptr_t)b->workspace_dev & 127) || ((uintptr_t)b->s_dev & 15) ||
      ((uintptr_t)b->u_dev & 15) || ((uintptr_t)b->tgt_dev & 15) || ((uintptr_t)b->reward_dev & 7) ||
      ((uintptr_t)b->step_count_dev & 7) || ((uintptr_t)b->flags_dev & 1))
    return fail(NP_EINVAL, "np_env_bind: buffers must be 16-byte (workspace 128-byte) aligned");
  DeviceGuard guard(env->device);
  env->buf = *b;
  env->bound = true;
  // invalidate the coefficient-cache keys: 0xFFFFFFFF is a NaN pattern no stored alpha/beta can equal bitwise.  Enqueued
  // on the caller's stream, like every later step.
  NP_CUDA(cudaMemsetAsync(reinterpret_cast<float*>(b->workspace_dev) + (size_t)kNumAB2 * env->cfg.ld, 0xFF,
                          2 * (size_t)env->cfg.ld * sizeof(float), (cudaStream_t)stream));
  // termination counters and the RNG epoch word (np_env_rng_advance) start at zero
  NP_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(b->workspace_dev) + (((size_t)kWorkspaceRows * env->cfg.ld * 4 + 127) / 128) * 128, 0, 256,
                          (cudaStream_t)stream));
  return NP_OK;
}

int np_env_rebind_outputs(np_env* env, float* obs_dev, float* reward_dev) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_rebind_outputs: env not bound");
  if (!obs_dev || !reward_dev || ((uintptr_t)obs_dev & 7) || ((uintptr_t)reward_dev & 7))
    return fail(NP_EINVAL, "np_env_rebind_outputs: obs / reward must be 8-byte aligned device pointers");
  env->buf.obs_dev = obs_dev;
  env->buf.reward_dev = reward_dev;
  return NP_OK;
}

int np_rollout_returns(const float* rewards_dev, float* value_preds_dev, const float* masks_dev, const float* bad_masks_dev,
                       float* returns_dev, int T, int M, double gamma, double gae_lambda, int use_gae, int use_proper_time_limits,
                       void* stream) {
  if (!rewards_dev || !value_preds_dev || !masks_dev || !returns_dev || T <= 0 || M <= 0 || (use_proper_time_limits && !bad_masks_dev))
    return fail(NP_EINVAL, "np_rollout_returns: bad argument");
  DeviceGuard guard(device_of(returns_dev));
  const int want = (M + 255) / 256;
  rollout_returns_kernel<<<want < 2368 ? want : 2368, 256, 0, (cudaStream_t)stream>>>(
      rewards_dev, value_preds_dev, masks_dev, bad_masks_dev, returns_dev, T, M, (float)gamma, (float)(gamma * gae_lambda), use_gae,
      use_proper_time_limits);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_rollout_masks(const uint8_t* flags_dev, int ld, int num_envs, int num_agents, float* masks_dev, float* bad_masks_dev,
                     uint8_t* reset_env_dev, void* stream) {
  if (!flags_dev || !masks_dev || !bad_masks_dev || num_envs <= 0 || num_agents <= 0 || (long long)num_envs * num_agents > ld)
    return fail(NP_EINVAL, "np_rollout_masks: bad argument");
  DeviceGuard guard(device_of(masks_dev));
  const int want = (num_envs + 255) / 256;
  rollout_masks_kernel<<<want < 2368 ? want : 2368, 256, 0, (cudaStream_t)stream>>>(flags_dev, ld, num_envs, num_agents, masks_dev,
                                                                                   bad_masks_dev, reset_env_dev);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_env_set_cfg(np_env* env, const np_env_cfg* cfg) {
  if (!env || !cfg) return fail(NP_EINVAL, "np_env_set_cfg: null argument");
  if (cfg->n != env->cfg.n || cfg->ld != env->cfg.ld || cfg->task != env->cfg.task || cfg->model != env->cfg.model)
    return fail(NP_EINVAL, "np_env_set_cfg: n, ld, task and model are fixed at creation");
  env->cfg = *cfg;
  return NP_OK;
}

int np_env_destroy(np_env* env) {
  if (!env) return NP_OK;
  DeviceGuard guard(env->device);
  if (env->hs[0]) {
    for (cudaStream_t st : env->hs) cudaStreamDestroy(st);
    cudaEventDestroy(env->hev_start);
    cudaEventDestroy(env->hev_done);
    for (int c = 0; c < np_env::kMaxHostChunks; ++c) { cudaEventDestroy(env->hev_up[c]); cudaEventDestroy(env->hev_run[c]); }
  }
  delete env;
  return NP_OK;
}

int np_env_reset(np_env* env, const float* draws_dev, const float* noise_dev, void* stream) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_reset: env not bound");
  DeviceGuard guard(env->device);
  StepParams p = make_params(env, nullptr, draws_dev, noise_dev);
  env->step_index++;
  const int want = (env->cfg.n + 255) / 256;
  const int grid = want < env->num_sms * 8 ? want : env->num_sms * 8;
  if (env->cfg.model == NP_MODEL_UAV) uav_env_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  else f16_reset_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

static int step_range_impl(np_env* env, const float* action_dev, const float* draws_dev, const float* noise_dev, int first,
                           int count, bool advance, void* stream);

int np_env_step(np_env* env, const float* action_dev, const float* draws_dev, const float* noise_dev, void* stream) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_step: env not bound");
  DeviceGuard guard(env->device);
  return step_range_impl(env, action_dev, draws_dev, noise_dev, 0, env->cfg.n, true, stream);
}

int np_env_step_range(np_env* env, const float* action_dev, const float* draws_dev, const float* noise_dev, int first_aircraft,
                      int count, int advance_step_index, void* stream) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_step_range: env not bound");
  if (first_aircraft < 0 || count < 0 || first_aircraft + count > env->cfg.n || (first_aircraft & 1) ||
      ((count & 1) && first_aircraft + count != env->cfg.n))
    return fail(NP_EINVAL, "np_env_step_range: the range must start on an even aircraft and cover whole pairs (except the tail)");
  DeviceGuard guard(env->device);
  return step_range_impl(env, action_dev, draws_dev, noise_dev, first_aircraft, count, advance_step_index != 0, stream);
}

static int step_range_impl(np_env* env, const float* action_dev, const float* draws_dev, const float* noise_dev, int first,
                           int count, bool advance, void* stream) {
  if (!action_dev || ((uintptr_t)action_dev & 15)) return fail(NP_EINVAL, "np_env_step: action must be a 16-byte aligned device pointer");
  if (advance) env->step_index++;           // every range of one logical step sees the same RNG counter
  StepParams p = make_params(env, action_dev, draws_dev, noise_dev);
  p.step_index = env->step_index - 1;
  p.pair_begin = first / 2;
  p.pair_end = (first + count + 1) / 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (env->cfg.model == NP_MODEL_UAV) {
    if (count == 0) return NP_OK;
    const int want = (count + 255) / 256;
    // TMA bulk copies need 16-byte aligned global addresses: row bases, and the first aircraft a multiple of 16 (u8 flag rows)
    const uintptr_t bases = (uintptr_t)p.s | (uintptr_t)p.u | (uintptr_t)p.tgt | (uintptr_t)p.step_count | (uintptr_t)p.flags |
                            (uintptr_t)p.obs | (uintptr_t)p.reward | (uintptr_t)p.action;
    const bool slab_ok = !(bases & 15) && !(first & 15) && !(env->cfg.ld & 15) && count >= uavslab::kSlab && !getenv("NPLANE_UAV_SCALAR");
    if (slab_ok) {
      static bool attr_set[64] = {};  // per device: the attribute lives in the device's context
      const int dev = env->device;
      if (!attr_set[dev & 63]) {
        NP_CUDA(cudaFuncSetAttribute(uav_step_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, uavslab::SMEM_BYTES));
        attr_set[dev & 63] = true;
      }
      env->grid = want < env->num_sms * 4 ? want : env->num_sms * 4;
      env->smem = uavslab::SMEM_BYTES;
      uav_step_slab_kernel<<<env->grid, uavslab::kSlab, uavslab::SMEM_BYTES, st>>>(p);
    } else {
      env->grid = want < env->num_sms * 8 ? want : env->num_sms * 8;
      env->smem = 0;
      uav_env_kernel<true><<<env->grid, 256, 0, st>>>(p);
    }
    NP_CUDA(cudaGetLastError());
    return NP_OK;
  }
  if (env->tables && !env->tab_pairs) {   // K1t: one aircraft per thread
    if (count == 0) return NP_OK;
    const int smem = table_step_smem_bytes();
    const int want = (count + kTabBS - 1) / kTabBS;
    env->grid = want < env->num_sms * kTabMinB ? want : env->num_sms * kTabMinB;
    env->smem = smem;
    env->last_block = kTabBS;
    static bool attr_set[64][3] = {};
    auto launch = [&](auto kern, int t) -> int {
      if (!attr_set[env->device & 63][t]) {
        NP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[env->device & 63][t] = true;
      }
      kern<<<env->grid, kTabBS, smem, st>>>(p);
      NP_CUDA(cudaGetLastError());
      return NP_OK;
    };
    switch (env->cfg.task) {
      case NP_TASK_HEADING: return launch(f16_table_step_kernel<NP_TASK_HEADING>, 0);
      case NP_TASK_CONTROL: return launch(f16_table_step_kernel<NP_TASK_CONTROL>, 1);
      default: return launch(f16_table_step_kernel<NP_TASK_TRACKING>, 2);
    }
  }
  if (env->tables) {                      // NPLANE_TAB_KERNEL=pairs: K1's two-aircraft-per-thread shape with table coefficients
    switch (env->tab_block) {
#if defined(NPLANE_ALL_BLOCKS) || defined(NPLANE_TAB_BLOCKS)
      case 128: return launch_env_step<128, 4, true>(env, p, st);
      case 256: return launch_env_step<256, 2, true>(env, p, st);
      case 512: return launch_env_step<512, 1, true>(env, p, st);
#endif
      case 384: return launch_env_step<384, 1, true>(env, p, st);  // 4.36e9 vs 4.16e9 (256 x 2) / 4.20e9 (512) / 4.10e9 (128 x 4) at n = 10^6
      default: return fail(NP_EINVAL, "np_env_step: table back-end block size not compiled in (build with -DNPLANE_ALL_BLOCKS)");
    }
  }
  if (coop_eligible(env, p)) return launch_coop<MODE_STEP>(env, p, st);   // K1c: the warps of a CTA share each pair's MLPs
  switch (pick_block(env, p.pair_end - p.pair_begin)) {
    case 128: return launch_env_step<128, 2>(env, p, st);   // two CTAs may share an SM (2 x 110 KB of shared memory)
#ifdef NPLANE_ALL_BLOCKS
    case 256: return launch_env_step<256, 2>(env, p, st);
    case 320: return launch_env_step<320, 1>(env, p, st);
    case 352: return launch_env_step<352, 1>(env, p, st);
    case 448: return launch_env_step<448, 1>(env, p, st);
#endif
    case 384: return launch_env_step<384, 1>(env, p, st);
    case 512: return launch_env_step<512, 1>(env, p, st);
    default: return fail(NP_EINVAL, "np_env_step: block size not compiled in (build with -DNPLANE_ALL_BLOCKS)");
  }
}

// The numpy boundary of GPUVecEnv.step (env_wrappers.py:93-103) in one native call: host actions in, host observations /
// rewards / flags out.  The population is cut into chunks (edges[0..n_chunks], whole pairs) and pipelined over three
// in-order streams, one per engine: chunk c's actions are staged (memcpy into pinned memory) and uploaded while chunk
// c-1 runs and chunk c-2's 88 B/aircraft observations -- the resource this boundary is bound by -- travel to the host.
// Issuing a chunk costs a few microseconds here against ~110 us of interpreter time in the Python pipeline it replaces,
// which is what delayed the start of the download (profiles/r01_variants.txt).  Returns when the host buffers are ready.
int np_env_step_host(np_env* env, const float* action_host, float* action_pinned, float* action_dev, float* obs_pinned,
                     float* reward_pinned, uint8_t* flags_pinned, const int* edges, int n_chunks, void* stream) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_step_host: env not bound");
  if (!action_host || !action_pinned || !action_dev || !obs_pinned || !reward_pinned || !flags_pinned || !edges || n_chunks < 1 ||
      n_chunks > np_env::kMaxHostChunks || edges[0] != 0 || edges[n_chunks] != env->cfg.n)
    return fail(NP_EINVAL, "np_env_step_host: bad argument (1..16 chunks, edges[0] = 0, edges[n_chunks] = n)");
  DeviceGuard guard(env->device);
  for (int c = 0; c < n_chunks; ++c)
    if (edges[c + 1] <= edges[c] || (edges[c] & 1)) return fail(NP_EINVAL, "np_env_step_host: chunks must be non-empty and start on whole pairs");
  if (!env->hs[0]) {
    for (cudaStream_t& st : env->hs) NP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    NP_CUDA(cudaEventCreateWithFlags(&env->hev_start, cudaEventDisableTiming));
    NP_CUDA(cudaEventCreateWithFlags(&env->hev_done, cudaEventDisableTiming));
    for (int c = 0; c < np_env::kMaxHostChunks; ++c) {
      NP_CUDA(cudaEventCreateWithFlags(&env->hev_up[c], cudaEventDisableTiming));
      NP_CUDA(cudaEventCreateWithFlags(&env->hev_run[c], cudaEventDisableTiming));
    }
  }
  cudaStream_t main_st = (cudaStream_t)stream, up = env->hs[0], run = env->hs[1], down = env->hs[2];
  const int n = env->cfg.n, ld = env->cfg.ld, A = 4, D = NP_NUM_OBS;
  // NPLANE_HOST_PIPE_V2=1 (experiment, F16 plug-in): only the 88 B/aircraft observation block goes through the copy engine; the
  // range kernels read their actions straight from the pinned staging buffer (zero-copy: no H2D copy, no upload stream) and
  // write reward + flags (7 B/aircraft) straight into the pinned outputs: 2 operations per chunk instead of 8.  MEASURED
  // SLOWER on the same box (2.34-2.37 vs 2.20-2.26 ms per 10^6-aircraft step, profiles/r02_e2e_boundaries.txt): kernels that
  // touch host memory run longer, which delays every download behind them.  The explicit upload / download copies stay.
  const bool direct = env->cfg.model == NP_MODEL_F16 && !(n & 1) && getenv("NPLANE_HOST_PIPE_V2");
  float* act_src = action_dev;
  struct Restore {   // the env's own reward buffer / no mirror again on every way out
    np_env* e;
    float* rew;
    ~Restore() { e->buf.reward_dev = rew; e->mirror = nullptr; }
  } restore_on_exit{env, env->buf.reward_dev};
  if (direct) {
    float* rew_d = nullptr;
    uint8_t* flg_d = nullptr;
    NP_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&act_src), action_pinned, 0));
    NP_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&rew_d), reward_pinned, 0));
    NP_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&flg_d), flags_pinned, 0));
    env->buf.reward_dev = rew_d;
    env->mirror = flg_d;
    env->mirror_ld = n;
  }
  NP_CUDA(cudaEventRecord(env->hev_start, main_st));   // everything already queued on the caller's stream comes first
  for (cudaStream_t st : env->hs) NP_CUDA(cudaStreamWaitEvent(st, env->hev_start, 0));
  for (int c = 0; c < n_chunks; ++c) {
    const int i0 = edges[c], cnt = edges[c + 1] - edges[c];
    const size_t a_off = (size_t)i0 * A, a_bytes = (size_t)cnt * A * sizeof(float);
    if (action_host != action_pinned) memcpy(action_pinned + a_off, action_host + a_off, a_bytes);   // overlaps the GPU
    if (!direct) {
      NP_CUDA(cudaMemcpyAsync(action_dev + a_off, action_pinned + a_off, a_bytes, cudaMemcpyHostToDevice, up));
      NP_CUDA(cudaEventRecord(env->hev_up[c], up));
      NP_CUDA(cudaStreamWaitEvent(run, env->hev_up[c], 0));
    }
    const int rc = step_range_impl(env, act_src, nullptr, nullptr, i0, cnt, c == 0, run);
    if (rc != NP_OK) return rc;
    NP_CUDA(cudaEventRecord(env->hev_run[c], run));
    NP_CUDA(cudaStreamWaitEvent(down, env->hev_run[c], 0));
    NP_CUDA(cudaMemcpyAsync(obs_pinned + (size_t)i0 * D, env->buf.obs_dev + (size_t)i0 * D, (size_t)cnt * D * sizeof(float),
                            cudaMemcpyDeviceToHost, down));
    if (!direct) {
      NP_CUDA(cudaMemcpyAsync(reward_pinned + i0, env->buf.reward_dev + i0, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToHost, down));
      NP_CUDA(cudaMemcpy2DAsync(flags_pinned + i0, (size_t)n, env->buf.flags_dev + i0, (size_t)ld, (size_t)cnt, 3,
                                cudaMemcpyDeviceToHost, down));   // the three flag rows of the chunk in one strided copy
    }
  }
  if (direct) {   // the last kernels' direct writes must have landed too (the download stream only covers the observations)
    NP_CUDA(cudaEventRecord(env->hev_up[0], run));
    NP_CUDA(cudaStreamWaitEvent(down, env->hev_up[0], 0));
  }
  NP_CUDA(cudaEventRecord(env->hev_done, down));
  NP_CUDA(cudaStreamWaitEvent(main_st, env->hev_done, 0));   // later work on the caller's stream sees the stepped state
  NP_CUDA(cudaEventSynchronize(env->hev_done));
  return NP_OK;
}

int np_env_plan_step(np_env* env, const float* action3_dev, int n_sub, const float* draws_dev, const float* noise_dev,
                     void* stream) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_plan_step: env not bound");
  if (env->cfg.model != NP_MODEL_F16) return fail(NP_EINVAL, "np_env_plan_step: the fused PID controller flies the F16 plug-in");
  if (env->cfg.task != NP_TASK_TRACKING) return fail(NP_EINVAL, "np_env_plan_step: PlanningEnv flies the tracking task (planning_env.py:33)");
  if (!action3_dev || ((uintptr_t)action3_dev & 3) || n_sub < 1) return fail(NP_EINVAL, "np_env_plan_step: bad action pointer or n_sub");
  DeviceGuard guard(env->device);
  StepParams p = make_params(env, action3_dev, draws_dev, noise_dev);
  p.n_sub = n_sub;
  p.pid_first = env->pid_started ? 0 : 1;
  env->pid_started = true;
  env->step_index++;
  if (env->tables) return launch_step<384, 1, MODE_PLAN, true, NP_TASK_TRACKING>(env, p, (cudaStream_t)stream);
  // train_tracking.sh flies 10 000 planning envs: K1c up to 18 944 aircraft, 128-thread CTAs up to 75 776 (latency bound, as the step)
  if (coop_eligible(env, p)) return launch_coop<MODE_PLAN>(env, p, (cudaStream_t)stream);
  // a strong-scaling shard (125 k aircraft = 1.1 waves of 148 x 384 pairs) runs in ONE wave of 512-thread CTAs (pick_block)
  switch (pick_block(env, p.pair_end - p.pair_begin)) {
    case 128: return launch_step<128, 2, MODE_PLAN, false, NP_TASK_TRACKING>(env, p, (cudaStream_t)stream);
    case 512: return launch_step<512, 1, MODE_PLAN, false, NP_TASK_TRACKING>(env, p, (cudaStream_t)stream);
    case 384: return launch_step<384, 1, MODE_PLAN, false, NP_TASK_TRACKING>(env, p, (cudaStream_t)stream);
    default: return fail(NP_EINVAL, "np_env_plan_step: block size not compiled in for the planning step (128 / 384 / 512)");
  }
}

// K5 by population, as the step: 128-thread CTAs up to 75 776 aircraft (latency bound), 384, 512 where that saves a wave
static int launch_combat(np_env* env, const StepParams& p, cudaStream_t st) {
  switch (pick_block(env, p.pair_end - p.pair_begin)) {
    case 128: return launch_step<128, 2, MODE_COMBAT>(env, p, st);
    case 512: return launch_step<512, 1, MODE_COMBAT>(env, p, st);
    case 384: return launch_step<384, 1, MODE_COMBAT>(env, p, st);
    default: return fail(NP_EINVAL, "np_env_combat_step: block size not compiled in for the combat step (128 / 384 / 512)");
  }
}

int np_env_combat_step(np_env* env, const float* action_dev, int n_sub, const float* draws_dev, void* stream) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_combat_step: env not bound");
  if (env->cfg.model != NP_MODEL_F16 || (env->cfg.n & 1)) return fail(NP_EINVAL, "np_env_combat_step: needs the F16 plug-in and an even population (pairs)");
  if (env->cfg.combat_pairs_per_env == 2 && (env->cfg.n & 3)) return fail(NP_EINVAL, "np_env_combat_step: a 2-v-2 population is a multiple of 4 aircraft");
  if (n_sub < 0 || (n_sub > 0 && (!action_dev || ((uintptr_t)action_dev & 15)))) return fail(NP_EINVAL, "np_env_combat_step: bad action pointer or n_sub");
  DeviceGuard guard(env->device);
  StepParams p = make_params(env, action_dev ? action_dev : reinterpret_cast<const float*>(env->buf.s_dev), draws_dev, nullptr);
  p.n_sub = n_sub;
  p.pid_first = env->pid_started ? 0 : 1;
  if (n_sub > 0) env->pid_started = true;
  env->step_index++;
  if (env->tables) return launch_step<384, 1, MODE_COMBAT, true>(env, p, (cudaStream_t)stream);
  if (coop_eligible(env, p)) return launch_coop<MODE_COMBAT>(env, p, (cudaStream_t)stream);   // pair-sharded duels at small populations
  return launch_combat(env, p, (cudaStream_t)stream);
}

// Role-sharded combat step, local half: np_env_combat_step for a rank that holds ONE aircraft of every env (all egos or all
// opponents): env-level reset from the pair_reset flags, n_sub sub-steps, per-aircraft terminations, and the record the
// partner rank needs (records_dev [n][28], see kCombatRecFloats) instead of obs / reward.
int np_env_combat_role_local(np_env* env, const float* action_dev, int n_sub, const float* draws_dev, float* records_dev, void* stream) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_combat_role_local: env not bound");
  if (env->cfg.model != NP_MODEL_F16 || (env->cfg.n & 1) || env->tables) return fail(NP_EINVAL, "np_env_combat_role_local: needs the F16 MLP plug-in and an even local population");
  if (n_sub < 0 || n_sub > kCombatMaxSub || (n_sub > 0 && (!action_dev || ((uintptr_t)action_dev & 15))) || !records_dev || ((uintptr_t)records_dev & 15))
    return fail(NP_EINVAL, "np_env_combat_role_local: bad action / records pointer or n_sub (0..5)");
  DeviceGuard guard(env->device);
  StepParams p = make_params(env, action_dev ? action_dev : reinterpret_cast<const float*>(env->buf.s_dev), draws_dev, nullptr);
  p.n_sub = n_sub;
  p.pid_first = env->pid_started ? 0 : 1;
  if (n_sub > 0) env->pid_started = true;
  p.records = records_dev;
  env->step_index++;
  return launch_combat(env, p, (cudaStream_t)stream);
}

// Role-sharded combat step, pair half (combat_pair_kernel): to be enqueued after a cross-device barrier that makes the
// partner rank's records visible.  partner_records_dev: the partner rank's [n][28] slab -- peer-mapped memory
// (partner_is_peer = 1: pulled over NVLink by the kernel's own loads) or a slice of an all-gathered array (0).
// role: 0 = this rank holds the egos, 1 = the opponents.  Writes obs [n][15], reward, blood, the final flags and the
// env-level reset flags of the next step.
int np_env_combat_role_pair(np_env* env, const float* own_records_dev, const float* partner_records_dev, int partner_is_peer, int role,
                            int n_sub, void* stream) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_combat_role_pair: env not bound");
  if (!own_records_dev || !partner_records_dev || (((uintptr_t)own_records_dev | (uintptr_t)partner_records_dev) & 15) || role < 0 || role > 1 ||
      n_sub < 0 || n_sub > kCombatMaxSub)
    return fail(NP_EINVAL, "np_env_combat_role_pair: bad argument");
  DeviceGuard guard(env->device);
  StepParams sp = make_params(env, nullptr, nullptr, nullptr);
  PairParams p;
  p.own = own_records_dev; p.partner = partner_records_dev;
  p.obs = sp.obs; p.reward = sp.reward; p.blood = sp.blood; p.flags = sp.flags; p.pair_reset = sp.pair_reset; p.counters = sp.counters;
  p.n = env->cfg.n; p.ld = env->cfg.ld; p.role = role; p.n_sub = n_sub; p.distance_limit = env->cfg.distance_limit; p.reward_scale = env->cfg.combat_reward_scale;
  const int want = (p.n + 255) / 256;
  const int grid = want < env->num_sms * 8 ? want : env->num_sms * 8;
  if (partner_is_peer) combat_pair_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  else combat_pair_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

size_t np_env_pair_reset_offset_bytes(const np_env_cfg* cfg) {
  return cfg ? (size_t)(kCacheRows + kPidRows + 1) * (size_t)cfg->ld * sizeof(float) : 0;
}

int np_env_combat_records(np_env* env, float* records_dev, void* stream) {
  if (!env || !env->bound || !records_dev || ((uintptr_t)records_dev & 15)) return fail(NP_EINVAL, "np_env_combat_records: bad argument");
  DeviceGuard guard(env->device);
  StepParams p = make_params(env, nullptr, nullptr, nullptr);
  const int want = (env->cfg.n + 255) / 256;
  combat_records_kernel<<<want < env->num_sms * 8 ? want : env->num_sms * 8, 256, 0, (cudaStream_t)stream>>>(p.s, p.blood, records_dev,
                                                                                                          env->cfg.n, env->cfg.ld);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_combat_relgeo(const float* records_dev, const int32_t* ego_idx_dev, const int32_t* enm_idx_dev, float* out_dev, int m,
                     void* stream) {
  if (!records_dev || !ego_idx_dev || !enm_idx_dev || !out_dev || m <= 0 || (((uintptr_t)records_dev | (uintptr_t)out_dev) & 15))
    return fail(NP_EINVAL, "np_combat_relgeo: bad argument");
  DeviceGuard guard(device_of(out_dev));
  const int want = (m + 255) / 256;
  combat_relgeo_kernel<<<want < 1184 ? want : 1184, 256, 0, (cudaStream_t)stream>>>(records_dev, ego_idx_dev, enm_idx_dev, out_dev, m);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_combat_relgeo_peers(const float* const* slabs_dev, int world, int n_local, const int32_t* ego_idx_dev, const int32_t* enm_idx_dev,
                           float* out_dev, int m, void* stream) {
  if (!slabs_dev || world < 1 || n_local < 1 || !ego_idx_dev || !enm_idx_dev || !out_dev || m <= 0 || ((uintptr_t)out_dev & 15))
    return fail(NP_EINVAL, "np_combat_relgeo_peers: bad argument");
  DeviceGuard guard(device_of(out_dev));
  const int want = (m + 255) / 256;
  combat_relgeo_peers_kernel<<<want < 1184 ? want : 1184, 256, 0, (cudaStream_t)stream>>>(slabs_dev, n_local, ego_idx_dev, enm_idx_dev,
                                                                                         out_dev, m);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

size_t np_env_blood_offset_bytes(const np_env_cfg* cfg) {
  return cfg ? (size_t)(kCacheRows + kPidRows) * (size_t)cfg->ld * sizeof(float) : 0;
}

size_t np_env_pid_offset_bytes(const np_env_cfg* cfg) {
  return cfg ? (size_t)kCacheRows * (size_t)cfg->ld * sizeof(float) : 0;
}

int np_env_set_pid_started(np_env* env, int started) {
  if (!env) return fail(NP_EINVAL, "np_env_set_pid_started: null env");
  env->pid_started = started != 0;
  return NP_OK;
}

int np_env_counters(np_env* env, uint64_t* out, void* stream) {
  if (!env || !env->bound || !out) return fail(NP_ESTATE, "np_env_counters: env not bound");
  DeviceGuard guard(env->device);
  StepParams p = make_params(env, nullptr, nullptr, nullptr);
  NP_CUDA(cudaMemcpyAsync(out, p.counters, NP_NUM_COUNTERS * sizeof(uint64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  NP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return NP_OK;
}

// CUDA-graph capture freezes every host-side launch argument, including the RNG counter the step calls advance: each replay
// would draw the same reset / noise streams.  This adds `delta` to a DEVICE word that every kernel adds to its counter; captured
// at the end of a graph of K steps with delta = K, every replay continues the sequence an eager loop would have produced.
__global__ void rng_epoch_add_kernel(uint32_t* epoch, uint32_t delta) { *epoch += delta; }

int np_env_rng_advance(np_env* env, uint32_t delta, void* stream) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_rng_advance: env not bound");
  DeviceGuard guard(env->device);
  StepParams p = make_params(env, nullptr, nullptr, nullptr);
  rng_epoch_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(const_cast<uint32_t*>(p.rng_epoch), delta);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_env_launch_info(const np_env* env, int* grid, int* block, int* smem_bytes, int* num_sms) {
  if (!env) return fail(NP_EINVAL, "np_env_launch_info: null env");
  if (grid) *grid = env->grid;
  if (block) *block = env->cfg.model == NP_MODEL_UAV ? 256 : env->last_block;
  if (smem_bytes) *smem_bytes = env->smem;
  if (num_sms) *num_sms = env->num_sms;
  return NP_OK;
}

int np_f16_nlplant(const np_aero* aero, const float* s_dev, const float* u_dev, float* xdot_dev, int n, int ld, void* stream) {
  if (!aero || !s_dev || !u_dev || !xdot_dev || n <= 0 || ld < n + (n & 1) || (ld & 1))
    return fail(NP_EINVAL, "np_f16_nlplant: bad argument (need even ld >= n rounded up to even)");
  if (((uintptr_t)s_dev | (uintptr_t)u_dev | (uintptr_t)xdot_dev) & 7) return fail(NP_EINVAL, "np_f16_nlplant: 8-byte alignment");
  DeviceGuard guard(aero->device);
  const int smem = aux_smem_bytes(aero->bytes);
  static int configured[64] = {};
  const int dev = aero->device;
  if (configured[dev & 63] < smem) {
    NP_CUDA(cudaFuncSetAttribute(f16_nlplant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev & 63] = smem;
  }
  const int want = ((n + 1) / 2 + kAuxBS - 1) / kAuxBS;
  f16_nlplant_kernel<<<want < 296 ? want : 296, kAuxBS, smem, (cudaStream_t)stream>>>(aero->image_dev, aero->bytes, s_dev, u_dev,
                                                                                   xdot_dev, n, ld);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_tables_create(const float* breakpoints, const int32_t* bp_sizes, const float* values, const int32_t* offsets, int n_tables,
                     size_t n_values, np_tables** out) {
  if (!breakpoints || !bp_sizes || !values || !offsets || !out) return fail(NP_EINVAL, "np_tables_create: null argument");
  const int want_sizes[5] = {kNA1, kNA2, kNB1, kND1, kND2};
  for (int j = 0; j < 5; ++j)
    if (bp_sizes[j] != want_sizes[j]) return fail(NP_EINVAL, "np_tables_create: breakpoint grid sizes must be 20, 14, 19, 5, 3");
  if (n_tables != kNumTables || n_values != (size_t)kTableValues) return fail(NP_EINVAL, "np_tables_create: expected 43 tables / 13405 values");
  for (int t = 0; t < kNumTables; ++t)
    if (offsets[t] != table_offset(t) - kBpFloats) return fail(NP_EINVAL, "np_tables_create: table offsets do not match the fixed table order");
  for (int j = 0, o = 0; j < 5; o += want_sizes[j], ++j)
    for (int i = 1; i < want_sizes[j]; ++i)
      if (!(breakpoints[o + i] > breakpoints[o + i - 1])) return fail(NP_EINVAL, "np_tables_create: breakpoints must increase");
  for (int i = 0; i < kNA2; ++i)   // the step kernels derive the ALPHA2 cell from the ALPHA1 cell (tables_device.cuh)
    if (breakpoints[kBpA2 + i] != breakpoints[kBpA1 + i]) return fail(NP_EINVAL, "np_tables_create: ALPHA2 must be the first 14 points of ALPHA1");
  std::vector<float> host(kTablesFloats, 0.0f);
  memcpy(host.data(), breakpoints, 61 * sizeof(float));
  memcpy(host.data() + kBpFloats, values, (size_t)kTableValues * sizeof(float));
  std::unique_ptr<np_tables, int (*)(np_tables*)> t(new np_tables(), np_tables_destroy);   // freed on every error return below
  NP_CUDA(cudaGetDevice(&t->device));
  NP_CUDA(cudaMalloc(&t->image_dev, kTablesFloats * sizeof(float)));
  NP_CUDA(cudaMemcpy(t->image_dev, host.data(), kTablesFloats * sizeof(float), cudaMemcpyHostToDevice));
  *out = t.release();
  return NP_OK;
}

int np_tables_destroy(np_tables* t) {
  if (!t) return NP_OK;
  DeviceGuard guard(t->device);
  cudaFree(t->image_dev);
  delete t;
  return NP_OK;
}

int np_f16_table_coeffs(const np_tables* tables, const float* alpha_deg_dev, const float* beta_deg_dev, const float* el_deg_dev,
                        float* out_dev, int n, int ld, void* stream) {
  if (!tables || !alpha_deg_dev || !beta_deg_dev || !el_deg_dev || !out_dev || n <= 0 || ld < n)
    return fail(NP_EINVAL, "np_f16_table_coeffs: bad argument");
  DeviceGuard guard(tables->device);
  const int smem = kTablesFloats * 4 + 16;
  static int configured[64] = {};
  const int dev = tables->device;
  if (!configured[dev & 63]) {
    NP_CUDA(cudaFuncSetAttribute(f16_table_coeffs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev & 63] = 1;
  }
  const int want = (n + 255) / 256;
  f16_table_coeffs_kernel<<<want < 444 ? want : 444, 256, smem, (cudaStream_t)stream>>>(tables->image_dev, alpha_deg_dev, beta_deg_dev,
                                                                                    el_deg_dev, out_dev, n, ld);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_f16_table_nlplant(const np_tables* tables, const float* s_dev, const float* u_dev, float* xdot_dev, int n, int ld,
                         void* stream) {
  if (!tables || !s_dev || !u_dev || !xdot_dev || n <= 0 || ld < n) return fail(NP_EINVAL, "np_f16_table_nlplant: bad argument");
  DeviceGuard guard(tables->device);
  const int smem = kTablesFloats * 4 + 16;
  static int configured[64] = {};
  const int dev = tables->device;
  if (!configured[dev & 63]) {
    NP_CUDA(cudaFuncSetAttribute(f16_table_nlplant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev & 63] = 1;
  }
  const int want = (n + 255) / 256;
  f16_table_nlplant_kernel<<<want < 444 ? want : 444, 256, smem, (cudaStream_t)stream>>>(tables->image_dev, s_dev, u_dev, xdot_dev, n, ld);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_uav_nlplant(const float* s_dev, const float* u_dev, float* xdot_dev, int n, int ld, void* stream) {
  if (!s_dev || !u_dev || !xdot_dev || n <= 0 || ld < n) return fail(NP_EINVAL, "np_uav_nlplant: bad argument");
  DeviceGuard guard(device_of(xdot_dev));
  const int want = (n + 255) / 256;
  uav_nlplant_kernel<<<want < 1184 ? want : 1184, 256, 0, (cudaStream_t)stream>>>(s_dev, u_dev, xdot_dev, n, ld);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_f16_coeffs(const np_aero* aero, const float* alpha_deg_dev, const float* beta_deg_dev, const float* el_deg_dev,
                  float* out_dev, int n, int ld, void* stream) {
  if (!aero || !alpha_deg_dev || !beta_deg_dev || !el_deg_dev || !out_dev || n <= 0 || ld < n)
    return fail(NP_EINVAL, "np_f16_coeffs: bad argument");
  DeviceGuard guard(aero->device);
  const int smem = aux_smem_bytes(aero->bytes);
  static int configured[64] = {};
  const int dev = aero->device;
  if (configured[dev & 63] < smem) {
    NP_CUDA(cudaFuncSetAttribute(f16_coeffs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev & 63] = smem;
  }
  const int want = ((n + 1) / 2 + kAuxBS - 1) / kAuxBS;
  f16_coeffs_kernel<<<want < 296 ? want : 296, kAuxBS, smem, (cudaStream_t)stream>>>(aero->image_dev, aero->bytes, alpha_deg_dev,
                                                                                  beta_deg_dev, el_deg_dev, out_dev, n, ld);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_f16_update(const np_aero* aero, float* s_dev, float* u_dev, float* recent_s_dev, float* recent_u_dev, const float* action_dev,
                  int n, int ld, double dt, void* stream) {
  if (!aero || !s_dev || !u_dev || !action_dev || n <= 0 || ld < n + (n & 1) || (ld & 1))
    return fail(NP_EINVAL, "np_f16_update: bad argument (need even ld >= n rounded up to even)");
  if ((((uintptr_t)s_dev | (uintptr_t)u_dev | (uintptr_t)recent_s_dev | (uintptr_t)recent_u_dev) & 7) || ((uintptr_t)action_dev & 15))
    return fail(NP_EINVAL, "np_f16_update: rows 8-byte, action 16-byte aligned");
  DeviceGuard guard(aero->device);
  const int smem = aux_smem_bytes(aero->bytes);
  static int configured[64] = {};
  const int dev = aero->device;
  if (configured[dev & 63] < smem) {
    NP_CUDA(cudaFuncSetAttribute(f16_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev & 63] = smem;
  }
  const int want = ((n + 1) / 2 + kAuxBS - 1) / kAuxBS;
  f16_update_kernel<<<want < 296 ? want : 296, kAuxBS, smem, (cudaStream_t)stream>>>(aero->image_dev, aero->bytes, s_dev, u_dev, recent_s_dev,
                                                                                  recent_u_dev, action_dev, n, ld, (float)dt);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_f16_table_update(const np_tables* tables, float* s_dev, float* u_dev, float* recent_s_dev, float* recent_u_dev,
                        const float* action_dev, int n, int ld, double dt, void* stream) {
  if (!tables || !s_dev || !u_dev || !action_dev || n <= 0 || ld < n || ((uintptr_t)action_dev & 15))
    return fail(NP_EINVAL, "np_f16_table_update: bad argument");
  DeviceGuard guard(tables->device);
  const int smem = kTablesFloats * 4 + 16;
  static int configured[64] = {};
  const int dev = tables->device;
  if (!configured[dev & 63]) {
    NP_CUDA(cudaFuncSetAttribute(f16_table_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev & 63] = 1;
  }
  const int want = (n + 255) / 256;
  f16_table_update_kernel<<<want < 444 ? want : 444, 256, smem, (cudaStream_t)stream>>>(tables->image_dev, s_dev, u_dev, recent_s_dev,
                                                                                    recent_u_dev, action_dev, n, ld, (float)dt);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_uav_update(float* s_dev, float* u_dev, float* recent_s_dev, const float* action_dev, int n, int ld, double dt, void* stream) {
  if (!s_dev || !u_dev || !action_dev || n <= 0 || ld < n || ((uintptr_t)action_dev & 15)) return fail(NP_EINVAL, "np_uav_update: bad argument");
  DeviceGuard guard(device_of(s_dev));
  const int want = (n + 255) / 256;
  uav_update_kernel<<<want < 1184 ? want : 1184, 256, 0, (cudaStream_t)stream>>>(s_dev, u_dev, recent_s_dev, action_dev, n, ld, (float)dt);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

// GPUVecEnv.step with HOST buffers and no copy engine in the path: the step kernel reads the actions from, and writes the
// observation rows / rewards / flags straight into, MAPPED page-locked host memory.  The kernel is compute-bound at
// ~0.46 ms per 10^6 aircraft while the 95 B/aircraft need ~1.7 ms of PCIe, so the posted writes (whole 128-B lines, see
// kObsTileFloats) drain under the arithmetic and the chunked upload / kernel / download pipeline of np_env_step_host
// disappears.  action_dev != NULL selects an explicit H2D copy of the actions before the launch instead of zero-copy reads.
int np_env_step_mapped(np_env* env, const float* action_host, float* action_mapped, float* action_dev, float* obs_mapped,
                       float* reward_mapped, uint8_t* flags_mapped, int flags_ld, void* stream) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_step_mapped: env not bound");
  if (env->cfg.model != NP_MODEL_F16) return fail(NP_EINVAL, "np_env_step_mapped: F16 plug-in only (the UAV step is TMA-staged: use np_env_step_host)");
  const int n = env->cfg.n;
  if (!action_host || !action_mapped || !obs_mapped || !reward_mapped || !flags_mapped || flags_ld < n + (n & 1) || (flags_ld & 1))
    return fail(NP_EINVAL, "np_env_step_mapped: null buffer or bad flags pitch (even, >= n rounded up to even)");
  DeviceGuard guard(env->device);
  float *act_d = nullptr, *obs_d = nullptr, *rew_d = nullptr;
  uint8_t* flg_d = nullptr;
  NP_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&act_d), action_mapped, 0));
  NP_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&obs_d), obs_mapped, 0));
  NP_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&rew_d), reward_mapped, 0));
  NP_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&flg_d), flags_mapped, 0));
  if (((uintptr_t)act_d & 15) || ((uintptr_t)obs_d & 15) || ((uintptr_t)rew_d & 7) || ((uintptr_t)flg_d & 1) || ((uintptr_t)action_dev & 15))
    return fail(NP_EINVAL, "np_env_step_mapped: action / obs 16-byte, reward 8-byte, flags 2-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (action_host != action_mapped) memcpy(action_mapped, action_host, (size_t)n * NP_NUM_ACT * sizeof(float));
  const float* act = act_d;
  if (action_dev) {
    NP_CUDA(cudaMemcpyAsync(action_dev, action_mapped, (size_t)n * NP_NUM_ACT * sizeof(float), cudaMemcpyHostToDevice, st));
    act = action_dev;
  }
  // the ordinary step, with the env's obs / reward outputs and a flag mirror pointed at the mapped host buffers for its duration
  struct Redirect {
    np_env* e;
    float *obs, *rew;
    ~Redirect() { e->buf.obs_dev = obs; e->buf.reward_dev = rew; e->mirror = nullptr; }
  } redirect{env, env->buf.obs_dev, env->buf.reward_dev};
  env->buf.obs_dev = obs_d;
  env->buf.reward_dev = rew_d;
  env->mirror = flg_d;
  env->mirror_ld = flags_ld;
  const int rc = step_range_impl(env, act, nullptr, nullptr, 0, n, true, st);
  if (rc != NP_OK) return rc;
  NP_CUDA(cudaStreamSynchronize(st));
  return NP_OK;
}

}  // extern "C"
