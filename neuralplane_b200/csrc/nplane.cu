// libnplane.so -- kernels and C ABI (include/nplane.h) of the B200 F-16 flight-dynamics step.
//
// One persistent kernel launch per BaseEnv.step() (reference: envs/env_base.py:99-109):
//   masked episodic reset -> control low-pass -> nlplant(s,u') -> explicit Euler -> step_count -> 22-D obs
//   -> nlplant(s',u') for the Overload check -> six termination predicates -> reward -> stores.
// Layout: SoA state rows [F][ld] (coalesced 128 B per warp per field), AoS obs rows staged in shared memory and
// written with one cp.async.bulk (TMA 1-D bulk store) per warp, the 43-net weight blob TMA-bulk-loaded into
// shared memory once per persistent CTA and read as warp-broadcast LDS.128.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (see f16_device.cuh for why).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/nplane.h"
#include "f16_device.cuh"

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
constexpr int kC0Off = kBlobFloats;                 // 36 (alpha,beta)-net outputs at alpha = beta = 0
constexpr int kAeroFloats = kBlobFloats + pad4(kNumAB);
constexpr int kAeroBytes = kAeroFloats * 4;
constexpr int kCacheRows = kNumAB + 2;              // + alpha key, beta key
static_assert(kAeroBytes % 16 == 0, "bulk copy granularity");

struct np_aero {
  float* blob_dev = nullptr;  // kAeroFloats
  int device = 0;
};

struct np_env {
  np_env_cfg cfg;
  const np_aero* aero = nullptr;
  np_buffers buf;
  bool bound = false;
  uint32_t step_index = 0;
  int block = 512, grid = 0, smem = 0, num_sms = 0;
};

struct StepParams {
  np_env_cfg cfg;
  float* s;
  float* u;
  float* tgt;
  int32_t* step_count;
  uint8_t* flags;
  float* obs;
  float* reward;
  float* cache;                  // [kCacheRows][ld]
  unsigned long long* counters;  // [NP_NUM_COUNTERS]
  const float* aero;             // kAeroFloats, global
  const float* action;           // [n][4]
  const float* draws;            // [n][5] or null
  const float* noise;            // [n][22] or null
  uint32_t step_index;
};

// ------------------------------------------------------------------------------------------------
// small PTX wrappers (TMA 1-D bulk copies + mbarrier)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Stage the aero blob into shared memory once per CTA: one elected thread issues TMA bulk copies that
// complete on an mbarrier; everyone waits on it.
__device__ __forceinline__ void stage_aero(float* blob_s, const float* aero_g, uint64_t* bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, kAeroBytes);
    constexpr uint32_t kChunk = 16384;
    for (uint32_t off = 0; off < (uint32_t)kAeroBytes; off += kChunk) {
      const uint32_t nb = min(kChunk, (uint32_t)kAeroBytes - off);
      bulk_g2s(reinterpret_cast<char*>(blob_s) + off, reinterpret_cast<const char*>(aero_g) + off, nb, bar);
    }
  }
  __syncthreads();
  mbar_wait(bar, 0);
}

// ------------------------------------------------------------------------------------------------
// shared pieces of reset / obs / task logic
// ------------------------------------------------------------------------------------------------
struct Draws {
  float d[NP_NUM_DRAWS];
};
__device__ __forceinline__ Draws reset_draws(const StepParams& p, int i) {
  Draws r;
  if (p.draws) {
#pragma unroll
    for (int j = 0; j < NP_NUM_DRAWS; ++j) r.d[j] = p.draws[(size_t)i * NP_NUM_DRAWS + j];
  } else {
    const uint64_t gi = p.cfg.index_base + (uint64_t)i;
    const uint2 key = make_uint2((uint32_t)p.cfg.seed, (uint32_t)(p.cfg.seed >> 32));
    const uint4 a = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), p.step_index, 0x5EED0000u), key);
    const uint4 b = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), p.step_index, 0x5EED0001u), key);
    r.d[0] = u01(a.x); r.d[1] = u01(a.y); r.d[2] = u01(a.z); r.d[3] = u01(a.w); r.d[4] = u01(b.x);
  }
  return r;
}

// F16Model.reset (F16_model.py:38-45) + task.reset (heading_task.py:63-69, control_task.py:59-68,
// tracking_task.py:57-71) for one aircraft.
__device__ __forceinline__ void reset_aircraft(const np_env_cfg& c, const Draws& r, float* s, float* u, float* tgt) {
#pragma unroll
  for (int j = 0; j < 12; ++j) s[j] = 0.0f;
  s[2] = r.d[0] * (c.max_altitude - c.min_altitude) + c.min_altitude;
  s[6] = r.d[1] * (c.max_vt - c.min_vt) + c.min_vt;
  u[0] = c.init_T; u[1] = 0.0f; u[2] = 0.0f; u[3] = 0.0f;
  if (c.task == NP_TASK_HEADING) {
    tgt[0] = s[2] + 1000.0f;
    tgt[1] = wrap_pi(s[5] + (float)(2.0 * 3.141592653589793 / 3.0));
    tgt[2] = s[6] + 0.0f;
  } else if (c.task == NP_TASK_CONTROL) {
    tgt[0] = wrap_pi(s[4] + 2.0f * (r.d[2] - 0.5f) * c.max_pitch_increment);
    tgt[1] = wrap_pi(s[5] + 2.0f * (r.d[3] - 0.5f) * c.max_heading_increment);
    tgt[2] = s[6] + 2.0f * (r.d[4] - 0.5f) * c.max_velocities_u_increment;
  } else {
    const float dist = r.d[2] * (c.max_distance - c.min_distance) + c.min_distance;
    const float th1 = r.d[3] * kPi / 3.0f - (float)(3.141592653589793 / 6.0);
    const float th2 = r.d[4] * kPi / 3.0f - (float)(3.141592653589793 / 6.0);
    tgt[0] = s[0] + dist * cosf(th1) * cosf(th2);
    tgt[1] = s[1] + dist * cosf(th1) * sinf(th2);
    tgt[2] = s[2] + dist * sinf(th1);
  }
}

// 22-D observation row (heading_task.py:113-151; control_task.py:109-111; tracking_task.py:112-114).
__device__ __forceinline__ void make_obs(const np_env_cfg& c, const float* s, const float* u, const float* tgt,
                                         const Trig& g, float e2t, float* o) {
  if (c.task == NP_TASK_HEADING) {
    o[0] = (s[2] - tgt[0]) * 0.3048f / 1000.0f;
    o[1] = wrap_pi(s[5] - tgt[1]);
    o[2] = (s[6] - tgt[2]) * 0.3048f / 340.0f;
  } else if (c.task == NP_TASK_CONTROL) {
    o[0] = wrap_pi(s[4] - tgt[0]);
    o[1] = wrap_pi(s[5] - tgt[1]);
    o[2] = (s[6] - tgt[2]) * 0.3048f / 340.0f;
  } else {
    o[0] = (s[0] - tgt[0]) * 0.3048f / 1000.0f;
    o[1] = (s[1] - tgt[1]) * 0.3048f / 1000.0f;
    o[2] = (s[2] - tgt[2]) * 0.3048f / 1000.0f;
  }
  const float eas = (s[6] + c.airspeed * 1.0f) / e2t;  // F16_model.py:96-103
  o[3] = s[2] * 0.3048f / 5000.0f;
  o[4] = g.sphi; o[5] = g.cphi; o[6] = g.st; o[7] = g.ct;
  o[8] = eas * 0.3048f / 340.0f;
  o[9] = g.sa; o[10] = g.ca; o[11] = g.sb; o[12] = g.cb;
  o[13] = s[9]; o[14] = s[10]; o[15] = s[11];
  o[16] = u[0] / 0.225f / 76300.0f * 0.3048f;
  o[17] = u[1] / 45.0f; o[18] = u[2] / 45.0f; o[19] = u[3] / 45.0f;
  o[20] = 0.0f / 45.0f;  // lef
  o[21] = e2t;
}

__device__ __forceinline__ void add_obs_noise(const StepParams& p, int i, float* o) {
  const float sc = p.cfg.noise_scale;
  if (p.noise) {  // injected standard normals (parity runs): obs + randn * noise_scale (heading_task.py:152)
#pragma unroll
    for (int j = 0; j < NP_NUM_OBS; ++j) o[j] = o[j] + p.noise[(size_t)i * NP_NUM_OBS + j] * sc;
  } else if (sc != 0.0f) {
    const uint64_t gi = p.cfg.index_base + (uint64_t)i;
    const uint2 key = make_uint2((uint32_t)p.cfg.seed, (uint32_t)(p.cfg.seed >> 32));
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const uint4 r = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), p.step_index, 0x0B5E0000u + q), key);
      float n0, n1, n2, n3;
      box_muller(r.x, r.y, n0, n1);
      box_muller(r.z, r.w, n2, n3);
      o[4 * q + 0] = o[4 * q + 0] + n0 * sc;
      o[4 * q + 1] = o[4 * q + 1] + n1 * sc;
      if (4 * q + 2 < NP_NUM_OBS) o[4 * q + 2] = o[4 * q + 2] + n2 * sc;
      if (4 * q + 3 < NP_NUM_OBS) o[4 * q + 3] = o[4 * q + 3] + n3 * sc;
    }
  }
}

// Write a block's obs rows: staged [BS][22] in smem -> one TMA bulk store per fully-populated warp.
template <int BS>
__device__ __forceinline__ void store_obs(float* __restrict__ obs_g, float* stage, const float* o, int i, int n,
                                          bool active) {
  const int lane = threadIdx.x & 31;
  const int warp_first = i - lane;  // aircraft index of lane 0
  float* wstage = stage + (threadIdx.x - lane) * NP_NUM_OBS;
  if (warp_first + 32 <= n) {
#pragma unroll
    for (int j = 0; j < NP_NUM_OBS; ++j) wstage[lane * NP_NUM_OBS + j] = o[j];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA engine
    __syncwarp();
    if (lane == 0) bulk_s2g(obs_g + (size_t)warp_first * NP_NUM_OBS, wstage, 32 * NP_NUM_OBS * 4);
  } else if (active) {
#pragma unroll
    for (int j = 0; j < NP_NUM_OBS; ++j) obs_g[(size_t)i * NP_NUM_OBS + j] = o[j];
  }
}

__device__ __forceinline__ void count_cause(unsigned long long* counters, int which, bool pred) {
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m != 0 && (threadIdx.x & 31) == 0) atomicAdd(&counters[which], (unsigned long long)__popc(m));
}

// ------------------------------------------------------------------------------------------------
// K1: the fused step kernel
// ------------------------------------------------------------------------------------------------
template <int BS>
constexpr int step_smem_bytes() {
  return kAeroBytes + kNumUsed * BS * 4 + BS * NP_NUM_OBS * 4 + 16;
}

template <int BS, int MINB, bool CACHE>
__global__ void __launch_bounds__(BS, MINB) f16_step_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* blob = reinterpret_cast<float*>(smem_raw);
  float* coef_all = blob + kAeroFloats;                 // [kNumUsed][BS]
  float* ostage = coef_all + kNumUsed * BS;             // [BS][22]
  uint64_t* bar = reinterpret_cast<uint64_t*>(ostage + BS * NP_NUM_OBS);

  stage_aero(blob, p.aero, bar);
  const uint32_t wb0 = aero_base_after_staging(blob);

  const np_env_cfg& c = p.cfg;
  const int n = c.n, ld = c.ld;
  float* coef = coef_all + threadIdx.x;                 // coefficient k of this thread: coef[k * BS]
  bool obs_pending = false;

  for (int base = blockIdx.x * BS; base < n; base += gridDim.x * BS) {
    const int i = base + threadIdx.x;
    const bool active = i < n;
    const int il = active ? i : n - 1;  // inactive lanes shadow the last aircraft and never store
    uint32_t wb = opaque_u32(wb0);

    // ---- load ----------------------------------------------------------------------------------
    float s[12], u[4], tgt[3], a[4];
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = p.s[(size_t)j * ld + il];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = p.u[(size_t)j * ld + il];
#pragma unroll
    for (int j = 0; j < 3; ++j) tgt[j] = p.tgt[(size_t)j * ld + il];
    int steps = p.step_count[il];
    const bool rst = (p.flags[il] | p.flags[ld + il] | p.flags[2 * ld + il]) != 0;
    {
      const float4 av = reinterpret_cast<const float4*>(p.action)[il];
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
    }

    // ---- episodic reset of terminated aircraft (env_base.py:83-97) --------------------------------
    if (rst) {
      const Draws r = reset_draws(p, il);
      reset_aircraft(c, r, s, u, tgt);
      steps = 0;
    }
    count_cause(p.counters, 7, rst && active);

    // ---- control low-pass (F16_model.py:52-57) ---------------------------------------------------
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = fminf(fmaxf(a[j], -1.0f), 1.0f);
    u[0] = 0.9f * u[0] + 0.1f * a[0] * 0.225f * 76300.0f / 0.3048f;
    u[1] = 0.9f * u[1] + 0.1f * a[1] * 45.0f;
    u[2] = 0.9f * u[2] + 0.1f * a[2] * 45.0f;
    u[3] = 0.9f * u[3] + 0.1f * a[3] * 45.0f;

    // ---- coefficients at (s, u') -------------------------------------------------------------------
    ZIn zi;
    zscores_el(blob, u[1], zi);
    {
      bool need_eval = true;
      if (CACHE) {
        if (rst) {  // alpha = beta = 0 after a reset: constants precomputed at np_aero_create
#pragma unroll 4
          for (int k = 0; k < kNumAB; ++k) coef[(kFirstAB + k) * BS] = blob[kC0Off + k];
          need_eval = false;
        } else {
          const float ka = p.cache[(size_t)kNumAB * ld + il], kb = p.cache[(size_t)(kNumAB + 1) * ld + il];
          if (__float_as_uint(ka) == __float_as_uint(s[7]) && __float_as_uint(kb) == __float_as_uint(s[8])) {
#pragma unroll 4
            for (int k = 0; k < kNumAB; ++k) coef[(kFirstAB + k) * BS] = p.cache[(size_t)k * ld + il];
            need_eval = false;
          }
        }
      }
      if (need_eval) {  // first step after external state writes, or CACHE == false
        zscores_ab(blob, s[7] * kR2D, s[8] * kR2D, zi);
        eval_ab_nets(blob, wb, zi, coef, BS);
      } else {
        const float* zn = blob + kZnormOff;
        zi.z[kZaC] = (s[7] * kR2D - zn[2 * kZaC]) / zn[2 * kZaC + 1];
        zi.z[kZbC] = (s[8] * kR2D - zn[2 * kZbC]) / zn[2 * kZbC + 1];
      }
    }
    eval_el_nets(blob, wb, zi, coef, BS);

    // ---- Euler step (F16_model.py:64-67; torchdiffeq fixed-grid euler on t=[0,dt]) -----------------
    {
      const Trig g = make_trig(s);
      const float tp = tfac_pow(s[2]);
      float xdot[12];
      nlplant_from_coefs(s, u[0], u[2], u[3], 0.0f, g, tp, coef, BS, xdot);
      const float h = c.dt - 0.0f;
#pragma unroll
      for (int j = 0; j < 12; ++j) s[j] = s[j] + h * xdot[j];
    }
    steps += 1;  // env_base.py:102

    // ---- observation of the new state (env_base.py:103) -------------------------------------------
    const Trig g2 = make_trig(s);
    const float tp2 = tfac_pow(s[2]);
    {
      float o[NP_NUM_OBS];
      make_obs(c, s, u, tgt, g2, eas2tas_of(tp2), o);
      add_obs_noise(p, il, o);
      if (obs_pending) {  // the previous slab's bulk store must have finished reading the staging rows
        if ((threadIdx.x & 31) == 0) bulk_wait_read();
        __syncwarp();
      }
      store_obs<BS>(p.obs, ostage, o, i, n, active);
      obs_pending = true;
    }

    // ---- coefficients at (s', u'): Overload check now, Euler derivative of the next step later ---------
    wb = opaque_u32(wb0);
    zscores_ab(blob, s[7] * kR2D, s[8] * kR2D, zi);
    eval_ab_nets(blob, wb, zi, coef, BS);
    if (CACHE && active) {
#pragma unroll 4
      for (int k = 0; k < kNumAB; ++k) p.cache[(size_t)k * ld + i] = coef[(kFirstAB + k) * BS];
      p.cache[(size_t)kNumAB * ld + i] = s[7];
      p.cache[(size_t)(kNumAB + 1) * ld + i] = s[8];
    }
    eval_el_force_nets(blob, wb, zi, coef, BS);

    // ---- terminations (task_base.py:75-96) --------------------------------------------------------
    bool bad, done;
    {
      const float vt_c = s[6] <= 0.01f ? 0.01f : s[6];
      const AeroTotals t = force_totals(coef, BS, vt_c, s[9], s[10], s[11], u[2] / 21.5f, u[3] / 30.0f, 1.0f);
      const ForceOut f = force_eqs(t, body_vel(vt_c, g2), g2, vt_c, s[9], s[10], s[11], qbar_of(tp2, vt_c), u[0]);
      float ax, ay, az;
      body_accel(s, g2, f, ax, ay, az);
      const float acc = sqrtf(ax * ax + ay * ay + az * az);
      const bool overload = (acc - c.acceleration_limit) > 0.0f;            // overload.py:37-42
      const bool low_alt = (s[2] - c.altitude_limit) < 0.0f;                // low_altitude.py:29-30
      const float vel = (s[6] + c.airspeed * 1.0f) * 0.3048f / 340.0f;
      const bool hi = (vel - c.max_velocity) >= 0.0f;                       // high_speed.py:29-30
      const bool lo = (vel - c.min_velocity) <= 0.0f;                       // low_speed.py:29-30
      const float a_deg = s[7] * 180.0f / kPi, b_deg = s[8] * 180.0f / kPi; // extreme_state.py:32-36
      const bool ext = (a_deg < c.min_alpha) | (a_deg > c.max_alpha) | (b_deg < c.min_beta) | (b_deg > c.max_beta);
      const bool late = steps >= c.max_check_interval;
      bool off;
      if (c.task == NP_TASK_HEADING) {                                      // unreach_heading.py:38-53
        off = (fabsf(wrap_pi(s[5] - tgt[1])) >= (float)(3.141592653589793 / 36.0)) | (fabsf(s[2] - tgt[0]) >= 100.0f) |
              (fabsf(s[6] - tgt[2]) >= 20.0f);
        done = !off && !late && (steps >= c.min_check_interval);
      } else if (c.task == NP_TASK_CONTROL) {                               // unreach_posture.py:37-55
        off = (fabsf(wrap_pi(s[5] - tgt[1])) >= (float)(3.141592653589793 / 36.0)) |
              (fabsf(s[4] - tgt[0]) >= (float)(3.141592653589793 / 36.0)) | (fabsf(s[6] - tgt[2]) >= 20.0f);
        done = !off && !late;
      } else {                                                              // unreach_target.py:35-47
        off = (fabsf(s[0] - tgt[0]) >= 100.0f) | (fabsf(s[1] - tgt[1]) >= 100.0f) | (fabsf(s[2] - tgt[2]) >= 100.0f);
        done = !off && !late;
      }
      const bool unreach = late && off;
      bad = overload | low_alt | hi | lo | ext | unreach;
      count_cause(p.counters, 0, overload && active);
      count_cause(p.counters, 1, low_alt && active);
      count_cause(p.counters, 2, hi && active);
      count_cause(p.counters, 3, lo && active);
      count_cause(p.counters, 4, ext && active);
      count_cause(p.counters, 5, unreach && active);
      count_cause(p.counters, 6, done && active);
    }

    // ---- reward (task_base.py:60-73) --------------------------------------------------------------
    float rew;
    {
      float d0, d1, d2;
      if (c.task == NP_TASK_HEADING) {                                      // heading_reward.py:26-35
        d0 = (s[2] - tgt[0]) * 0.3048f / 1000.0f;
        d1 = wrap_pi(s[5] - tgt[1]) / kPi;
        d2 = (s[6] - tgt[2]) * 0.3048f / 340.0f;
        rew = -(d0 * d0) + -(d1 * d1) + -(d2 * d2);
      } else if (c.task == NP_TASK_CONTROL) {                               // posture_reward.py:26-34
        d0 = wrap_pi(s[4] - tgt[0]) / kPi;
        d1 = wrap_pi(s[5] - tgt[1]) / kPi;
        d2 = (s[6] - tgt[2]) * 0.3048f / 340.0f;
        rew = -(d0 * d0) + -(d1 * d1) + -(d2 * d2);
      } else {                                                              // position_reward.py:26-34
        d0 = (s[0] - tgt[0]) * 0.3048f / 1000.0f;
        d1 = (s[1] - tgt[1]) * 0.3048f / 1000.0f;
        d2 = (s[2] - tgt[2]) * 0.3048f / 1000.0f;
        rew = 0.1f * (-(d0 * d0) + -(d1 * d1) + -(d2 * d2));
      }
      rew = rew + (float)(-200 * (int)bad + 200 * (int)done);               // event_driven_reward.py:28
    }

    // ---- store ----------------------------------------------------------------------------------
    if (active) {
#pragma unroll
      for (int j = 0; j < 12; ++j) p.s[(size_t)j * ld + i] = s[j];
#pragma unroll
      for (int j = 0; j < 4; ++j) p.u[(size_t)j * ld + i] = u[j];
#pragma unroll
      for (int j = 0; j < 3; ++j) p.tgt[(size_t)j * ld + i] = tgt[j];
      p.step_count[i] = steps;
      p.flags[i] = done ? 1 : 0;
      p.flags[ld + i] = bad ? 1 : 0;
      p.flags[2 * ld + i] = 0;  // Timeout is commented out of the control tasks (heading_task.py:45)
      p.reward[i] = rew;
    }
  }
  if (obs_pending && (threadIdx.x & 31) == 0) bulk_wait_read();
}

// ------------------------------------------------------------------------------------------------
// K3: standalone reset (BaseEnv.reset, env_base.py:83-97)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) f16_reset_kernel(const __grid_constant__ StepParams p) {
  const np_env_cfg& c = p.cfg;
  const int n = c.n, ld = c.ld;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s[12], u[4], tgt[3];
    const bool rst = (p.flags[i] | p.flags[ld + i] | p.flags[2 * ld + i]) != 0;
    if (rst) {
      const Draws r = reset_draws(p, i);
      reset_aircraft(c, r, s, u, tgt);
#pragma unroll
      for (int j = 0; j < 12; ++j) p.s[(size_t)j * ld + i] = s[j];
#pragma unroll
      for (int j = 0; j < 4; ++j) p.u[(size_t)j * ld + i] = u[j];
      p.u[(size_t)4 * ld + i] = 0.0f;
#pragma unroll
      for (int j = 0; j < 3; ++j) p.tgt[(size_t)j * ld + i] = tgt[j];
      p.step_count[i] = 0;
      for (int k = 0; k < kNumAB; ++k) p.cache[(size_t)k * ld + i] = p.aero[kC0Off + k];
      p.cache[(size_t)kNumAB * ld + i] = 0.0f;
      p.cache[(size_t)(kNumAB + 1) * ld + i] = 0.0f;
      atomicAdd(&p.counters[7], 1ull);
    } else {
#pragma unroll
      for (int j = 0; j < 12; ++j) s[j] = p.s[(size_t)j * ld + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = p.u[(size_t)j * ld + i];
#pragma unroll
      for (int j = 0; j < 3; ++j) tgt[j] = p.tgt[(size_t)j * ld + i];
    }
    p.flags[i] = 0; p.flags[ld + i] = 0; p.flags[2 * ld + i] = 0;
    const Trig g = make_trig(s);
    float o[NP_NUM_OBS];
    make_obs(c, s, u, tgt, g, eas2tas_of(tfac_pow(s[2])), o);
    add_obs_noise(p, i, o);
#pragma unroll
    for (int j = 0; j < NP_NUM_OBS; ++j) p.obs[(size_t)i * NP_NUM_OBS + j] = o[j];
  }
}

// ------------------------------------------------------------------------------------------------
// stand-alone nlplant / coefficient kernels (model plug-in getters, parity tests)
// ------------------------------------------------------------------------------------------------
constexpr int kAuxBS = 128;
constexpr int kAuxSmem = kAeroBytes + kNumNets * kAuxBS * 4 + 16;

__global__ void __launch_bounds__(kAuxBS) f16_nlplant_kernel(const float* __restrict__ aero, const float* __restrict__ S,
                                                             const float* __restrict__ U, float* __restrict__ X, int n,
                                                             int ld) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* blob = reinterpret_cast<float*>(smem_raw);
  float* coef_all = blob + kAeroFloats;
  uint64_t* bar = reinterpret_cast<uint64_t*>(coef_all + kNumNets * kAuxBS);
  stage_aero(blob, aero, bar);
  const uint32_t wb0 = aero_base_after_staging(blob);
  float* coef = coef_all + threadIdx.x;
  for (int base = blockIdx.x * kAuxBS; base < n; base += gridDim.x * kAuxBS) {
    const int i = base + threadIdx.x;
    const int il = i < n ? i : n - 1;
    const uint32_t wb = opaque_u32(wb0);
    float s[12], u[5];
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = S[(size_t)j * ld + il];
#pragma unroll
    for (int j = 0; j < 5; ++j) u[j] = U[(size_t)j * ld + il];
    ZIn zi;
    zscores_ab(blob, s[7] * kR2D, s[8] * kR2D, zi);
    zscores_el(blob, u[1], zi);
    eval_el_nets(blob, wb, zi, coef, kAuxBS);
    eval_ab_nets(blob, wb, zi, coef, kAuxBS);
    const Trig g = make_trig(s);
    float xdot[12];
    nlplant_from_coefs(s, u[0], u[2], u[3], u[4], g, tfac_pow(s[2]), coef, kAuxBS, xdot);
    if (i < n) {
#pragma unroll
      for (int j = 0; j < 12; ++j) X[(size_t)j * ld + i] = xdot[j];
    }
  }
}

__global__ void __launch_bounds__(kAuxBS) f16_coeffs_kernel(const float* __restrict__ aero, const float* __restrict__ A,
                                                            const float* __restrict__ Bd, const float* __restrict__ E,
                                                            float* __restrict__ out, int n, int ld) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* blob = reinterpret_cast<float*>(smem_raw);
  float* coef_all = blob + kAeroFloats;
  uint64_t* bar = reinterpret_cast<uint64_t*>(coef_all + kNumNets * kAuxBS);
  stage_aero(blob, aero, bar);
  const uint32_t wb0 = aero_base_after_staging(blob);
  float* coef = coef_all + threadIdx.x;
  for (int base = blockIdx.x * kAuxBS; base < n; base += gridDim.x * kAuxBS) {
    const int i = base + threadIdx.x;
    const int il = i < n ? i : n - 1;
    const uint32_t wb = opaque_u32(wb0);
    ZIn zi;
    zscores_ab(blob, A[il], Bd[il], zi);
    zscores_el(blob, E[il], zi);
    eval_el_nets(blob, wb, zi, coef, kAuxBS);
    eval_ab_nets(blob, wb, zi, coef, kAuxBS);
    eval_group<kdCzq_lef, kdCzq_lef + 1>(blob, wb, zi, coef, kAuxBS);
    if (i < n) {
      for (int k = 0; k < kNumNets; ++k) out[(size_t)k * ld + i] = coef[k * kAuxBS];
    }
  }
}

// (alpha,beta)-net outputs at alpha = beta = 0, appended to the device blob at np_aero_create.
__global__ void __launch_bounds__(128) f16_c0_kernel(float* aero) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* blob = reinterpret_cast<float*>(smem_raw);
  float* coef = blob + kAeroFloats;
  uint64_t* bar = reinterpret_cast<uint64_t*>(coef + pad4(kNumNets));  // 8-byte aligned
  stage_aero(blob, aero, bar);
  const uint32_t wb = aero_base_after_staging(blob);
  if (threadIdx.x == 0) {
    ZIn zi;
    zscores_ab(blob, 0.0f * kR2D, 0.0f * kR2D, zi);
    eval_ab_nets(blob, wb, zi, coef, 1);
    for (int k = 0; k < kNumAB; ++k) aero[kC0Off + k] = coef[kFirstAB + k];
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

int np_aero_create(const float* blob, size_t n_floats, const np_net_desc* descs, const double* norm, int n_nets,
                   np_aero** out) {
  if (!blob || !descs || !norm || !out) return fail(NP_EINVAL, "np_aero_create: null argument");
  if (n_nets != kNumNets) return fail(NP_EINVAL, "np_aero_create: expected 43 nets");
  std::vector<float> host(kAeroFloats, 0.0f);
  bool zset[kNumZ] = {};
  auto set_z = [&](int zid, double mean, double sd, std::string* err) {
    const float m = (float)mean, s = (float)sd;
    if (zset[zid] && (host[kZnormOff + 2 * zid] != m || host[kZnormOff + 2 * zid + 1] != s))
      *err = "np_aero_create: nets of one normalisation group disagree on (mean, std)";
    host[kZnormOff + 2 * zid] = m;
    host[kZnormOff + 2 * zid + 1] = s;
    zset[zid] = true;
  };
  for (int k = 0; k < kNumNets; ++k) {
    const np_net_desc& d = descs[k];
    const NetArch a = arch_of(k);
    const ZSel z = zsel_of(k);
    const int nl = a.h3 ? 4 : 3;
    const int dims[5] = {a.nin, a.h1, a.h2, a.h3 ? a.h3 : 1, a.h3 ? 1 : 0};
    if (d.n_in != a.nin || d.n_layers != nl) return fail(NP_EINVAL, "np_aero_create: net architecture mismatch (depth)");
    for (int l = 0; l <= nl; ++l)
      if (d.dims[l] != dims[l]) return fail(NP_EINVAL, "np_aero_create: net architecture mismatch (width)");
    const int want_sel[3] = {z.a >= 0 ? 0 : 2, a.nin >= 2 ? 1 : -1, a.nin == 3 ? 2 : -1};
    for (int j = 0; j < a.nin; ++j)
      if (d.sel[j] != want_sel[j]) return fail(NP_EINVAL, "np_aero_create: net input selection mismatch");
    std::string err;
    const double* nm = norm + 8 * k;
    for (int j = 0; j < a.nin; ++j) {
      const int zid = d.sel[j] == 0 ? z.a : (d.sel[j] == 1 ? z.b : z.e);
      set_z(zid, nm[j], nm[3 + j], &err);
    }
    if (!err.empty()) return fail(NP_EINVAL, err);
    host[kOnormOff + 2 * k] = (float)nm[6];
    host[kOnormOff + 2 * k + 1] = (float)nm[7];
    // weights: source per layer W[out][in] then b[out]; device per layer b[out] then W^T[in][out], padded to 4
    size_t src = (size_t)d.w_off;
    int dst = net_offset(k);
    for (int l = 0; l < nl; ++l) {
      const int in = dims[l], o = dims[l + 1];
      if (src + (size_t)in * o + o > n_floats) return fail(NP_EINVAL, "np_aero_create: blob too short");
      const float* W = blob + src;
      const float* b = W + (size_t)in * o;
      for (int j = 0; j < o; ++j) host[dst + j] = b[j];
      for (int i = 0; i < in; ++i)
        for (int j = 0; j < o; ++j) host[dst + o + i * o + j] = W[(size_t)j * in + i];
      src += (size_t)in * o + o;
      dst += layer_floats(in, o);
    }
  }
  for (int zid = 0; zid < kNumZ; ++zid)
    if (!zset[zid]) return fail(NP_EINVAL, "np_aero_create: normalisation group without nets");
  np_aero* a = new np_aero();
  cudaGetDevice(&a->device);
  NP_CUDA(cudaMalloc(&a->blob_dev, kAeroBytes));
  NP_CUDA(cudaMemcpy(a->blob_dev, host.data(), kAeroBytes, cudaMemcpyHostToDevice));
  constexpr int c0_smem = kAeroBytes + pad4(kNumNets) * 4 + 16;
  NP_CUDA(cudaFuncSetAttribute(f16_c0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c0_smem));
  f16_c0_kernel<<<1, 128, c0_smem>>>(a->blob_dev);
  NP_CUDA(cudaGetLastError());
  NP_CUDA(cudaDeviceSynchronize());
  *out = a;
  return NP_OK;
}

int np_aero_destroy(np_aero* aero) {
  if (!aero) return NP_OK;
  cudaFree(aero->blob_dev);
  delete aero;
  return NP_OK;
}

size_t np_env_workspace_bytes(const np_env_cfg* cfg) {
  if (!cfg) return 0;
  return (size_t)kCacheRows * (size_t)cfg->ld * sizeof(float) + 256 /* counters, 128-B aligned tail */;
}

}  // extern "C"

template <int BS, int MINB>
static int launch_step(np_env* env, const StepParams& p, cudaStream_t st) {
  constexpr int smem = step_smem_bytes<BS>();
  static bool configured[2] = {false, false};
  const bool cache = env->cfg.use_coef_cache != 0;
  auto kern = cache ? f16_step_kernel<BS, MINB, true> : f16_step_kernel<BS, MINB, false>;
  if (!configured[cache]) {
    NP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[cache] = true;
  }
  const int want = (env->cfg.n + BS - 1) / BS;
  const int grid = want < env->num_sms * MINB ? want : env->num_sms * MINB;
  env->grid = grid;
  env->smem = smem;
  kern<<<grid, BS, smem, st>>>(p);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
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
  p.counters = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(env->buf.workspace_dev) +
                                                     (((size_t)kCacheRows * env->cfg.ld * 4 + 127) / 128) * 128);
  p.aero = env->aero->blob_dev;
  p.action = action;
  p.draws = draws;
  p.noise = noise;
  p.step_index = env->step_index;
  return p;
}

extern "C" {

int np_env_create(const np_env_cfg* cfg, const np_aero* aero, np_env** out) {
  if (!cfg || !aero || !out) return fail(NP_EINVAL, "np_env_create: null argument");
  if (cfg->n <= 0 || cfg->ld < cfg->n || cfg->ld % 4) return fail(NP_EINVAL, "np_env_create: need n > 0, ld >= n, ld % 4 == 0");
  if (cfg->task < NP_TASK_HEADING || cfg->task > NP_TASK_TRACKING) return fail(NP_EINVAL, "np_env_create: unknown task");
  np_env* e = new np_env();
  e->cfg = *cfg;
  e->aero = aero;
  memset(&e->buf, 0, sizeof(e->buf));
  int dev = 0;
  NP_CUDA(cudaGetDevice(&dev));
  NP_CUDA(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, dev));
  e->block = 512;  // 16 warps/SM: measured 1.155 ms vs 1.49 ms (256) per 10^6-aircraft step (profiles/r01_block_sweep.txt)
  if (const char* b = getenv("NPLANE_BLOCK")) e->block = atoi(b);
  if (e->block != 128 && e->block != 256 && e->block != 384 && e->block != 512)
    return fail(NP_EINVAL, "NPLANE_BLOCK must be 128, 256, 384 or 512");
  *out = e;
  return NP_OK;
}

int np_env_bind(np_env* env, const np_buffers* b) {
  if (!env || !b) return fail(NP_EINVAL, "np_env_bind: null argument");
  if (!b->s_dev || !b->u_dev || !b->tgt_dev || !b->step_count_dev || !b->flags_dev || !b->obs_dev || !b->reward_dev ||
      !b->workspace_dev)
    return fail(NP_EINVAL, "np_env_bind: null buffer");
  if (((uintptr_t)b->obs_dev & 15) || ((uintptr_t)b->workspace_dev & 127))
    return fail(NP_EINVAL, "np_env_bind: obs must be 16-byte and workspace 128-byte aligned");
  env->buf = *b;
  env->bound = true;
  // invalidate the coefficient-cache keys: 0xFFFFFFFF is a NaN pattern no stored alpha/beta can equal bitwise
  NP_CUDA(cudaMemset(reinterpret_cast<float*>(b->workspace_dev) + (size_t)kNumAB * env->cfg.ld, 0xFF,
                     2 * (size_t)env->cfg.ld * sizeof(float)));
  return NP_OK;
}

int np_env_set_cfg(np_env* env, const np_env_cfg* cfg) {
  if (!env || !cfg) return fail(NP_EINVAL, "np_env_set_cfg: null argument");
  if (cfg->n != env->cfg.n || cfg->ld != env->cfg.ld || cfg->task != env->cfg.task)
    return fail(NP_EINVAL, "np_env_set_cfg: n, ld and task are fixed at creation");
  env->cfg = *cfg;
  return NP_OK;
}

int np_env_destroy(np_env* env) {
  delete env;
  return NP_OK;
}

int np_env_reset(np_env* env, const float* draws_dev, const float* noise_dev, void* stream) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_reset: env not bound");
  StepParams p = make_params(env, nullptr, draws_dev, noise_dev);
  env->step_index++;
  const int want = (env->cfg.n + 255) / 256;
  const int grid = want < env->num_sms * 8 ? want : env->num_sms * 8;
  f16_reset_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_env_step(np_env* env, const float* action_dev, const float* draws_dev, const float* noise_dev, void* stream) {
  if (!env || !env->bound) return fail(NP_ESTATE, "np_env_step: env not bound");
  if (!action_dev || ((uintptr_t)action_dev & 15)) return fail(NP_EINVAL, "np_env_step: action must be a 16-byte aligned device pointer");
  StepParams p = make_params(env, action_dev, draws_dev, noise_dev);
  env->step_index++;
  cudaStream_t st = (cudaStream_t)stream;
  switch (env->block) {
#ifdef NPLANE_ALL_BLOCKS
    case 128: return launch_step<128, 2>(env, p, st);
    case 256: return launch_step<256, 1>(env, p, st);
    case 384: return launch_step<384, 1>(env, p, st);
#endif
    case 512: return launch_step<512, 1>(env, p, st);
    default: return fail(NP_EINVAL, "np_env_step: block size not compiled in (build with -DNPLANE_ALL_BLOCKS)");
  }
}

int np_env_counters(np_env* env, uint64_t* out, void* stream) {
  if (!env || !env->bound || !out) return fail(NP_ESTATE, "np_env_counters: env not bound");
  StepParams p = make_params(env, nullptr, nullptr, nullptr);
  NP_CUDA(cudaMemcpyAsync(out, p.counters, NP_NUM_COUNTERS * sizeof(uint64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  NP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return NP_OK;
}

int np_env_launch_info(const np_env* env, int* grid, int* block, int* smem_bytes, int* num_sms) {
  if (!env) return fail(NP_EINVAL, "np_env_launch_info: null env");
  if (grid) *grid = env->grid;
  if (block) *block = env->block;
  if (smem_bytes) *smem_bytes = env->smem;
  if (num_sms) *num_sms = env->num_sms;
  return NP_OK;
}

int np_f16_nlplant(const np_aero* aero, const float* s_dev, const float* u_dev, float* xdot_dev, int n, int ld, void* stream) {
  if (!aero || !s_dev || !u_dev || !xdot_dev || n <= 0 || ld < n) return fail(NP_EINVAL, "np_f16_nlplant: bad argument");
  static bool configured = false;
  if (!configured) {
    NP_CUDA(cudaFuncSetAttribute(f16_nlplant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAuxSmem));
    configured = true;
  }
  const int want = (n + kAuxBS - 1) / kAuxBS;
  f16_nlplant_kernel<<<want < 296 ? want : 296, kAuxBS, kAuxSmem, (cudaStream_t)stream>>>(aero->blob_dev, s_dev, u_dev, xdot_dev, n, ld);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

int np_f16_coeffs(const np_aero* aero, const float* alpha_deg_dev, const float* beta_deg_dev, const float* el_deg_dev,
                  float* out_dev, int n, int ld, void* stream) {
  if (!aero || !alpha_deg_dev || !beta_deg_dev || !el_deg_dev || !out_dev || n <= 0 || ld < n)
    return fail(NP_EINVAL, "np_f16_coeffs: bad argument");
  static bool configured = false;
  if (!configured) {
    NP_CUDA(cudaFuncSetAttribute(f16_coeffs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAuxSmem));
    configured = true;
  }
  const int want = (n + kAuxBS - 1) / kAuxBS;
  f16_coeffs_kernel<<<want < 296 ? want : 296, kAuxBS, kAuxSmem, (cudaStream_t)stream>>>(aero->blob_dev, alpha_deg_dev, beta_deg_dev,
                                                                                      el_deg_dev, out_dev, n, ld);
  NP_CUDA(cudaGetLastError());
  return NP_OK;
}

}  // extern "C"
