// Per-aircraft env logic shared by the step kernels (K1, K1t, K3 and the UAV kernels): the kernel parameter block, reset draws,
// BaseEnv.reset / task.reset, the 22-D observation and its noise, the control lag, the six terminations + task reward, and the
// warp-aggregated termination counters.  Reference: envs/env_base.py:83-109, envs/tasks/*.py, envs/termination_conditions/*.py,
// envs/reward_functions/*.py.
#pragma once
#include "../../include/nplane.h"
#include "f16_device.cuh"

using namespace npl;

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
  const uint32_t* aero;          // device image, aero_bytes
  int aero_bytes;
  int tab;                       // 1: `aero` is the table image (tables_device.cuh), not the MLP image
  const float* action;           // [n][4]; planning step: [n][3]
  float* pid;                    // [kPidRows][ld] controller state (planning / combat step)
  float* blood;                  // [ld] combat damage state (singlecombat_env.py:45)
  int n_sub;                     // FDM sub-steps per env step (planning: 50, planning_env.py:153)
  int pid_first;                 // 1: the controllers have never run (PID.reset, pid.py:13)
  int pair_begin, pair_end;      // aircraft pairs [pair_begin, pair_end) this launch works on (whole population by default)
  const float* draws;            // [n][5] or null
  const float* noise;            // [n][22] or null
  int obs_stg;                   // 1: the staged observation tile leaves through per-lane 16-byte stores instead of a TMA bulk store
  uint8_t* flags_mirror;         // null, or a second [3][flags_mirror_ld] copy of the new flags (mapped host memory)
  int flags_mirror_ld;
  // role-sharded combat (egos and opponents on different ranks): this rank's two lanes are two DIFFERENT envs' aircraft
  float* records;                // null (pair-sharded), or this rank's record slab [n][kCombatRecFloats]
  uint8_t* pair_reset;           // [ld] env-level reset flag of each local aircraft's env (own | partner flags of the last step)
  int index_stride;              // global aircraft index = index_base + index_stride * local index (RNG streams)
  uint32_t step_index;
  const uint32_t* rng_epoch;     // device word added to step_index (np_env_rng_advance): fresh RNG streams for each replay of a CUDA graph
};

// The RNG counter of this launch.  step_index is advanced on the host by every step call, i.e. it is a constant of a captured
// CUDA graph; the epoch word lives on the device and is advanced by np_env_rng_advance (capturable), 0 otherwise.
// (read through L2: the word may have been written by the kernel just ahead in the stream)
#ifdef NPLANE_NO_EPOCH   // experiment: what the epoch read costs
__device__ __forceinline__ uint32_t rng_step(const StepParams& p) { return p.step_index; }
#else
__device__ __forceinline__ uint32_t rng_step(const StepParams& p) { return p.step_index + __ldcg(p.rng_epoch); }
#endif


struct Draws {
  float d[NP_NUM_DRAWS];
};
// `step` = rng_step(p); the latency kernel reads it once up front, the others where it is needed (rare / hidden by other warps)
__device__ __forceinline__ Draws reset_draws(const StepParams& p, int i, uint32_t step) {
  Draws r;
  if (p.draws) {
#pragma unroll
    for (int j = 0; j < NP_NUM_DRAWS; ++j) r.d[j] = p.draws[(size_t)i * NP_NUM_DRAWS + j];
  } else {
    const uint64_t gi = p.cfg.index_base + (uint64_t)p.index_stride * (uint64_t)i;
    const uint2 key = make_uint2((uint32_t)p.cfg.seed, (uint32_t)(p.cfg.seed >> 32));
    const uint4 a = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), step, 0x5EED0000u), key);
    const uint4 b = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), step, 0x5EED0001u), key);
    r.d[0] = u01(a.x); r.d[1] = u01(a.y); r.d[2] = u01(a.z); r.d[3] = u01(a.w); r.d[4] = u01(b.x);
  }
  return r;
}

// F16Model.reset (F16_model.py:38-45) + task.reset (heading_task.py:63-69, control_task.py:59-68,
// tracking_task.py:57-71) for one aircraft.
// `task` is a compile-time constant in the step kernel (one instantiation per task: the other two tasks' code -- each with
// its own wrap / trig calls -- would only be instruction-cache ballast) and c.task in the stand-alone reset kernel.
__device__ __forceinline__ void reset_aircraft(const np_env_cfg& c, int task, const Draws& r, float* s, float* u, float* tgt) {
#pragma unroll
  for (int j = 0; j < 12; ++j) s[j] = 0.0f;
  s[2] = r.d[0] * (c.max_altitude - c.min_altitude) + c.min_altitude;
  s[6] = r.d[1] * (c.max_vt - c.min_vt) + c.min_vt;
  u[0] = c.init_T; u[1] = 0.0f; u[2] = 0.0f; u[3] = 0.0f;
  if (task == NP_TASK_HEADING) {
    tgt[0] = s[2] + 1000.0f;
    tgt[1] = wrap_pi(s[5] + (float)(2.0 * 3.141592653589793 / 3.0));
    tgt[2] = s[6] + 0.0f;
  } else if (task == NP_TASK_CONTROL) {
    tgt[0] = wrap_pi(s[4] + 2.0f * (r.d[2] - 0.5f) * c.max_pitch_increment);
    tgt[1] = wrap_pi(s[5] + 2.0f * (r.d[3] - 0.5f) * c.max_heading_increment);
    tgt[2] = s[6] + 2.0f * (r.d[4] - 0.5f) * c.max_velocities_u_increment;
  } else {
    const float dist = r.d[2] * (c.max_distance - c.min_distance) + c.min_distance;
    const float th1 = r.d[3] * kPi / DC(3.0f) - (float)(3.141592653589793 / 6.0);
    const float th2 = r.d[4] * kPi / DC(3.0f) - (float)(3.141592653589793 / 6.0);
    const float2 sc1 = sincos_shared(th1), sc2 = sincos_shared(th2);   // (sin, cos): the bits of sinf / cosf
    tgt[0] = s[0] + dist * sc1.y * sc2.y;
    tgt[1] = s[1] + dist * sc1.y * sc2.x;
    tgt[2] = s[2] + dist * sc1.x;
  }
}

// 22-D observation row (heading_task.py:113-151; control_task.py:109-111; tracking_task.py:112-114).
__device__ __forceinline__ void make_obs(const np_env_cfg& c, int task, const float* s, const float* u, const float* tgt,
                                         const Trig& g, float e2t, float* o) {
  if (task == NP_TASK_HEADING) {
    o[0] = (s[2] - tgt[0]) * 0.3048f / DC(1000.0f);
    o[1] = wrap_pi(s[5] - tgt[1]);
    o[2] = (s[6] - tgt[2]) * 0.3048f / DC(340.0f);
  } else if (task == NP_TASK_CONTROL) {
    o[0] = wrap_pi(s[4] - tgt[0]);
    o[1] = wrap_pi(s[5] - tgt[1]);
    o[2] = (s[6] - tgt[2]) * 0.3048f / DC(340.0f);
  } else {
    o[0] = (s[0] - tgt[0]) * 0.3048f / DC(1000.0f);
    o[1] = (s[1] - tgt[1]) * 0.3048f / DC(1000.0f);
    o[2] = (s[2] - tgt[2]) * 0.3048f / DC(1000.0f);
  }
  const float eas = (s[6] + c.airspeed * 1.0f) / e2t;  // F16_model.py:96-103
  o[3] = s[2] * 0.3048f / DC(5000.0f);
  o[4] = g.sphi; o[5] = g.cphi; o[6] = g.st; o[7] = g.ct;
  o[8] = eas * 0.3048f / DC(340.0f);
  o[9] = g.sa; o[10] = g.ca; o[11] = g.sb; o[12] = g.cb;
  o[13] = s[9]; o[14] = s[10]; o[15] = s[11];
  o[16] = u[0] / DC(0.225f) / DC(76300.0f) * 0.3048f;
  o[17] = u[1] / DC(45.0f); o[18] = u[2] / DC(45.0f); o[19] = u[3] / DC(45.0f);
  o[20] = 0.0f / DC(45.0f);  // lef
  o[21] = e2t;
}

__device__ __forceinline__ Draws reset_draws(const StepParams& p, int i) { return reset_draws(p, i, p.draws ? 0u : rng_step(p)); }

__device__ __forceinline__ void add_obs_noise(const StepParams& p, int i, float* o, uint32_t step) {
  const float sc = p.cfg.noise_scale;
  if (p.noise) {  // injected standard normals (parity runs): obs + randn * noise_scale (heading_task.py:152)
#pragma unroll
    for (int j = 0; j < NP_NUM_OBS; ++j) o[j] = o[j] + p.noise[(size_t)i * NP_NUM_OBS + j] * sc;
  } else if (sc != 0.0f) {
    const uint64_t gi = p.cfg.index_base + (uint64_t)p.index_stride * (uint64_t)i;
    const uint2 key = make_uint2((uint32_t)p.cfg.seed, (uint32_t)(p.cfg.seed >> 32));
#pragma unroll
    for (int q = 0; q < 3; ++q) {  // 3 x 4 words -> 12 pairs of normals, 22 used
      const uint4 r = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), step, 0x0B5E0000u + q), key);
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = 8 * q + 2 * k;
        if (j < NP_NUM_OBS) {
          float r, cs, sn;
          box_muller16(w[k], sc, r, cs, sn);
          o[j] = fmaf(r, cs, o[j]);
          o[j + 1] = fmaf(r, sn, o[j + 1]);
        }
      }
    }
  }
}
__device__ __forceinline__ void add_obs_noise(const StepParams& p, int i, float* o) {
  add_obs_noise(p, i, o, (p.noise || p.cfg.noise_scale == 0.0f) ? 0u : rng_step(p));
}

// one aircraft pair of an SoA row: a single 8-byte store, or only the first aircraft for the odd tail
__device__ __forceinline__ void store_pair(float* row, int pr, float2 v, bool both) {
  if (both) reinterpret_cast<float2*>(row)[pr] = v;
  else row[2 * pr] = v.x;
}

// one atomic per warp per cause: the thread's two aircraft contribute p0 and p1
__device__ __forceinline__ void count_cause2(unsigned long long* counters, int which, bool p0, bool p1) {
  const int k = __popc(__ballot_sync(0xffffffffu, p0)) + __popc(__ballot_sync(0xffffffffu, p1));
  if (k != 0 && (threadIdx.x & 31) == 0) atomicAdd(&counters[which], (unsigned long long)k);
}

// F16Model.update's control lag (F16_model.py:52-57): clamp, then first-order low-pass towards the scaled action
__device__ __forceinline__ void lowpass_controls(const float* a_in, float* u) {
  float a[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = fminf(fmaxf(a_in[j], -1.0f), 1.0f);
  u[0] = 0.9f * u[0] + 0.1f * a[0] * 0.225f * 76300.0f / DC(0.3048f);
  u[1] = 0.9f * u[1] + 0.1f * a[1] * 45.0f;
  u[2] = 0.9f * u[2] + 0.1f * a[2] * 45.0f;
  u[3] = 0.9f * u[3] + 0.1f * a[3] * 45.0f;
}

// The six termination predicates (task_base.py:75-96) and the task reward (task_base.py:60-73) of ONE aircraft at its new
// state; f = the force part of nlplant at that state (Overload needs the body accelerations, F16_model.py:132-148).
// causes: bit 0 overload, 1 low altitude, 2 high speed, 3 low speed, 4 extreme state, 5 unreach, 6 reached.
struct Verdict {
  bool bad, done, exc;
  float rw;
  int causes;
};
template <bool COMBAT, int TASK>
__device__ __forceinline__ Verdict judge_state(const np_env_cfg& c, const float* sq, const float* tq, const Trig& g, const ForceOut& f,
                                               int steps) {
  float ax, ay, az;
  body_accel(sq, g, f, ax, ay, az);
  const float acc = sqrtf(ax * ax + ay * ay + az * az);
  const bool overload = (acc - c.acceleration_limit) > 0.0f;            // overload.py:37-42
  const bool low_alt = (sq[2] - c.altitude_limit) < 0.0f;               // low_altitude.py:29-30
  const float vel = (sq[6] + c.airspeed * 1.0f) * 0.3048f / DC(340.0f);
  const bool hi = (vel - c.max_velocity) >= 0.0f;                       // high_speed.py:29-30
  const bool lo = (vel - c.min_velocity) <= 0.0f;                       // low_speed.py:29-30
  const float a_deg = sq[7] * 180.0f / DC(kPi), b_deg = sq[8] * 180.0f / DC(kPi);  // extreme_state.py:32-36
  const bool ext = (a_deg < c.min_alpha) | (a_deg > c.max_alpha) | (b_deg < c.min_beta) | (b_deg > c.max_beta);
  const bool late = steps >= c.max_check_interval;
  bool off = false, dn = false, exc = false;
  float d0, d1, d2, rw = 0.0f;
  if (COMBAT) {
    exc = (steps - c.max_steps) >= 0;                                   // timeout.py:29
  } else if (TASK == NP_TASK_HEADING) {                                 // unreach_heading.py:38-53
    const float dpsi = wrap_pi(sq[5] - tq[1]);
    off = (fabsf(dpsi) >= (float)(3.141592653589793 / 36.0)) | (fabsf(sq[2] - tq[0]) >= 100.0f) |
          (fabsf(sq[6] - tq[2]) >= 20.0f);
    dn = !off && !late && (steps >= c.min_check_interval);
    d0 = (sq[2] - tq[0]) * 0.3048f / DC(1000.0f);                       // heading_reward.py:26-35
    d1 = dpsi / DC(kPi);
    d2 = (sq[6] - tq[2]) * 0.3048f / DC(340.0f);
    rw = -(d0 * d0) + -(d1 * d1) + -(d2 * d2);
  } else if (TASK == NP_TASK_CONTROL) {                                 // unreach_posture.py:37-55
    const float dpsi = wrap_pi(sq[5] - tq[1]);
    off = (fabsf(dpsi) >= (float)(3.141592653589793 / 36.0)) |
          (fabsf(sq[4] - tq[0]) >= (float)(3.141592653589793 / 36.0)) | (fabsf(sq[6] - tq[2]) >= 20.0f);
    dn = !off && !late;
    d0 = wrap_pi(sq[4] - tq[0]) / DC(kPi);                              // posture_reward.py:26-34
    d1 = dpsi / DC(kPi);
    d2 = (sq[6] - tq[2]) * 0.3048f / DC(340.0f);
    rw = -(d0 * d0) + -(d1 * d1) + -(d2 * d2);
  } else {                                                              // unreach_target.py:35-47
    off = (fabsf(sq[0] - tq[0]) >= 100.0f) | (fabsf(sq[1] - tq[1]) >= 100.0f) | (fabsf(sq[2] - tq[2]) >= 100.0f);
    dn = !off && !late;
    d0 = (sq[0] - tq[0]) * 0.3048f / DC(1000.0f);                       // position_reward.py:26-34
    d1 = (sq[1] - tq[1]) * 0.3048f / DC(1000.0f);
    d2 = (sq[2] - tq[2]) * 0.3048f / DC(1000.0f);
    rw = 0.1f * (-(d0 * d0) + -(d1 * d1) + -(d2 * d2));
  }
  const bool unreach = late && off;
  Verdict v;
  v.bad = overload | low_alt | hi | lo | ext | unreach;
  v.done = dn;
  v.exc = exc;
  v.rw = rw;
  v.causes = (int)overload | ((int)low_alt << 1) | ((int)hi << 2) | ((int)lo << 3) | ((int)ext << 4) | ((int)unreach << 5) | ((int)dn << 6);
  return v;
}

