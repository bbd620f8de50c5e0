// K2: the UAV plug-in's env kernels (per-thread and TMA-staged slab pipeline) and its stand-alone nlplant kernel.
#pragma once
#include "env_device.cuh"
#include "ptx_device.cuh"
#include "uav_device.cuh"

// ------------------------------------------------------------------------------------------------
// K2: the UAV plug-in (envs/models/UAV_model.py, UAV/UAV_dynamics.py) behind the same tasks.  Trivial arithmetic,
// HBM-bound: one aircraft per thread, coalesced SoA rows, one launch per BaseEnv.step().
// Algorithmic bytes per aircraft-step: 268 = read 96 (s 48, F 12, tgt 12, step 4, flags 4, action 16) +
// write 172 (s 48, F 12, tgt 12, step 4, flags 4, obs 88, reward 4).
// ------------------------------------------------------------------------------------------------
// task.reset on the getter view (heading_task.py:49-69, control_task.py:49-68, tracking_task.py:48-71)
__device__ __forceinline__ void uav_task_reset(const np_env_cfg& c, const UavView& v, const Draws& r, float* tgt) {
  if (c.task == NP_TASK_HEADING) {
    tgt[0] = v.alt + 1000.0f;
    tgt[1] = wrap_pi(v.heading + (float)(2.0 * 3.141592653589793 / 3.0));
    tgt[2] = v.vt + 0.0f;
  } else if (c.task == NP_TASK_CONTROL) {
    tgt[0] = wrap_pi(v.pitch + 2.0f * (r.d[2] - 0.5f) * c.max_pitch_increment);
    tgt[1] = wrap_pi(v.heading + 2.0f * (r.d[3] - 0.5f) * c.max_heading_increment);
    tgt[2] = v.vt + 2.0f * (r.d[4] - 0.5f) * c.max_velocities_u_increment;
  } else {
    const float dist = r.d[2] * (c.max_distance - c.min_distance) + c.min_distance;
    const float th1 = r.d[3] * kPi / DC(3.0f) - (float)(3.141592653589793 / 6.0);
    const float th2 = r.d[4] * kPi / DC(3.0f) - (float)(3.141592653589793 / 6.0);
    tgt[0] = v.npos + dist * cosf(th1) * cosf(th2);
    tgt[1] = v.epos + dist * cosf(th1) * sinf(th2);
    tgt[2] = v.alt + dist * sinf(th1);
  }
}

// UAVModel.reset (UAV_model.py:32-45): SI state, zero forces except u[0] = init_T
__device__ __forceinline__ void uav_reset_aircraft(const np_env_cfg& c, const Draws& r, float* s, float* F) {
#pragma unroll
  for (int j = 0; j < 12; ++j) s[j] = 0.0f;
  s[2] = (r.d[0] * (c.max_altitude - c.min_altitude) + c.min_altitude) * 0.3048f;
  s[6] = (r.d[1] * (c.max_vt - c.min_vt) + c.min_vt) * 0.3048f;
  F[0] = c.init_T; F[1] = 0.0f; F[2] = 0.0f;
}

// 22-D observation through the getters (heading_task.py:93-152): AOA = AOS = thrust = surfaces = 0 for this model
__device__ __forceinline__ void uav_make_obs(const np_env_cfg& c, const float* s, const UavView& v, const UavTrig& t, const float* tgt,
                                             float* o) {
  if (c.task == NP_TASK_HEADING) {
    o[0] = (v.alt - tgt[0]) * 0.3048f / DC(1000.0f);
    o[1] = wrap_pi(v.heading - tgt[1]);
    o[2] = (v.vt - tgt[2]) * 0.3048f / DC(340.0f);
  } else if (c.task == NP_TASK_CONTROL) {
    o[0] = wrap_pi(v.pitch - tgt[0]);
    o[1] = wrap_pi(v.heading - tgt[1]);
    o[2] = (v.vt - tgt[2]) * 0.3048f / DC(340.0f);
  } else {
    o[0] = (v.npos - tgt[0]) * 0.3048f / DC(1000.0f);
    o[1] = (v.epos - tgt[1]) * 0.3048f / DC(1000.0f);
    o[2] = (v.alt - tgt[2]) * 0.3048f / DC(1000.0f);
  }
  const float eas = (v.vt + c.airspeed * 1.0f) / v.e2t;  // UAV_model.py:94-102
  o[3] = v.alt * 0.3048f / DC(5000.0f);
  o[4] = t.sphi; o[5] = t.cphi; o[6] = t.st; o[7] = t.ct;
  o[8] = eas * 0.3048f / DC(340.0f);
  o[9] = 0.0f; o[10] = 1.0f; o[11] = 0.0f; o[12] = 1.0f;  // sin / cos of get_AOA() = get_AOS() = 0
  o[13] = s[9]; o[14] = s[10]; o[15] = s[11];
  o[16] = 0.0f / DC(0.225f) / DC(76300.0f) * 0.3048f;             // get_thrust() = 0
  o[17] = 0.0f / DC(45.0f); o[18] = 0.0f / DC(45.0f); o[19] = 0.0f / DC(45.0f); o[20] = 0.0f / DC(45.0f);
  o[21] = v.e2t;
}

// One aircraft of BaseEnv.step / reset for the UAV plug-in, entirely in registers: masked reset -> (STEP) force
// low-pass + Euler step -> observation row -> (STEP) terminations + reward.  Shared by the per-thread kernel and the
// TMA-staged slab kernel below, so both produce identical bits.
template <bool STEP>
__device__ __forceinline__ void uav_aircraft(const StepParams& p, uint32_t rng, int i, bool rst, const float4 av, float* s, float* F, float* tgt,
                                             int& steps, float* o, float& rew, bool& done, bool& bad) {
  const np_env_cfg& c = p.cfg;
  // ---- BaseEnv.reset (env_base.py:83-97) ---------------------------------------------------------
  if (rst) {
    const Draws r = reset_draws(p, i, rng);
    uav_reset_aircraft(c, r, s, F);
    uav_task_reset(c, uav_view(s), r, tgt);
    steps = 0;
    atomicAdd(&p.counters[7], 1ull);
  }
  bad = false; done = false; rew = 0.0f;
  if (STEP) {
    // ---- UAVModel.update (UAV_model.py:51-62): clamp, force low-pass, one explicit Euler step ------------
    const float a[3] = {fminf(fmaxf(av.x, -1.0f), 1.0f), fminf(fmaxf(av.y, -1.0f), 1.0f), fminf(fmaxf(av.z, -1.0f), 1.0f)};
#pragma unroll
    for (int j = 0; j < 3; ++j) F[j] = 0.9f * F[j] + 0.1f * a[j] * 27000.0f;
    float xdot[12];
    uav_nlplant(s, F, xdot);
    const float h = c.dt - 0.0f;
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = s[j] + h * xdot[j];
    steps += 1;
  }
  // ---- obs (env_base.py:103) ----------------------------------------------------------------------
  const UavView v = uav_view(s);
  const UavTrig trig = uav_trig(s);   // of the state the observation, the Overload check and the reward all see
  uav_make_obs(c, s, v, trig, tgt, o);
  add_obs_noise(p, i, o, rng);
  if (STEP) {
    // ---- terminations (task_base.py:75-96) through the getters ------------------------------------------
    float xdot[12];
    uav_nlplant(s, F, trig, xdot);                                        // get_acceleration (UAV_model.py:120-130)
    const float vu = s[6] / DC(0.3048f), vv = s[7] / DC(0.3048f), vw = s[8] / DC(0.3048f);
    const float ax = xdot[6] / DC(0.3048f) + s[10] * vw - s[11] * vv;
    const float ay = xdot[7] / DC(0.3048f) + s[11] * vu - s[9] * vw;
    const float az = xdot[8] / DC(0.3048f) + s[9] * vv - s[10] * vu;
    const float acc = sqrtf(ax * ax + ay * ay + az * az);
    const bool overload = (acc - c.acceleration_limit) > 0.0f;
    const bool low_alt = (v.alt - c.altitude_limit) < 0.0f;
    const float vel = (v.vt + c.airspeed * 1.0f) * 0.3048f / DC(340.0f);
    const bool hi = (vel - c.max_velocity) >= 0.0f;
    const bool lo = (vel - c.min_velocity) <= 0.0f;
    const float a_deg = 0.0f * 180.0f / DC(kPi);                              // get_AOA() = get_AOS() = 0
    const bool ext = (a_deg < c.min_alpha) | (a_deg > c.max_alpha) | (a_deg < c.min_beta) | (a_deg > c.max_beta);
    const bool late = steps >= c.max_check_interval;
    bool off;
    float d0, d1, d2;
    if (c.task == NP_TASK_HEADING) {
      const float dpsi = wrap_pi(v.heading - tgt[1]);
      off = (fabsf(dpsi) >= (float)(3.141592653589793 / 36.0)) | (fabsf(v.alt - tgt[0]) >= 100.0f) |
            (fabsf(v.vt - tgt[2]) >= 20.0f);
      done = !off && !late && (steps >= c.min_check_interval);
      d0 = (v.alt - tgt[0]) * 0.3048f / DC(1000.0f); d1 = dpsi / DC(kPi); d2 = (v.vt - tgt[2]) * 0.3048f / DC(340.0f);
      rew = -(d0 * d0) + -(d1 * d1) + -(d2 * d2);
    } else if (c.task == NP_TASK_CONTROL) {
      const float dpsi = wrap_pi(v.heading - tgt[1]);
      off = (fabsf(dpsi) >= (float)(3.141592653589793 / 36.0)) | (fabsf(v.pitch - tgt[0]) >= (float)(3.141592653589793 / 36.0)) |
            (fabsf(v.vt - tgt[2]) >= 20.0f);
      done = !off && !late;
      d0 = wrap_pi(v.pitch - tgt[0]) / DC(kPi); d1 = dpsi / DC(kPi); d2 = (v.vt - tgt[2]) * 0.3048f / DC(340.0f);
      rew = -(d0 * d0) + -(d1 * d1) + -(d2 * d2);
    } else {
      off = (fabsf(v.npos - tgt[0]) >= 100.0f) | (fabsf(v.epos - tgt[1]) >= 100.0f) | (fabsf(v.alt - tgt[2]) >= 100.0f);
      done = !off && !late;
      d0 = (v.npos - tgt[0]) * 0.3048f / DC(1000.0f); d1 = (v.epos - tgt[1]) * 0.3048f / DC(1000.0f);
      d2 = (v.alt - tgt[2]) * 0.3048f / DC(1000.0f);
      rew = 0.1f * (-(d0 * d0) + -(d1 * d1) + -(d2 * d2));
    }
    const bool unreach = late && off;
    bad = overload | low_alt | hi | lo | ext | unreach;
    rew = rew + (float)(-200 * (int)bad + 200 * (int)done);
    const bool cause[7] = {overload, low_alt, hi, lo, ext, unreach, done};
#pragma unroll
    for (int w = 0; w < 7; ++w)
      if (cause[w]) atomicAdd(&p.counters[w], 1ull);   // ptxas aggregates warp-uniform-address atomics (REDUX + one ATOM)
  }
}

// Per-thread variant (reset, unaligned ranges): scalar SoA accesses straight to global memory.
template <bool STEP>
__global__ void __launch_bounds__(256, 4) uav_env_kernel(const __grid_constant__ StepParams p) {
  const np_env_cfg& c = p.cfg;
  const int n = c.n, ld = c.ld;
  const int i_end = min(n, 2 * p.pair_end);
  const uint32_t rng = rng_step(p);   // the RNG counter of this launch: read once per thread, not once per aircraft
  for (int i = 2 * p.pair_begin + blockIdx.x * blockDim.x + threadIdx.x; i < i_end; i += gridDim.x * blockDim.x) {
    float s[12], F[3], tgt[3], o[NP_NUM_OBS], rew;
    bool done, bad;
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = p.s[(size_t)j * ld + i];
#pragma unroll
    for (int j = 0; j < 3; ++j) F[j] = p.u[(size_t)j * ld + i];
#pragma unroll
    for (int j = 0; j < 3; ++j) tgt[j] = p.tgt[(size_t)j * ld + i];
    int steps = p.step_count[i];
    const bool rst = (p.flags[i] | p.flags[ld + i] | p.flags[2 * (size_t)ld + i]) != 0;
    const float4 av = STEP ? reinterpret_cast<const float4*>(p.action)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    uav_aircraft<STEP>(p, rng, i, rst, av, s, F, tgt, steps, o, rew, done, bad);
    float2* orow = reinterpret_cast<float2*>(p.obs + (size_t)i * NP_NUM_OBS);
#pragma unroll
    for (int j = 0; j < NP_NUM_OBS / 2; ++j) orow[j] = make_float2(o[2 * j], o[2 * j + 1]);
    if (STEP) p.reward[i] = rew;
#pragma unroll
    for (int j = 0; j < 12; ++j) p.s[(size_t)j * ld + i] = s[j];
#pragma unroll
    for (int j = 0; j < 3; ++j) p.u[(size_t)j * ld + i] = F[j];
#pragma unroll
    for (int j = 0; j < 3; ++j) p.tgt[(size_t)j * ld + i] = tgt[j];
    p.step_count[i] = steps;
    p.flags[i] = done ? 1 : 0;
    p.flags[ld + i] = bad ? 1 : 0;
    p.flags[2 * (size_t)ld + i] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// K2 (TMA-staged): the HBM-bound UAV step as a persistent slab pipeline.  A CTA owns 256-aircraft slabs.
//   in : every SoA row segment (1 KB), the action block (4 KB) and the flag rows arrive in shared memory through TMA bulk
//        copies completing on an mbarrier (23 copies, one per lane of warp 0); slab k+1's inputs are requested as soon as
//        slab k's are in registers and land during slab k's arithmetic;
//   out: the strided part -- the 256 x 22 observation block (22.5 KB, contiguous in the row-major obs array) -- is staged
//        and leaves as ONE bulk store that drains during slab k+1's arithmetic; SoA rows, reward and flags are already
//        coalesced (128 B per warp and row) and go straight from registers.
// The per-thread variant is LSU-queue / latency limited (23 scalar loads + 34 stores per aircraft, the 11 observation
// stores touching 32 separate sectors each).  46.9 KB and 64 registers per thread -> 4 CTAs (32 warps) per SM.
// The ragged tail slab (< 256 aircraft) goes through guarded per-thread accesses in the same kernel.
// ------------------------------------------------------------------------------------------------
namespace uavslab {
constexpr int kSlab = 256;
constexpr int IN_S = 0;                            // [12][256] f32
constexpr int IN_U = IN_S + 12 * kSlab * 4;        // [3][256] f32
constexpr int IN_T = IN_U + 3 * kSlab * 4;         // [3][256] f32
constexpr int IN_STEP = IN_T + 3 * kSlab * 4;      // [256] i32
constexpr int IN_ACT = IN_STEP + kSlab * 4;        // [256][4] f32
constexpr int IN_FLG = IN_ACT + kSlab * 16;        // [3][256] u8
constexpr int IN_BYTES = IN_FLG + 3 * kSlab;       // 24 320
constexpr int OUT_OBS = IN_BYTES;                  // [256][22] f32: the only output staged in shared memory
constexpr int BAR = OUT_OBS + kSlab * NP_NUM_OBS * 4;
constexpr int SMEM_BYTES = BAR + 32;               // 46 880 (four mbarriers) -> 4 CTAs = 32 warps per SM
static_assert(OUT_OBS % 16 == 0 && BAR % 8 == 0, "bulk copies need 16-byte aligned shared addresses");
}  // namespace uavslab


// warp 0 requests the inputs of the full slab starting at aircraft i0: lane r fetches row r
__device__ __forceinline__ void uav_slab_request(const StepParams& p, unsigned char* sm, uint64_t* bar, int i0, int lane) {
  using namespace uavslab;
  const size_t ld = (size_t)p.cfg.ld;
  if (lane == 0) mbar_expect_tx(bar, IN_BYTES);
  __syncwarp();
  if (lane < 12) bulk_g2s(sm + IN_S + lane * kSlab * 4, p.s + lane * ld + i0, kSlab * 4, bar);
  else if (lane < 15) bulk_g2s(sm + IN_U + (lane - 12) * kSlab * 4, p.u + (lane - 12) * ld + i0, kSlab * 4, bar);
  else if (lane < 18) bulk_g2s(sm + IN_T + (lane - 15) * kSlab * 4, p.tgt + (lane - 15) * ld + i0, kSlab * 4, bar);
  else if (lane == 18) bulk_g2s(sm + IN_STEP, p.step_count + i0, kSlab * 4, bar);
  else if (lane == 19) bulk_g2s(sm + IN_ACT, p.action + (size_t)i0 * 4, kSlab * 16, bar);
  else if (lane < 23) bulk_g2s(sm + IN_FLG + (lane - 20) * kSlab, p.flags + (lane - 20) * ld + i0, kSlab, bar);
}

// one lane of warp 0 sends the slab's 256 x 22 observation block (contiguous in the row-major obs array) as one bulk store
__device__ __forceinline__ void uav_slab_send(const StepParams& p, unsigned char* sm, int i0, int lane) {
  using namespace uavslab;
  if (lane == 0) {
    bulk_s2g(p.obs + (size_t)i0 * NP_NUM_OBS, sm + OUT_OBS, kSlab * NP_NUM_OBS * 4);
    bulk_commit();
  }
}

__global__ void __launch_bounds__(uavslab::kSlab, 4) uav_step_slab_kernel(const __grid_constant__ StepParams p) {
  using namespace uavslab;
  extern __shared__ __align__(128) unsigned char sm[];
  const int ld = p.cfg.ld, t = threadIdx.x, lane = t & 31;
  const bool warp0 = t < 32;
  const int i_begin = 2 * p.pair_begin, i_end = min(p.cfg.n, 2 * p.pair_end);
  const int nslab = (i_end - i_begin + kSlab - 1) / kSlab;
  // No CTA-wide barrier in the slab loop: the eight warps are coupled only through four mbarriers, so a warp that is
  // ahead keeps issuing (a __syncthreads version measured 3.2 barrier-stall cycles per issued instruction).
  uint64_t* in_full = reinterpret_cast<uint64_t*>(sm + BAR);  // TMA: the slab's inputs have landed            (tx bytes)
  uint64_t* in_read = in_full + 1;                            // every warp holds its inputs in registers        (8 warps)
  uint64_t* written = in_full + 2;                            // every warp has staged its observation rows      (8 warps)
  uint64_t* out_free = in_full + 3;                           // the previous bulk store has read the obs block      (1)
  if (t == 0) {
    mbar_init(in_full, 1);
    mbar_init(in_read, kSlab / 32);
    mbar_init(written, kSlab / 32);
    mbar_init(out_free, 1);
  }
  __syncthreads();

  const uint32_t rng = rng_step(p);   // the RNG counter of this launch: read once per thread, not once per aircraft
  int slab = blockIdx.x;
  if (warp0 && slab < nslab && i_begin + (slab + 1) * kSlab <= i_end) uav_slab_request(p, sm, in_full, i_begin + slab * kSlab, lane);

  for (int it = 0; slab < nslab; slab += gridDim.x, ++it) {
    const int i0 = i_begin + slab * kSlab, i = i0 + t;
    const bool full = i0 + kSlab <= i_end;       // CTA-uniform; only the last slab of the range can be ragged
    const bool live = full || i < i_end;
    const uint32_t ph = it & 1;
    float s[12], F[3], tgt[3], o[NP_NUM_OBS], rew = 0.0f;
    bool done = false, bad = false, rst = false;
    int steps = 0;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (full) {
      mbar_wait(in_full, ph);
#pragma unroll
      for (int j = 0; j < 12; ++j) s[j] = reinterpret_cast<const float*>(sm + IN_S)[j * kSlab + t];
#pragma unroll
      for (int j = 0; j < 3; ++j) F[j] = reinterpret_cast<const float*>(sm + IN_U)[j * kSlab + t];
#pragma unroll
      for (int j = 0; j < 3; ++j) tgt[j] = reinterpret_cast<const float*>(sm + IN_T)[j * kSlab + t];
      steps = reinterpret_cast<const int*>(sm + IN_STEP)[t];
      av = reinterpret_cast<const float4*>(sm + IN_ACT)[t];
      rst = (sm[IN_FLG + t] | sm[IN_FLG + kSlab + t] | sm[IN_FLG + 2 * kSlab + t]) != 0;
      __syncwarp();
      if (lane == 0) mbar_arrive(in_read);
      if (warp0) {
        if (it > 0) {            // the previous slab's stores were issued a whole input wait ago: normally drained by now
          bulk_wait_read0();
          __syncwarp();
          if (lane == 0) mbar_arrive(out_free);
        }
        const int next = slab + gridDim.x;
        if (next < nslab && i_begin + (next + 1) * kSlab <= i_end) {
          mbar_wait(in_read, ph);   // the input slab is free again: its next contents land during this slab's arithmetic
          uav_slab_request(p, sm, in_full, i_begin + next * kSlab, lane);
        }
      }
    } else if (live) {
#pragma unroll
      for (int j = 0; j < 12; ++j) s[j] = p.s[(size_t)j * ld + i];
#pragma unroll
      for (int j = 0; j < 3; ++j) F[j] = p.u[(size_t)j * ld + i];
#pragma unroll
      for (int j = 0; j < 3; ++j) tgt[j] = p.tgt[(size_t)j * ld + i];
      steps = p.step_count[i];
      av = reinterpret_cast<const float4*>(p.action)[i];
      rst = (p.flags[i] | p.flags[ld + i] | p.flags[2 * (size_t)ld + i]) != 0;
    }

    if (live) uav_aircraft<true>(p, rng, i, rst, av, s, F, tgt, steps, o, rew, done, bad);

    if (full) {
      // SoA rows, reward and flags: fully coalesced 4-byte stores straight from registers (128 B per warp and row)
#pragma unroll
      for (int j = 0; j < 12; ++j) p.s[(size_t)j * ld + i] = s[j];
#pragma unroll
      for (int j = 0; j < 3; ++j) p.u[(size_t)j * ld + i] = F[j];
#pragma unroll
      for (int j = 0; j < 3; ++j) p.tgt[(size_t)j * ld + i] = tgt[j];
      p.step_count[i] = steps;
      p.reward[i] = rew;
      p.flags[i] = done ? 1 : 0;
      p.flags[ld + i] = bad ? 1 : 0;
      p.flags[2 * (size_t)ld + i] = 0;
      // the 88-byte observation rows are the strided part: staged, then one 22.5 KB bulk store per slab
      if (it > 0) mbar_wait(out_free, ph ^ 1);
      float2* orow = reinterpret_cast<float2*>(sm + OUT_OBS) + t * (NP_NUM_OBS / 2);  // 8-byte stride 11: conflict-free
#pragma unroll
      for (int j = 0; j < NP_NUM_OBS / 2; ++j) orow[j] = make_float2(o[2 * j], o[2 * j + 1]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(written);
      if (warp0) {
        mbar_wait(written, ph);
        uav_slab_send(p, sm, i0, lane);
      }
    } else if (live) {
      float2* orow = reinterpret_cast<float2*>(p.obs + (size_t)i * NP_NUM_OBS);
#pragma unroll
      for (int j = 0; j < NP_NUM_OBS / 2; ++j) orow[j] = make_float2(o[2 * j], o[2 * j + 1]);
      p.reward[i] = rew;
#pragma unroll
      for (int j = 0; j < 12; ++j) p.s[(size_t)j * ld + i] = s[j];
#pragma unroll
      for (int j = 0; j < 3; ++j) p.u[(size_t)j * ld + i] = F[j];
#pragma unroll
      for (int j = 0; j < 3; ++j) p.tgt[(size_t)j * ld + i] = tgt[j];
      p.step_count[i] = steps;
      p.flags[i] = done ? 1 : 0;
      p.flags[ld + i] = bad ? 1 : 0;
      p.flags[2 * (size_t)ld + i] = 0;
    }
  }
  if (warp0) bulk_wait0();  // shared memory must outlive the last bulk stores
}

__global__ void __launch_bounds__(256) uav_nlplant_kernel(const float* __restrict__ S, const float* __restrict__ U,
                                                          float* __restrict__ X, int n, int ld) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s[12], F[3], xdot[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = S[(size_t)j * ld + i];
#pragma unroll
    for (int j = 0; j < 3; ++j) F[j] = U[(size_t)j * ld + i];
    uav_nlplant(s, F, xdot);
#pragma unroll
    for (int j = 0; j < 12; ++j) X[(size_t)j * ld + i] = xdot[j];
  }
}

