// Combat device code (K5 / K5r): env-level re-initialisation, AO / TA / R geometry, orientation / range reward and blood
// functions, the per-aircraft combat record and the pair terms built from two records.  Reference: envs/singlecombat_env.py,
// envs/multiplecombat_env.py, envs/utils/utils.py:156-249, termination_conditions/{crash,shutdown,timeout}.py.
#pragma once
#include "env_device.cuh"

// SingleCombatEnv.reset_done_envs re-initialisation of one aircraft (:219-225): draws npos, epos, altitude, heading, vt
__device__ __forceinline__ void combat_reset_aircraft(const np_env_cfg& c, const Draws& r, float* s, float* u) {
#pragma unroll
  for (int j = 0; j < 12; ++j) s[j] = 0.0f;
  s[0] = r.d[0] * (c.max_npos - c.min_npos) + c.min_npos;
  s[1] = r.d[1] * (c.max_epos - c.min_epos) + c.min_epos;
  s[2] = r.d[2] * (c.max_altitude - c.min_altitude) + c.min_altitude;
  s[5] = r.d[3] * (c.max_heading - c.min_heading) + c.min_heading;
  s[6] = r.d[4] * (c.max_vt - c.min_vt) + c.min_vt;
  u[0] = c.init_T; u[1] = 0.0f; u[2] = 0.0f; u[3] = 0.0f;
}

// AO / TA / R of get_AO_TA_R (3-D) or get2d_AO_TA_R (utils.py:156-206): dp = enemy - ego position, ve / vm = ego / enemy
// inertial velocity (xdot[0:3]); DIM = 3 or 2.
template <int DIM>
__device__ __forceinline__ void ao_ta_r(const float* dp, const float* ve, const float* vm, float& AO, float& TA, float& R) {
  float ev = 0.f, mv = 0.f, d2 = 0.f, pe = 0.f, pm = 0.f;
#pragma unroll
  for (int j = 0; j < DIM; ++j) {
    ev = ev + ve[j] * ve[j]; mv = mv + vm[j] * vm[j]; d2 = d2 + dp[j] * dp[j];
    pe = pe + dp[j] * ve[j]; pm = pm + dp[j] * vm[j];
  }
  R = sqrtf(d2);
  AO = acosf(fminf(fmaxf(pe / (R * sqrtf(ev) + 1e-8f), -1.0f), 1.0f));
  TA = acosf(fminf(fmaxf(pm / (R * sqrtf(mv) + 1e-8f), -1.0f), 1.0f));
}
__device__ __forceinline__ float orientation_reward_v2(float AO, float TA) {  // utils.py:215-217
  const float t = atanhf(1.0f - fmaxf(1.9f * TA / DC(kPi), 1e-4f * 1.0f)) / (2.0f * kPi);
  return 1.0f / (50.0f * AO / DC(kPi) + 2.0f) + (float)(1.0 / 2) + fminf(t, 0.0f) + 0.5f;
}
__device__ __forceinline__ float range_reward_v3(float Rkm) {  // utils.py:230-231
  const float poly = fminf(fmaxf(-0.032f * (Rkm * Rkm) + 0.284f * Rkm + 0.38f, 0.0f), 1.0f);
  return (Rkm < 5.0f ? 1.0f : 0.0f) + (Rkm >= 5.0f ? 1.0f : 0.0f) * poly + fminf(fmaxf(expf(-0.16f * Rkm), 0.0f), 0.2f);
}
__device__ __forceinline__ float orientation_fn(float AO) {  // utils.py:235-243
  constexpr float k6 = (float)(3.141592653589793 / 6);
  const bool m3 = (AO >= 0.0f) & (AO <= k6), m4 = (AO <= 0.0f) & (AO >= -k6);
  return (1.0f - 6.0f * AO / DC(kPi)) * (m3 ? 1.0f : 0.0f) + (1.0f + 6.0f * AO / DC(kPi)) * (m4 ? 1.0f : 0.0f);
}
__device__ __forceinline__ float distance_fn(float Rkm) {  // utils.py:245-249
  return (Rkm <= 1.0f ? 1.0f : 0.0f) + (3.0f - Rkm) / 2.0f * (((Rkm > 1.0f) & (Rkm <= 3.0f)) ? 1.0f : 0.0f);
}

// What the pairwise terms need of ONE aircraft at its final state: position, inertial velocity es = xdot[0:3] of nlplant
// (F16_dynamics.py:104,129-135), body-axis velocity (F16Model.get_velocity), roll / pitch trigonometry and vt.  In the
// pair-sharded layout both records are built in the thread that owns the pair; in the role-sharded layout each rank builds
// its own and pulls the partner's from the peer's record slab -- the same code either way, so the outputs agree bit for bit.
struct CombatRec {
  float pos[3], es[3], vel[3], sphi, cphi, st, ct, vt;
};
__device__ __forceinline__ CombatRec combat_rec(const float* s) {
  CombatRec r;
  const Trig t = make_trig(s);
  const float vt = s[6];
  r.pos[0] = s[0]; r.pos[1] = s[1]; r.pos[2] = s[2];
  r.vel[0] = vt * t.cb * t.ca;
  r.vel[1] = vt * t.sb;
  r.vel[2] = vt * t.cb * t.sa;
  const float vtc = vt <= 0.01f ? 0.01f : vt;
  const BodyVel b = body_vel(vtc, t);
  r.es[0] = b.U * (t.ct * t.cpsi) + b.V * (t.sphi * t.cpsi * t.st - t.cphi * t.spsi) + b.W * (t.cphi * t.st * t.cpsi + t.sphi * t.spsi);
  r.es[1] = b.U * (t.ct * t.spsi) + b.V * (t.sphi * t.spsi * t.st + t.cphi * t.cpsi) + b.W * (t.cphi * t.st * t.spsi - t.sphi * t.cpsi);
  r.es[2] = b.U * t.st - b.V * (t.sphi * t.ct) - b.W * (t.cphi * t.ct);
  r.sphi = t.sphi; r.cphi = t.cphi; r.st = t.st; r.ct = t.ct; r.vt = vt;
  return r;
}
// pairwise geometry of (ego, enemy): get_AO_TA_R / get2d_AO_TA_R + the side flag (singlecombat_env.py:96-121,142-150)
struct CombatGeo {
  float AO, TA, R, AO2, TA2, R2, side, Rkm;
};
__device__ __forceinline__ CombatGeo combat_geo(const CombatRec& e, const CombatRec& m) {
  CombatGeo g;
  const float dp[3] = {m.pos[0] - e.pos[0], m.pos[1] - e.pos[1], m.pos[2] - e.pos[2]};
  ao_ta_r<2>(dp, e.es, m.es, g.AO2, g.TA2, g.R2);
  ao_ta_r<3>(dp, e.es, m.es, g.AO, g.TA, g.R);
  const float cz = e.es[0] * dp[1] - e.es[1] * dp[0];
  g.side = (cz > 0.0f ? 1.0f : 0.0f) - (cz < 0.0f ? 1.0f : 0.0f);
  g.Rkm = g.R * 0.3048f / DC(1000.0f);
  return g;
}
// 15-D observation row of one aircraft (`own`) given its partner and the pair geometry; q = 0: ego, 1: enemy (mirrored)
__device__ __forceinline__ void combat_obs_row(const CombatRec& own, const CombatRec& other, const CombatGeo& g, int q, float* o) {
  o[0] = own.pos[2] * 0.3048f / DC(5000.0f);
  o[1] = own.sphi; o[2] = own.cphi; o[3] = own.st; o[4] = own.ct;
  o[5] = own.vel[0] * 0.3048f / DC(340.0f); o[6] = own.vel[1] * 0.3048f / DC(340.0f); o[7] = own.vel[2] * 0.3048f / DC(340.0f);
  o[8] = own.vt * 0.3048f / DC(340.0f);
  o[9] = (other.vel[0] - own.vel[0]) * 0.3048f / DC(340.0f);
  o[10] = (other.pos[2] - own.pos[2]) * 0.3048f / DC(1000.0f);
  o[11] = q == 0 ? g.AO2 : kPi - g.TA2;
  o[12] = q == 0 ? g.TA2 : kPi - g.AO2;
  o[13] = g.R2 * 0.3048f / 10000.0f;
  o[14] = q == 0 ? g.side : -g.side;
}
// singlecombat_env.py:140-181 (scale 0.01) / multiplecombat_env.py:163-181 (scale 1: the product itself)
__device__ __forceinline__ float combat_reward(const CombatGeo& g, int q, float scale) {
  const float rr = range_reward_v3(g.Rkm);
  return q == 0 ? scale * (orientation_reward_v2(g.AO, g.TA) * rr) : scale * (orientation_reward_v2(kPi - g.TA, kPi - g.AO) * rr);
}
// blood model (singlecombat_env.py:263-271): what aircraft q loses in this env step
__device__ __forceinline__ float combat_damage(const CombatGeo& g, int q) {
  const float df = distance_fn(g.Rkm);
  return q == 0 ? orientation_fn(kPi - g.TA) * df : orientation_fn(g.AO) * df;
}

// obs (singlecombat_env.py:64-138), reward (:140-181) and the blood model (:263-271) of one pair at its final state.
__device__ __forceinline__ void combat_outputs(const StepParams& p, float (&s)[2][12], float (&blood)[2], float (&rew)[2],
                                               int pr, const bool (&act)[2], bool stepped) {
  const CombatRec rec[2] = {combat_rec(s[0]), combat_rec(s[1])};
  const CombatGeo g = combat_geo(rec[0], rec[1]);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    rew[q] = combat_reward(g, q, p.cfg.combat_reward_scale);
    float o[NP_NUM_OBS_COMBAT];
    combat_obs_row(rec[q], rec[1 - q], g, q, o);
    if (act[q]) {
      float* orow = p.obs + (size_t)(2 * pr + q) * NP_NUM_OBS_COMBAT;
#pragma unroll
      for (int j = 0; j < NP_NUM_OBS_COMBAT; ++j) orow[j] = o[j];
    }
  }
  if (stepped) {  // blood model, after obs / reward (singlecombat_env.py:263-271)
    blood[1] = blood[1] - combat_damage(g, 1);
    blood[0] = blood[0] - combat_damage(g, 0);
  }
}

// ---- role-sharded combat records: [n][kCombatRecFloats] f32, 16-byte rows -----------------------------------------------
//   0..2 final position | 3..5 es | 6 body vx | 7 blood (after the env-level reset, before this step's damage) |
//   8 this aircraft's own termination bits (1 done, 2 bad, 4 exceed; as an integer bit pattern) | 9 10 body vy vz | 11 vt |
//   12..15 sin / cos roll, sin / cos pitch | 16..27 position after sub-steps 0..3 (the Crash check runs every sub-step; the
//   last sub-step's position is the final one)
constexpr int kCombatRecFloats = 28;
constexpr int kCombatMaxSub = 5;
__device__ __forceinline__ void combat_rec_store(float* row, const CombatRec& r, float blood, int bits) {
  float4* v = reinterpret_cast<float4*>(row);
  v[0] = make_float4(r.pos[0], r.pos[1], r.pos[2], r.es[0]);
  v[1] = make_float4(r.es[1], r.es[2], r.vel[0], blood);
  v[2] = make_float4(__int_as_float(bits), r.vel[1], r.vel[2], r.vt);
  v[3] = make_float4(r.sphi, r.cphi, r.st, r.ct);
}
struct CombatRecFull {
  CombatRec r;
  float blood;
  int bits;
  float sub_pos[kCombatMaxSub - 1][3];
};
template <bool PEER>
__device__ __forceinline__ CombatRecFull combat_rec_load(const float* row) {
  const float4* v = reinterpret_cast<const float4*>(row);
  float4 q[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) q[j] = PEER ? __ldcg(v + j) : v[j];   // a peer GPU wrote it: never through the read-only path
  CombatRecFull f;
  f.r.pos[0] = q[0].x; f.r.pos[1] = q[0].y; f.r.pos[2] = q[0].z; f.r.es[0] = q[0].w;
  f.r.es[1] = q[1].x; f.r.es[2] = q[1].y; f.r.vel[0] = q[1].z; f.blood = q[1].w;
  f.bits = __float_as_int(q[2].x); f.r.vel[1] = q[2].y; f.r.vel[2] = q[2].z; f.r.vt = q[2].w;
  f.r.sphi = q[3].x; f.r.cphi = q[3].y; f.r.st = q[3].z; f.r.ct = q[3].w;
  const float sp[12] = {q[4].x, q[4].y, q[4].z, q[4].w, q[5].x, q[5].y, q[5].z, q[5].w, q[6].x, q[6].y, q[6].z, q[6].w};
#pragma unroll
  for (int k = 0; k < kCombatMaxSub - 1; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) f.sub_pos[k][j] = sp[3 * k + j];
  return f;
}

