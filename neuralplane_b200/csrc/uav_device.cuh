// Second aircraft plug-in: the force-driven 6-DoF rigid body `UAV` (reference: envs/models/UAV_model.py:10-176,
// envs/models/UAV/UAV_dynamics.py:15-84).  State is SI (metres, m/s, body velocities U V W in slots 6..8), the task
// layer sees feet through the model getters, exactly as the reference's tasks do (heading_task.py:93-102 etc.).
//
// Reference defect handled: UAVModel.update() shrinks `u` to 3 columns while the shipped yaml says
// num_controls: 5, so the reference crashes on its second step; the model has 3 controls (Fx, Fy, Fz) and that is
// what is implemented (the golden fixtures come from the unmodified reference run with num_controls = 3).
//
// Compiled with -fmad=false: every a*b+c rounds twice like the reference's eager ops.
#pragma once
#include "f16_device.cuh"

namespace npl {

// sin / cos / tan of the Euler angles, shared by nlplant and the observation of the same state
struct UavTrig {
  float st, ct, tt, sphi, cphi, spsi, cpsi;
};
__device__ __forceinline__ UavTrig uav_trig(const float* s) {
  UavTrig t;
  sincosf(s[4], &t.st, &t.ct);
  t.tt = t.st / t.ct;  // tan(theta) from the two values already needed (<= 3 ulp; tanf itself is a 4-ulp function)
  sincosf(s[3], &t.sphi, &t.cphi);
  sincosf(s[5], &t.spsi, &t.cpsi);
  return t;
}

// UAVDynamics.nlplant (UAV_dynamics.py:15-84) for one aircraft: xdot[0..11] from s[0..11], forces F[0..2].
__device__ __forceinline__ void uav_nlplant(const float* s, const float* F, const UavTrig& t, float* xdot) {
  constexpr float UAV_M = 300.0f, g = 9.81f;
  constexpr float M = 1.0f, N = 1.0f, L_bar = 1.0f, I_x = 1.0f, I_y = 1.0f, I_z = 1.0f, I_xz = 0.0f;
  const float U = s[6], V = s[7], W = s[8], P = s[9], Q = s[10], R = s[11];
  const float st = t.st, ct = t.ct, tt = t.tt, sphi = t.sphi, cphi = t.cphi, spsi = t.spsi, cpsi = t.cpsi;
  xdot[0] = U * (ct * cpsi) + V * (sphi * st * cpsi - cphi * spsi) + W * (sphi * spsi + cphi * st * cpsi);
  xdot[1] = U * (ct * spsi) + V * (sphi * st * spsi + cphi * cpsi) + W * (-sphi * cpsi + cphi * st * spsi);
  xdot[2] = U * st - V * (sphi * ct) - W * (cphi * ct);
  xdot[3] = P + (R * cphi + Q * sphi) * tt;
  xdot[4] = Q * cphi - R * sphi;
  xdot[5] = (R * cphi + Q * sphi) / ct;
  xdot[6] = V * R - W * Q - g * st + F[0] / DC(UAV_M);
  xdot[7] = -U * R + W * P + g * ct * sphi + F[1] / DC(UAV_M);
  xdot[8] = U * Q - V * P + g * ct * cphi + F[2] / DC(UAV_M);
  const float b0 = L_bar - Q * R * (I_z - I_y) + P * Q * I_xz;
  const float b1 = N - P * Q * (I_y - I_x) - Q * R * I_xz;
  const float b2 = M - P * R * (I_x - I_z) - (P * P - R * R) * I_xz;
  xdot[9] = (b0 * I_z + b1 * I_xz) / (I_z * I_x - I_xz * I_xz);
  xdot[10] = b2 / I_y;
  xdot[11] = (b0 * I_xz + b1 * I_x) / (I_z * I_x - I_xz * I_xz);
}
__device__ __forceinline__ void uav_nlplant(const float* s, const float* F, float* xdot) { uav_nlplant(s, F, uav_trig(s), xdot); }

// What the task layer reads through the model getters (UAV_model.py:69-134), in the reference's units (feet).
struct UavView {
  float npos, epos, alt, roll, pitch, heading, vt, e2t;
};
__device__ __forceinline__ UavView uav_view(const float* s) {
  UavView v;
  v.npos = s[0] / DC(0.3048f);                                            // get_position
  v.epos = s[1] / DC(0.3048f);
  v.alt = s[2] / DC(0.3048f);
  v.roll = s[3]; v.pitch = s[4]; v.heading = s[5];                    // get_posture
  v.vt = sqrtf(s[6] * s[6] + s[7] * s[7] + s[8] * s[8]) / DC(0.3048f);    // get_vt
  const float tfac = 1.0f - .703e-5f * v.alt;                         // get_EAS2TAS
  v.e2t = sqrtf(1.0f / powf(tfac, 4.14f));
  return v;
}

}  // namespace npl
