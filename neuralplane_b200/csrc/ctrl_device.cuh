// Low-level flight controller fused into the planning step (K4): the reference's ArduPilot-style PID stack, one
// aircraft per call, state in registers.
//   Controller.stabilize            algorithms/pid/controller.py:43-74   (speed scaler, roll / pitch / yaw loops)
//   RollController / PitchController / YawController rate loops
//                                   algorithms/pid/rollController.py:26-49, pitchController.py:30-94, yawController.py:69-84
//   PID                             algorithms/pid/pid.py:17-41  (gains: algorithms/pid/config/*.yaml)
//   L1 heading hold + nav_roll      algorithms/pid/L1Controller.py:230-271, controller.py:114-124
//   TAS loop                        the PID class with the gains of config/speedcontroller.yaml (the reference's
//                                   SpeedController is not runnable: speedController.py:24-45 reads undefined attributes)
// In the reference every rate loop calls env.model.get_euler_angular_velocity() -> one full nlplant each; only the
// kinematic rows (xdot[0..5]) are needed, which are pure trigonometry of the state -- computed once here.
// The PID class's population-wide NaN guard (pid.py:18-21) has no per-aircraft meaning and is not reproduced.
// Compiled with -fmad=false (two roundings per a*b+c, like the reference's eager ops).
#pragma once
#include "f16_device.cuh"

namespace npl {

constexpr int kPidRows = 12;  // {roll, pitch, yaw, speed} x {error, integrator, last_out}

struct PidGains {
  float Kp, Ki, Kd, Kff, Kimax;
};

// pid.py update_all + update_i; st = {error, integrator, last_out}; returns (ff, p + i + d pieces combined by caller)
__device__ __forceinline__ void pid_update(float* st, const PidGains& G, float target, float meas, bool limit, bool first,
                                           float dt, float& deriv) {
  const float err = target - meas;
  if (first) {
    deriv = 0.0f;
    st[1] = 0.0f;
  } else {
    deriv = (err - st[0]) / dt;
  }
  st[0] = err;
  if (G.Ki != 0.0f && dt > 0.0f) {
    const bool gate = (!limit) | (err * dt < 0.0f);
    st[1] = st[1] + err * G.Ki * dt * (gate ? 1.0f : 0.0f);
    st[1] = fminf(fmaxf(st[1], -G.Kimax), G.Kimax);
  } else {
    st[1] = 0.0f;
  }
}

// get_rate_out of the three attitude controllers (identical but for the >= / > of the integrator limit)
__device__ __forceinline__ float rate_out(float* st, const PidGains& G, float desired_rate, float rate, float scaler, float e2t,
                                          bool strict, bool first, float dt) {
  const bool limit = strict ? (fabsf(st[2]) > 45.0f) : (fabsf(st[2]) >= 45.0f);
  const float target = desired_rate * scaler * scaler;
  float deriv;
  pid_update(st, G, target, rate * scaler * scaler, limit, first, dt, deriv);
  float out = target * G.Kff / (scaler * e2t + 1e-8f) + st[0] * G.Kp + st[1] + deriv * G.Kd;
  out = 180.0f * out / DC(kPi);
  st[2] = out;
  return fminf(fmaxf(out, -45.0f), 45.0f);
}

// Kinematic quantities every loop of the stack reads through the model getters (F16_model.py:75-103,156-162).
struct CtrlKin {
  Trig g;
  float TAS, e2t, vx, vy, roll_rate, pitch_rate, yaw_rate;
};
__device__ __forceinline__ CtrlKin ctrl_kin(const float* s, float airspeed) {
  CtrlKin k;
  k.g = make_trig(s);
  const Trig& g = k.g;
  const float P = s[9], Q = s[10], R = s[11];
  k.TAS = s[6] + airspeed * 1.0f;
  k.e2t = eas2tas_of(tfac_pow(s[2]));
  // kinematic rows of nlplant (F16_dynamics.py:104,129-138): ground speed and Euler-angle rates
  const float vtc = s[6] <= 0.01f ? 0.01f : s[6];
  const BodyVel b = body_vel(vtc, g);
  k.vx = b.U * (g.ct * g.cpsi) + b.V * (g.sphi * g.cpsi * g.st - g.cphi * g.spsi) +
         b.W * (g.cphi * g.st * g.cpsi + g.sphi * g.spsi);
  k.vy = b.U * (g.ct * g.spsi) + b.V * (g.sphi * g.spsi * g.st + g.cphi * g.cpsi) +
         b.W * (g.cphi * g.st * g.spsi - g.sphi * g.cpsi);
  k.roll_rate = P + g.tt * (Q * g.sphi + R * g.cphi);
  k.pitch_rate = Q * g.cphi - R * g.sphi;
  k.yaw_rate = (Q * g.sphi + R * g.cphi) / g.ct;
  return k;
}

// Controller.stabilize (controller.py:43-74): speed scaler, roll / pitch / yaw loops.  pid[0..8] = the three rate
// PIDs; outputs surface demands in degrees (clamped to +-45).
__device__ __forceinline__ void pid_stabilize(const float* s, const CtrlKin& k, float dt, float roll_dem, float pitch_dem,
                                              float yaw_rate_dem, float* pid, bool first, float& el, float& ail, float& rud) {
  constexpr float gravity = 32.174f;
  const PidGains kRoll{10.0f, 0.3f, 0.0f, 0.3f, 0.666f}, kPitch{10.0f, 0.3f, 0.0f, 0.3f, 0.666f};
  const PidGains kYaw{1.0f, 0.3f, 0.05f, 0.3f, 0.666f};
  const Trig& g = k.g;
  const float roll = s[3], pitch = s[4], TAS = k.TAS, e2t = k.e2t;
  constexpr float scale_min = (float)(1000.0 / (2 * 2300)), scale_max = (float)(1000.0 / (0.7 * 100));
  const float scaler = fminf(fmaxf(1000.0f / (TAS + 1e-8f), scale_min), scale_max);
  ail = rate_out(pid + 0, kRoll, wrap_pi(roll_dem - roll) / 0.5f, k.roll_rate, scaler, e2t, false, first, dt);

  // pitch loop with turn coordination and inverted-flight handling (pitchController.py:47-94)
  float desired = wrap_pi(pitch_dem - pitch) / 0.5f;
  constexpr float kHalfPi = (float)(3.141592653589793 / 2);
  const bool m1 = fabsf(roll) < kHalfPi, m2 = roll >= kHalfPi, m3 = roll <= -kHalfPi;
  const float r1 = fminf(fmaxf(roll, -(float)(4 * 3.141592653589793 / 9)), (float)(4 * 3.141592653589793 / 9));
  const float r2 = fminf(fmaxf(roll, (float)(5 * 3.141592653589793 / 9)), kPi);
  const float r3 = fminf(fmaxf(roll, -kPi), -(float)(5 * 3.141592653589793 / 9));
  const bool inverted = !m1;
  const float rollc = (m1 ? r1 : 0.0f) + (m2 ? r2 : 0.0f) + (m3 ? r3 : 0.0f);
  const bool mp = fabsf(pitch) <= (float)(7 * 3.141592653589793 / 18);
  float rate_offset = (mp ? 1.0f : 0.0f) * g.ct * fabsf(gravity / TAS * tanf(rollc) * sinf(rollc) * e2t) * 1.0f;
  rate_offset = inverted ? (rate_offset * 0.0f - rate_offset * 1.0f) : (rate_offset * 1.0f - rate_offset * 0.0f);
  const float desired1 = desired + rate_offset;
  desired = inverted ? (rate_offset - desired) : desired1;
  float roll_wrapped = fabsf(roll);
  if (roll_wrapped > kHalfPi) roll_wrapped = kPi - roll_wrapped;
  const bool mk = (roll_wrapped > (float)(5 * 3.141592653589793 / 18)) & (fabsf(pitch) < (float)(7 * 3.141592653589793 / 18));
  const float roll_prop = mk ? (roll_wrapped - (float)(5 * 3.141592653589793 / 18)) / (float)(4 * 3.141592653589793 / 18) : 0.0f;
  desired = desired * (1.0f - roll_prop);
  el = rate_out(pid + 3, kPitch, desired, k.pitch_rate, scaler, e2t, true, first, dt);
  rud = rate_out(pid + 6, kYaw, yaw_rate_dem, k.yaw_rate, scaler, e2t, false, first, dt);
}

// PlanningEnv low level: one control decision for one aircraft from (pitch, heading, speed) targets:
// action[4] = (throttle, -el/45, -ail/45, -rud/45) (controller.py:140-148).  pid[12] is the persistent controller
// state; `first` marks the very first call (PID.reset, pid.py:13,22-27).
__device__ __forceinline__ void pid_controller(const float* s, float airspeed, float dt, float target_pitch, float target_heading,
                                               float target_vt, float* pid, bool first, float* action) {
  constexpr float gravity = 32.174f;
  const PidGains kSpeed{5.0f, 25.0f, 0.0f, 80.0f, 100.0f};
  const CtrlKin k = ctrl_kin(s, airspeed);
  const float yaw = s[5], TAS = k.TAS, e2t = k.e2t;

  // ---- L1 heading hold -> roll demand, yaw-rate demand (L1Controller.py:230-271, controller.py:114-124) -------
  constexpr float omegaA = (float)(4.4428 / 17);
  float Nu = wrap_pi(wrap_pi(target_heading) - wrap_pi(yaw));
  const float VomegaA = sqrtf(k.vx * k.vx + k.vy * k.vy) * omegaA;
  Nu = fminf(fmaxf(Nu, -(float)(3.141592653589793 / 2)), (float)(3.141592653589793 / 2));
  const float latAccDem = 2.0f * sinf(Nu) * VomegaA;
  float roll_dem = k.g.ct * atanf(latAccDem / gravity);
  roll_dem = fminf(fmaxf(roll_dem, -(float)(3.141592653589793 / 2)), (float)(3.141592653589793 / 2));
  roll_dem = fminf(fmaxf(roll_dem, -(float)(3.141592653589793 / 4)), (float)(3.141592653589793 / 4));
  const float yaw_rate_dem = gravity * tanf(roll_dem) / TAS * e2t;

  // ---- TAS loop -> throttle ----------------------------------------------------------------------------------
  float deriv;
  pid_update(pid + 9, kSpeed, target_vt * 0.3048f / DC(340.0f), TAS * 0.3048f / DC(340.0f), fabsf(pid[11]) >= 100.0f, first, dt, deriv);
  const float sp_out = (target_vt * 0.3048f / DC(340.0f)) * kSpeed.Kff + pid[9] * kSpeed.Kp + pid[10] + deriv * kSpeed.Kd;
  pid[11] = sp_out;
  const float throttle = fminf(fmaxf(sp_out / 100.0f, 0.0f), 1.0f);

  float el, ail, rud;
  pid_stabilize(s, k, dt, roll_dem, target_pitch, yaw_rate_dem, pid, first, el, ail, rud);
  action[0] = throttle;
  action[1] = -el / DC(45.0f);
  action[2] = -ail / DC(45.0f);
  action[3] = -rud / DC(45.0f);
}

// SingleCombatEnv low level (singlecombat_env.py:244-255): the 4-D action [throttle, roll_dem, pitch_dem, yaw] moves
// low-passed attitude demands (pid[9] = roll_dem, pid[10] = pitch_dem); yaw_rate_dem stays 0 (a3 only sets `yaw_dem`,
// which the current stabilize_yaw does not read, controller.py:60-65).
__device__ __forceinline__ void combat_controller(const float* s, float airspeed, float dt, const float* a4, float* pid, bool first,
                                                  float* action) {
  pid[9] = 0.9f * pid[9] + 0.1f * a4[1] * 4.0f * kPi / 9.0f;
  pid[10] = 0.9f * pid[10] + 0.1f * a4[2] * kPi / 12.0f;
  const CtrlKin k = ctrl_kin(s, airspeed);
  float el, ail, rud;
  pid_stabilize(s, k, dt, pid[9], pid[10], 0.0f, pid, first, el, ail, rud);
  action[0] = a4[0];
  action[1] = -el / DC(45.0f);
  action[2] = -ail / DC(45.0f);
  action[3] = -rud / DC(45.0f);
}

}  // namespace npl
