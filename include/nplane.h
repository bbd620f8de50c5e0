/*
 * nplane.h -- C ABI of libnplane.so: the B200 (sm_100a) implementation of NeuralPlane's
 * vectorised F-16 flight-dynamics step (ControlEnv.step / reset and the F16Model plug-in getters).
 *
 * The reference (xuecy22/NeuralPlane) is pure Python/PyTorch and has no FFI; the entry points below
 * are what a binding for this path replaces, cited as reference file:line (relative to the reference
 * repo root).  A maintainer-side ctypes stub is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns an int status (NP_OK == 0); np_last_error() gives the thread-local text.
 *   - all pointers named *_dev are DEVICE pointers owned by the caller (PyTorch allocates them); the
 *     library only borrows them.  The library owns the small aero blob it uploads in np_aero_create.
 *   - per-aircraft state is SoA, field-major: row f of a [F][ld] array starts at base + f*ld, ld >= n,
 *     ld % 4 == 0.  obs is row-major [n][22] (what GPUVecEnv hands to the runners, env_wrappers.py:97).
 *   - `stream` is a cudaStream_t passed as void*; step/reset only ENQUEUE work on it, never synchronise,
 *     and are CUDA-graph capturable.  Handles are not thread-safe; distinct handles are independent.
 *   - every entry point runs on the device its handle was created on (or that owns its output pointer), whatever the
 *     caller's current device is, and restores the current device before returning (the reference takes device=
 *     per env: env_base.py:21).  `stream` must belong to that device.
 */
#ifndef NPLANE_H_
#define NPLANE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NP_ABI_VERSION 8

enum { NP_OK = 0, NP_EINVAL = 1, NP_ECUDA = 2, NP_ESTATE = 3 };

enum { NP_TASK_HEADING = 0, NP_TASK_CONTROL = 1, NP_TASK_TRACKING = 2 };

/* aircraft plug-in (control_env.py:22-27): F16 = envs/models/F16_model.py, UAV = envs/models/UAV_model.py */
enum { NP_MODEL_F16 = 0, NP_MODEL_UAV = 1 };

#define NP_NUM_NETS 43       /* hifi_F16_AeroData.py:44-129 */
#define NP_NUM_STATE 12      /* F16_model.py:19 */
#define NP_NUM_CTRL 5        /* F16_model.py:21 (T, el, ail, rud, lef; lef is always 0, F16_model.py:57) */
#define NP_NUM_TGT 3         /* heading: alt,psi,vt | control: theta,psi,vt | tracking: npos,epos,alt */
#define NP_NUM_OBS 22        /* heading_task.py:71-152 */
#define NP_NUM_OBS_COMBAT 15 /* singlecombat_env.py:64-138 */
#define NP_NUM_ACT 4         /* F16_model.py:52-56 */
#define NP_NUM_DRAWS 5       /* uniforms a resetting aircraft consumes: alt, vt, 3 task draws */
#define NP_NUM_COUNTERS 8

typedef struct np_aero np_aero; /* the 43 MLP surrogates, device-resident */
typedef struct np_env np_env;   /* one ControlEnv(model='F16') population */
typedef struct np_tables np_tables; /* the NASA F-16 aerodynamic tables (example/data/*.dat), device-resident */

/* One row of the packed net table (neuralplane_b200/data/f16_aero.npz `desc`). */
typedef struct np_net_desc {
  int32_t n_in;      /* 1..3 */
  int32_t sel[3];    /* which input feeds column j: 0 alpha_deg, 1 beta_deg, 2 el_deg, -1 unused */
  int32_t n_layers;  /* Linear layers incl. the 1-wide output */
  int32_t dims[5];   /* widths d0 (= n_in) .. d_n_layers (= 1), 0 padded */
  int32_t w_off;     /* float offset into the blob: per layer W[out][in] row-major, then b[out] */
  int32_t used;      /* 0: evaluated by the reference but never consumed (delta_Czq_lef) */
} np_net_desc;

/* yaml task parameters (envs/configs/{heading,control,tracking}.yaml), read via getattr(config, key, default)
 * at F16_model.py:13-29, heading_task.py:29-32, the termination conditions' __init__ and env_base.py:22. */
typedef struct np_env_cfg {
  int32_t n;            /* aircraft on this rank = num_envs * num_agents (env_base.py:25) */
  int32_t ld;           /* row pitch of the SoA arrays, >= n, multiple of 4 */
  int32_t task;         /* NP_TASK_* (control_env.py:28-35) */
  int32_t use_coef_cache; /* 1: reuse the 16 (alpha,beta)-MLP outputs of the Overload evaluation in the next step */
  uint64_t seed;        /* Philox key for reset draws / observation noise when no tape is injected */
  uint64_t index_base;  /* global index of local aircraft 0 (rank sharding; RNG streams are per global index) */
  float dt, airspeed, noise_scale;
  float altitude_limit, acceleration_limit, max_velocity, min_velocity;
  float min_alpha, max_alpha, min_beta, max_beta;
  float max_heading_increment, max_pitch_increment, max_velocities_u_increment;
  float max_distance, min_distance;
  int32_t max_check_interval, min_check_interval;
  float init_T, max_altitude, min_altitude, max_vt, min_vt;
  int32_t model;        /* NP_MODEL_* ; the UAV plug-in needs no np_aero (pass NULL to np_env_create) */
  int32_t max_steps;    /* Timeout (timeout.py:16,29); combat only (selfplay.yaml max_steps: 2000) */
  /* combat (envs/configs/selfplay.yaml; singlecombat_env.py:33-44, crash.py:16) */
  float distance_limit, target_dist, max_heading, min_heading, max_npos, min_npos, max_epos, min_epos;
  int32_t index_stride; /* global aircraft index = index_base + index_stride * local index (0 = 1).  2 in the role-sharded
                           combat layout, where a rank holds every second aircraft (all egos or all opponents) */
  int32_t combat_pairs_per_env; /* 1: SingleCombat (1-v-1).  2: MultipleCombat (2-v-2, multiple_selfplay.yaml num_agents 4): an
                           env is two adjacent duels [ego0 enm0 ego1 enm1]; any flag re-initialises all four (0 = 1) */
  float combat_reward_scale; /* 0.01 (singlecombat_env.py:176-177; the default when 0) or 1 (multiplecombat_env.py:176-177) */
} np_env_cfg;

/* Device buffers the env works on (replace the tensors of F16_model.py:19-22, heading_task.py:26-28,
 * env_base.py:30-33). */
typedef struct np_buffers {
  /* all SoA rows: 16-byte aligned base, ld even and >= n rounded up to even (the step kernel moves float2 pairs) */
  float* s_dev;            /* [12][ld] state rows: npos epos alt phi theta psi vt alpha beta P Q R */
  float* u_dev;            /* [5][ld]  controls: T el ail rud lef (row 4 stays 0 and is never read) */
  float* tgt_dev;          /* [3][ld]  task targets */
  int32_t* step_count_dev; /* [ld] */
  uint8_t* flags_dev;      /* [3][ld]  is_done, bad_done, exceed_time_limit (0/1) */
  float* obs_dev;          /* [n][22] row-major */
  float* reward_dev;       /* [n] */
  void* workspace_dev;     /* np_env_workspace_bytes() bytes, zero-initialised by the caller */
} np_buffers;

int np_version(void);
/* Copies the calling thread's last error text (NUL terminated) into buf; returns its length. */
size_t np_last_error(char* buf, size_t cap);

/* Upload the 43 nets.  blob/descs/norm are HOST pointers in the layout of f16_aero.npz:
 * norm[k] = {in_mean[3], in_std[3], out_mean, out_std} (mean_std.csv; hifi_F16_AeroData.py:32-37,149-166). */
int np_aero_create(const float* blob, size_t n_floats, const np_net_desc* descs, const double* norm,
                   int n_nets, np_aero** out);
int np_aero_destroy(np_aero* aero);
/* Host-only: build the device image np_aero_create uploads (weights of the 22 multi-input MLPs + the exact
 * piecewise-linear tables of the 21 one-input nets; layout in neuralplane_b200/csrc/f16_layout.h).  With
 * out_words == NULL only *n_words is returned.  Used by the CPU tests of the table construction. */
int np_aero_pack_host(const float* blob, size_t n_floats, const np_net_desc* descs, const double* norm, int n_nets,
                      uint32_t* out_words, size_t cap_words, size_t* n_words);

size_t np_env_workspace_bytes(const np_env_cfg* cfg);
int np_env_create(const np_env_cfg* cfg, const np_aero* aero, np_env** out);
/* Binds the device buffers; invalidates the coefficient-cache keys with a memset ENQUEUED ON `stream` (the stream the
 * caller will step on), never on the legacy default stream. */
int np_env_bind(np_env* env, const np_buffers* bufs, void* stream);
/* Replace the task parameters (e.g. task.noise_scale = 0 for parity runs); n, ld and task must not change. */
int np_env_set_cfg(np_env* env, const np_env_cfg* cfg);
int np_env_destroy(np_env* env);

/* BaseEnv.reset() (env_base.py:83-97): re-initialise flagged aircraft (F16Model.reset F16_model.py:33-45,
 * task.reset e.g. heading_task.py:49-69), clear all flags, write obs.
 * draws_dev: [n][5] uniforms in [0,1) consumed by resetting aircraft, or NULL -> in-kernel Philox.
 * noise_dev: [n][22] standard normals (scaled by noise_scale), or NULL -> in-kernel Philox (skipped when
 * noise_scale == 0). */
int np_env_reset(np_env* env, const float* draws_dev, const float* noise_dev, void* stream);

/* BaseEnv.step(action) (env_base.py:99-109) as ONE kernel launch: reset -> control low-pass + Euler step
 * (F16_model.py:51-67, F16_dynamics.py:37-229) -> step_count -> obs -> terminations (task_base.py:75-96)
 * -> reward (task_base.py:60-73).  action_dev: [n][4] row-major (env_wrappers.py:94-95). */
int np_env_step(np_env* env, const float* action_dev, const float* draws_dev, const float* noise_dev,
                void* stream);

/* The same step restricted to aircraft [first_aircraft, first_aircraft + count) (first even; count even unless it ends
 * the population).  Aircraft are independent, so a step may be issued as several ranges on different streams: the
 * numpy boundary (GPUVecEnv, env_wrappers.py:93-103) pipelines H2D(actions) -> kernel -> D2H(obs) chunk by chunk.
 * action_dev is still the base of the full [n][4] array.  advance_step_index = 1 on the first range of a logical
 * step, 0 on the others (all ranges then share one RNG counter). */
int np_env_step_range(np_env* env, const float* action_dev, const float* draws_dev, const float* noise_dev, int first_aircraft,
                      int count, int advance_step_index, void* stream);

/* GPUVecEnv.step with HOST buffers (envs/env_wrappers.py:93-103: numpy actions in, numpy obs / rewards / dones out) in one
 * call: the population is cut into n_chunks aircraft ranges [edges[c], edges[c+1]) (edges[0] = 0, edges[n_chunks] = n, every
 * edge even, at most 16 chunks) and staging memcpy -> H2D -> np_env_step_range -> D2H are pipelined over three library-owned
 * streams, so the observation download (88 of the 95 B/aircraft that cross PCIe) starts after the first small chunk and never
 * idles.  action_host: [n][4] anywhere in host memory (may equal action_pinned); action_pinned / obs_pinned [n][22] /
 * reward_pinned [n] / flags_pinned [3][n] (is_done, bad_done, exceed_time_limit rows): page-locked host buffers;
 * action_dev: [n][4] device scratch.  Ordered after everything queued on `stream`; RETURNS WHEN THE HOST BUFFERS ARE READY.
 * Bit-identical to np_env_step (all chunks share one RNG counter).  ControlEnv steps only (F16, F16 tables, UAV). */
int np_env_step_host(np_env* env, const float* action_host, float* action_pinned, float* action_dev, float* obs_pinned,
                     float* reward_pinned, uint8_t* flags_pinned, const int* edges, int n_chunks, void* stream);

/* The same boundary with NO copy engine in the path: the step kernel reads the actions from, and writes observation rows,
 * rewards and flags straight into, MAPPED page-locked host memory (cudaHostAlloc / cudaHostRegister; torch pin_memory()).
 * The step is compute-bound (~0.46 ms per 10^6 aircraft) while the 95 B/aircraft that cross PCIe need ~1.7 ms, so the
 * posted writes drain under the arithmetic: one launch, no chunk pipeline.  action_host: [n][4] anywhere in host memory
 * (may equal action_mapped); action_mapped [n][4], obs_mapped [n][22], reward_mapped [n], flags_mapped [3][flags_ld]
 * (is_done, bad_done, exceed_time_limit rows; flags_ld even, >= n): page-locked, device-mapped host buffers;
 * action_dev: NULL (zero-copy action reads) or [n][4] device scratch (explicit H2D copy before the launch).  The env's own
 * obs / reward device buffers are NOT written by this call.  Ordered after everything queued on `stream`; RETURNS WHEN THE
 * HOST BUFFERS ARE READY.  Bit-identical to np_env_step.  F16 plug-in (MLP or table back-end) only. */
int np_env_step_mapped(np_env* env, const float* action_host, float* action_mapped, float* action_dev, float* obs_mapped,
                       float* reward_mapped, uint8_t* flags_mapped, int flags_ld, void* stream);

/* PlanningEnv.step(action) (envs/planning_env.py:144-177) as ONE kernel launch: reset -> clamp -> pitch / heading /
 * speed targets from the 3-D high-level action (:146-152) -> n_sub (reference: 50) x { low-level controller ->
 * F16Model.update -> freeze aircraft already terminated in this env step (:162-166) -> step_count -> terminations,
 * OR-accumulated } -> obs / reward of the last sub-step.  The low-level controller is the reference's PID stack
 * (algorithms/pid/controller.py:43-74,114-148 etc., see csrc/ctrl_device.cuh) in place of the GRU PPO actor whose
 * checkpoint the reference does not ship (planning_env.py:16).  action3_dev: [n][3] row-major.  The controller state
 * ([12][ld] floats) lives in the workspace and persists across episodes, as the reference's Controller does. */
int np_env_plan_step(np_env* env, const float* action3_dev, int n_sub, const float* draws_dev, const float* noise_dev,
                     void* stream);

/* SingleCombatEnv.step(action) (envs/singlecombat_env.py:240-274) as ONE kernel launch.  The class is stale at the
 * surveyed commit (not constructible); obs / reward / geometry / terminations / blood follow its code, the
 * orchestration is re-derived and documented in oracle/combat_oracle.py.  Population = pairs: ego = 2e, enemy = 2e+1
 * (n even); action_dev [n][4] = [throttle, roll_dem, pitch_dem, yaw]; obs is [n][15]; n_sub = 5 in the reference;
 * n_sub = 0 performs only the env-level reset + obs (SingleCombatEnv.reset).  draws_dev [n][5] = npos, epos, altitude,
 * heading, vt uniforms or NULL (Philox).  Per-pair relative geometry needs no exchange in this pair-sharded layout. */
int np_env_combat_step(np_env* env, const float* action_dev, int n_sub, const float* draws_dev, void* stream);
/* MultipleCombatEnv.step (envs/multiplecombat_env.py:240-274; envs/configs/multiple_selfplay.yaml) is the same call on an env
 * created with combat_pairs_per_env = 2, combat_reward_scale = 1 and n_sub = 1 (the reference takes ONE FDM step per env
 * step there).  The reference file pairs only agents 0 and 1 of each 4-agent env (:100-101,165-166, crash.py:32-33) and its
 * column stack cannot be built (2 E rows against 4 E aircraft), so the pairing is RESTATED: agents (0,1) and (2,3) of an env
 * are two duels evaluated with the reference's own 1-v-1 formulas, and the env-level reset (:207-238) spans all four. */
/* Role-sharded combat layout (egos and opponents on different ranks): each rank writes the 8-float record of
 * its aircraft -- position (3), inertial velocity xdot[0:3] (3), body-axis vx, blood -- into records_dev [n][8] (the
 * all-gather send slab), the ranks all-gather the slabs (NCCL, neuralplane_b200/combat_exchange.py), and
 * np_combat_relgeo evaluates the pairwise terms of singlecombat_env.py:96-121,142-177 (get_AO_TA_R / get2d_AO_TA_R,
 * envs/utils/utils.py:156-206) for m (ego, enemy) index pairs into the gathered array:
 * out_dev [m][8] = AO, TA, R (3-D), AO2, TA2, R2 (2-D), side flag, enemy - ego body-vx. */
int np_env_combat_records(np_env* env, float* records_dev, void* stream);
int np_combat_relgeo(const float* records_dev, const int32_t* ego_idx_dev, const int32_t* enm_idx_dev, float* out_dev, int m,
                     void* stream);
/* np_combat_relgeo with the exchange fused into the kernel (no gathered array, no NCCL call): slabs_dev is a DEVICE array
 * of `world` pointers, slabs_dev[r] = rank r's record slab [n_local][8] in peer-mapped memory (NVLink P2P, e.g. the
 * buffer_ptrs_dev of a torch symmetric-memory tensor); ego / enemy indices address the virtual rank-ordered gathered
 * array (record g = slabs_dev[g / n_local] + 8 (g % n_local)).  The caller orders it after every rank's
 * np_env_combat_records with a cross-device barrier.  Same out[m][8] as np_combat_relgeo on the all-gathered array. */
int np_combat_relgeo_peers(const float* const* slabs_dev, int world, int n_local, const int32_t* ego_idx_dev, const int32_t* enm_idx_dev,
                           float* out_dev, int m, void* stream);
/* ROLE-sharded combat step (BASELINE configs[4]: "NCCL all-gather for opponent relative geometry"): a rank holds
 * ONE aircraft of every env -- all egos (role 0) or all opponents (role 1); local aircraft i is global aircraft
 * index_base + 2 i (np_env_cfg.index_stride = 2).  One env step = np_env_combat_role_local (env-level reset from the
 * pair-reset flags, n_sub <= 5 FDM sub-steps under the attitude controller, per-aircraft terminations; publishes a 28-float
 * record per aircraft into records_dev [n][28]: final position, inertial velocity, body velocity, blood, own termination
 * bits, roll / pitch trigonometry, and the position after every sub-step) -> a cross-device barrier (or an all-gather of the
 * slabs) -> np_env_combat_role_pair (pulls the partner's record from partner_records_dev -- the PEER rank's slab mapped over
 * NVLink when partner_is_peer = 1, read by the kernel's own loads; or a slice of a gathered array -- and produces Crash at
 * every sub-step, Shutdown, the 15-D observation, reward, blood, the final flags and the next step's env-level reset flags:
 * singlecombat_env.py:64-181,207-238,263-271, crash.py:29-42, shutdown.py:30-40).  Every output is bit-identical to
 * np_env_combat_step on the pair-sharded layout.  n_sub = 0: SingleCombatEnv.reset. */
int np_env_combat_role_local(np_env* env, const float* action_dev, int n_sub, const float* draws_dev, float* records_dev, void* stream);
int np_env_combat_role_pair(np_env* env, const float* own_records_dev, const float* partner_records_dev, int partner_is_peer, int role,
                            int n_sub, void* stream);
#define NP_COMBAT_RECORD_FLOATS 28
/* Byte offset, inside the workspace, of the [ld] u8 env-level reset flags of the role-sharded layout (1 = re-initialise). */
size_t np_env_pair_reset_offset_bytes(const np_env_cfg* cfg);

/* Byte offset, inside the workspace, of the [ld] f32 blood row (singlecombat_env.py:45). */
size_t np_env_blood_offset_bytes(const np_env_cfg* cfg);

/* Byte offset, inside the workspace, of the controller state block [12][ld] f32: rows {roll, pitch, yaw, speed} x
 * {error, integrator, last_out} (pid.py:22-33, rollController.py:24,40); in combat mode rows 9, 10 hold the low-passed
 * roll / pitch demands (singlecombat_env.py:246-247).  Lets the host wrapper expose it. */
size_t np_env_pid_offset_bytes(const np_env_cfg* cfg);
/* started = 0 re-arms the first-call initialisation of the PIDs (PID.reset, pid.py:13,22-27). */
int np_env_set_pid_started(np_env* env, int started);

/* Device-resident rollout buffer (SURVEY f-2; reference algorithms/utils/buffer.py, runner/F16sim_runner.py:141-157).
 * np_env_rebind_outputs: point the env's obs [n][D] / reward [n] outputs at another device location (8-byte aligned), e.g.
 *   slot t+1 of the rollout buffer's obs array, so the step kernel writes the rollout in place (no copy).  Unlike
 *   np_env_bind it leaves the coefficient cache valid.
 * np_rollout_masks: masks / bad_masks [num_envs][num_agents] f32 (0 where ANY agent of the env is done / bad_done) and
 *   reset_env [num_envs] u8 (may be null) from the env's flag rows flags_dev [3][ld] (F16sim_runner.py:144-155).
 * np_rollout_returns: ReplayBuffer.compute_returns (buffer.py:139-172), all four (use_gae, use_proper_time_limits)
 *   variants, on [T][M] rewards / [T+1][M] value_preds, masks, bad_masks, returns (M = envs * agents); the caller has
 *   written next_value into value_preds[T] (use_gae) or returns[T].  Arithmetic is numpy's fp32 order: bit-exact. */
int np_env_rebind_outputs(np_env* env, float* obs_dev, float* reward_dev);
int np_rollout_masks(const uint8_t* flags_dev, int ld, int num_envs, int num_agents, float* masks_dev, float* bad_masks_dev,
                     uint8_t* reset_env_dev, void* stream);
int np_rollout_returns(const float* rewards_dev, float* value_preds_dev, const float* masks_dev, const float* bad_masks_dev,
                       float* returns_dev, int T, int M, double gamma, double gae_lambda, int use_gae, int use_proper_time_limits,
                       void* stream);

/* Termination-cause counters accumulated on device since creation (replaces the per-condition
 * print(torch.sum(bad_done)) host syncs, e.g. overload.py:32-34).  Synchronises `stream`.
 * out[0..7] = overload, low_altitude, high_speed, low_speed, extreme_state, unreach, reached(done), resets. */
int np_env_counters(np_env* env, uint64_t* out, void* stream);

/* (no reference counterpart: the reference draws from torch's global generator on the host, env_base.py:83-97,
 * heading_task.py:152.)  Every step call advances the env's RNG counter on the HOST, so a captured CUDA graph replays
 * with frozen counters -- the same reset / observation-noise draws on every replay.  This enqueues a one-thread kernel
 * that adds `delta` to a device-side epoch word all kernels add to their counter: capture it at the end of a graph of K
 * step calls with delta = K and replay r continues exactly where an eager loop of r*K steps would be (the results are
 * bit-identical to that loop).  Zero at bind. */
int np_env_rng_advance(np_env* env, uint32_t delta, void* stream);

/* F16Dynamics.nlplant (F16_dynamics.py:37-229) on SoA rows: xdot_dev [12][ld] from s_dev [12][ld], u_dev [5][ld].
 * Backs F16Model.get_extended_state and the getters built on it (F16_model.py:47-49,75-91,132-182). */
int np_f16_nlplant(const np_aero* aero, const float* s_dev, const float* u_dev, float* xdot_dev, int n, int ld,
                   void* stream);

/* The aircraft plug-in's stand-alone update(action) (envs/models/F16_model.py:51-67, UAV_model.py:51-62; callers:
 * planning_env.py:161, example/quick_start.ipynb): clamp -> control low-pass -> ONE explicit Euler step of nlplant
 * (torchdiffeq fixed-grid euler on t = [0, dt]) on the SoA rows, in place.  recent_s_dev [12][ld] / recent_u_dev [5][ld]
 * (may be NULL) receive the state / controls the update started from (the reference rebinds recent_s = s first).
 * action_dev [n][4] row-major (the UAV plug-in reads columns 0..2).  Same arithmetic as the fused env step: bit-identical.
 * np_f16_table_update is the same for a table-backed F16 plug-in. */
int np_f16_update(const np_aero* aero, float* s_dev, float* u_dev, float* recent_s_dev, float* recent_u_dev, const float* action_dev,
                  int n, int ld, double dt, void* stream);
int np_uav_update(float* s_dev, float* u_dev, float* recent_s_dev, const float* action_dev, int n, int ld, double dt, void* stream);

/* Table aero back-end (SURVEY f-3): the NASA tables the MLP surrogates were fitted to, evaluated by multilinear
 * interpolation as example/train_model/mexndinterp.py:84-110 and combined as example/train_model/hifi_F16_AeroData.py:
 * 406-483.  breakpoints = ALPHA1(20) ALPHA2(14) BETA1(19) DH1(5) DH2(3) concatenated; values / offsets = the 43
 * tables in the order of neuralplane_b200/data/f16_tables.npz (Fortran order, alpha fastest).  All HOST pointers.
 * np_f16_table_coeffs: out_dev [44][ld] in the row order of the reference's golden file envs/models/F16/model/coefs.csv
 * (rows 3..46) from alpha_deg / beta_deg / el_deg [n] (clamped to the grids). */
int np_tables_create(const float* breakpoints, const int32_t* bp_sizes, const float* values, const int32_t* offsets, int n_tables,
                     size_t n_values, np_tables** out);
int np_tables_destroy(np_tables* tables);
int np_f16_table_coeffs(const np_tables* tables, const float* alpha_deg_dev, const float* beta_deg_dev, const float* el_deg_dev,
                        float* out_dev, int n, int ld, void* stream);
/* An F16 env whose aero coefficients come from the tables instead of the MLP surrogates: the same BaseEnv.step/reset
 * (np_env_reset / np_env_step / np_env_step_range) with hifi_F16's 42 consumed coefficients (F16_dynamics.py:167-175)
 * interpolated in the step kernel.  `tables` must outlive the env.  The planning / combat steps are MLP-only.
 * np_f16_table_nlplant = np_f16_nlplant for such an env's getters. */
int np_env_create_tables(const np_env_cfg* cfg, const np_tables* tables, np_env** out);
int np_f16_table_nlplant(const np_tables* tables, const float* s_dev, const float* u_dev, float* xdot_dev, int n, int ld,
                         void* stream);
int np_f16_table_update(const np_tables* tables, float* s_dev, float* u_dev, float* recent_s_dev, float* recent_u_dev,
                        const float* action_dev, int n, int ld, double dt, void* stream);

/* UAVDynamics.nlplant (envs/models/UAV/UAV_dynamics.py:15-84) on SoA rows: xdot_dev [12][ld] from s_dev [12][ld] and
 * the three body forces u_dev [3+][ld].  Backs UAVModel.get_extended_state and the getters built on it
 * (UAV_model.py:47-49,72-84,120-157). */
int np_uav_nlplant(const float* s_dev, const float* u_dev, float* xdot_dev, int n, int ld, void* stream);

/* The 43 coefficient nets (hifi_F16_AeroData.py:748-819): out_dev [43][ld] in f16_aero.npz order from
 * alpha_deg/beta_deg/el_deg [n]. */
int np_f16_coeffs(const np_aero* aero, const float* alpha_deg_dev, const float* beta_deg_dev, const float* el_deg_dev,
                  float* out_dev, int n, int ld, void* stream);

/* Launch geometry the step kernel uses for this env (diagnostics / bench reporting). */
int np_env_launch_info(const np_env* env, int* grid, int* block, int* smem_bytes, int* num_sms);

#ifdef __cplusplus
}
#endif
#endif /* NPLANE_H_ */
