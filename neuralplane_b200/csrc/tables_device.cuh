// K6: the TABLE aero back-end (SURVEY f-3) -- the NASA F-16 tables the reference's MLP surrogates were fitted to
// (example/data/*.dat), evaluated by multilinear interpolation exactly as example/train_model/mexndinterp.py:84-110
// does (hyper-cube lookup, then successive linear interpolation, alpha fastest) and combined into the 44 coefficients
// of example/train_model/hifi_F16_AeroData.py:406-483 in the row order of the golden file
// envs/models/F16/model/coefs.csv.  All 13 405 table values + 61 breakpoints (54 KB) are staged in shared memory once
// per CTA; each thread owns one (alpha, beta, el) point.  Inputs are clamped to the grids (the reference returns
// garbage for the whole batch outside them, mexndinterp.py:20-21).
#pragma once
#include <stdint.h>

#include "f16_layout.h"

namespace npl {

constexpr int kNA1 = 20, kNA2 = 14, kNB1 = 19, kND1 = 5, kND2 = 3;
constexpr int kBpA1 = 0, kBpA2 = kBpA1 + kNA1, kBpB1 = kBpA2 + kNA2, kBpD1 = kBpB1 + kNB1, kBpD2 = kBpD1 + kND1;
constexpr int kBpFloats = 64;  // 61 used
constexpr int kNumTables = 43, kNumTableCoefs = 44, kTableValues = 13405;

// table order = neuralplane_b200/data/f16_tables.npz `names` (tools/pack_f16_tables.py)
enum Tab : int {
  tCx, tCz, tCm, tCy, tCn, tCl, tCx_lef, tCz_lef, tCm_lef, tCy_lef, tCn_lef, tCl_lef,
  tCXq, tCZq, tCMq, tCYp, tCYr, tCNr, tCNp, tCLp, tCLr,
  tdCXq_lef, tdCYr_lef, tdCYp_lef, tdCZq_lef, tdCLr_lef, tdCLp_lef, tdCMq_lef, tdCNr_lef, tdCNp_lef,
  tCy_r30, tCn_r30, tCl_r30, tCy_a20, tCy_a20_lef, tCn_a20, tCn_a20_lef, tCl_a20, tCl_a20_lef,
  tdCNbeta, tdCLbeta, tdCm, tEta
};
constexpr int table_size(int t) {
  return (t == tCx || t == tCz || t == tCm) ? kNA1 * kNB1 * kND1
       : (t == tCn || t == tCl) ? kNA1 * kNB1 * kND2
       : (t == tCy || (t >= tCy_r30 && t <= tCl_a20 && t != tCy_a20_lef && t != tCn_a20_lef)) ? kNA1 * kNB1
       : ((t >= tCx_lef && t <= tCl_lef) || t == tCy_a20_lef || t == tCn_a20_lef || t == tCl_a20_lef) ? kNA2 * kNB1
       : (t >= tCXq && t <= tCLr) ? kNA1
       : (t >= tdCXq_lef && t <= tdCNp_lef) ? kNA2
       : (t >= tdCNbeta && t <= tdCm) ? kNA1
       : kND1;
}
constexpr int table_offset(int t) {
  int off = kBpFloats;
  for (int i = 0; i < t; ++i) off += table_size(i);
  return off;
}
static_assert(table_offset(kNumTables) == kBpFloats + kTableValues, "table sizes must add up to the packed file");
constexpr int kTablesFloats = ((kBpFloats + kTableValues + 3) / 4) * 4;

struct Cell {
  int i;      // lower grid index
  float lam;  // (x - bp[i]) / (bp[i+1] - bp[i]) in [0, 1]
};
__device__ __forceinline__ Cell find_cell(const float* bp, int n, float x) {
  x = fminf(fmaxf(x, bp[0]), bp[n - 1]);
  int i = 0;
#pragma unroll 1
  for (int j = 1; j < n - 1; ++j) i += (bp[j] <= x) ? 1 : 0;   // grids are tiny (<= 20): a counted scan, no divergence
  Cell c;
  c.i = i;
  c.lam = (x - bp[i]) / (bp[i + 1] - bp[i]);
  return c;
}
// the same scan with the grid size known at compile time: every breakpoint load is independent of the others (warp-
// broadcast LDS), so the scan costs its instruction count, not N shared-memory latencies (the step kernel's version)
template <int N>
__device__ __forceinline__ Cell find_cell_u(const float* bp, float x) {
  float g[N];
#pragma unroll
  for (int j = 0; j < N; ++j) g[j] = bp[j];
  x = fminf(fmaxf(x, g[0]), g[N - 1]);
  int i = 0;
#pragma unroll
  for (int j = 1; j < N - 1; ++j) i += (g[j] <= x) ? 1 : 0;
  Cell c;
  c.i = i;
  c.lam = (x - bp[i]) / (bp[i + 1] - bp[i]);
  return c;
}
// The same cell -- i = #{interior breakpoints <= x}, lam -- from a GUESS of i that is then walked to the exact cell with
// compares against the grid itself: any guess gives the scan's result on any increasing grid (the walk is a loop), a good
// guess makes the walk 0 or 1 step.  ~15 instructions instead of ~60 for the 20-point alpha grid (the scan loads every
// breakpoint and compares it); the guesses below are for the NASA grids of example/data (ALPHA1 = -20..60 step 5, 70, 80,
// 90; BETA1 = -30..-10 step 5, -10..10 step 2, 10..30 step 5).
template <int N>
__device__ __forceinline__ Cell cell_from_guess(const float* bp, float x, int i) {
  x = fminf(fmaxf(x, bp[0]), bp[N - 1]);
  i = min(max(i, 0), N - 2);
  while (i > 0 && x < bp[i]) --i;
  while (i < N - 2 && x >= bp[i + 1]) ++i;
  Cell c;
  c.i = i;
  c.lam = (x - bp[i]) / (bp[i + 1] - bp[i]);
  return c;
}
__device__ __forceinline__ Cell find_cell_alpha1(const float* bp, float x) {
  const int guess = x < 60.0f ? (int)((x + 20.0f) * 0.2f) : 16 + (x >= 70.0f) + (x >= 80.0f);
  return cell_from_guess<kNA1>(bp, x, guess);
}
// ALPHA2 (-20..45 step 5) is the first 14 points of ALPHA1: below 45 deg the cell is ALPHA1's, at and above it the clamped
// point sits on the last breakpoint (i = 12, lam = (45 - 40) / (45 - 40) = 1) -- what find_cell_u<kNA2> returns, without a scan
__device__ __forceinline__ Cell cell_alpha2_from_alpha1(const float* bp2, const Cell& a1, float x) {
  Cell c = a1;
  if (!(x < bp2[kNA2 - 1])) { c.i = kNA2 - 2; c.lam = 1.0f; }
  return c;
}
__device__ __forceinline__ Cell find_cell_beta1(const float* bp, float x) {
  const int guess = x < -10.0f ? (int)((x + 30.0f) * 0.2f) : x < 10.0f ? 4 + (int)((x + 10.0f) * 0.5f) : 14 + (int)((x - 10.0f) * 0.2f);
  return cell_from_guess<kNB1>(bp, x, guess);
}

// successive linear interpolation, alpha first (mexndinterp.py:50-81): lambda * f2 + (1 - lambda) * f1
// (the second product is fused: one rounding fewer than the fp32 restatement, and FMUL + FFMA instead of FMUL, FMUL, FADD --
// the table path is checked against the float64 oracle, tests/test_gpu_tables*.py)
__device__ __forceinline__ float lerp(float f1, float f2, float lam) { return fmaf(lam, f2, (1.0f - lam) * f1); }
__device__ __forceinline__ float tab1(const float* t, Cell a) { return lerp(t[a.i], t[a.i + 1], a.lam); }
__device__ __forceinline__ float tab2(const float* t, int na, Cell a, Cell b) {
  const float* r0 = t + a.i + na * b.i;
  return lerp(lerp(r0[0], r0[1], a.lam), lerp(r0[na], r0[na + 1], a.lam), b.lam);
}
__device__ __forceinline__ float tab3(const float* t, int na, int nb, Cell a, Cell b, Cell d) {
  const float* p0 = t + na * nb * d.i;
  return lerp(tab2(p0, na, a, b), tab2(p0 + na * nb, na, a, b), d.lam);
}

// The 44 coefficients of one point, coefs.csv row order; out[k * stride].
__device__ __forceinline__ void table_coefficients(const float* T, float alpha, float beta, float el, float* out, int stride) {
  const Cell a1 = find_cell(T + kBpA1, kNA1, alpha), a2 = find_cell(T + kBpA2, kNA2, alpha), b1 = find_cell(T + kBpB1, kNB1, beta);
  const Cell d1 = find_cell(T + kBpD1, kND1, el), d2 = find_cell(T + kBpD2, kND2, el);
  const Cell z1 = find_cell(T + kBpD1, kND1, 0.0f), z2 = find_cell(T + kBpD2, kND2, 0.0f);   // el = 0 (hifi_C_lef etc.)
#define T3(t, d) tab3(T + table_offset(t), kNA1, kNB1, a1, b1, d)
#define T2(t) tab2(T + table_offset(t), kNA1, a1, b1)
#define T2L(t) tab2(T + table_offset(t), kNA2, a2, b1)
#define T1(t) tab1(T + table_offset(t), a1)
#define T1L(t) tab1(T + table_offset(t), a2)
  const float Cy = T2(tCy);
  const float Cx0 = T3(tCx, z1), Cz0 = T3(tCz, z1), Cm0 = T3(tCm, z1), Cn0 = T3(tCn, z2), Cl0 = T3(tCl, z2);
  const float Cy_lef = T2L(tCy_lef), Cn_lef = T2L(tCn_lef), Cl_lef = T2L(tCl_lef);
  const float dCy_a20 = T2(tCy_a20) - Cy, dCn_a20 = T2(tCn_a20) - Cn0, dCl_a20 = T2(tCl_a20) - Cl0;
  int k = 0;
#define PUT(v) out[(k++) * stride] = (v)
  PUT(T3(tCx, d1)); PUT(T3(tCz, d1)); PUT(T3(tCm, d1)); PUT(Cy); PUT(T3(tCn, d2)); PUT(T3(tCl, d2));                 // hifi_C
  PUT(T1(tCXq)); PUT(T1(tCYr)); PUT(T1(tCYp)); PUT(T1(tCZq)); PUT(T1(tCLr)); PUT(T1(tCLp)); PUT(T1(tCMq));           // hifi_damping
  PUT(T1(tCNr)); PUT(T1(tCNp));
  PUT(T2L(tCx_lef) - Cx0); PUT(T2L(tCz_lef) - Cz0); PUT(T2L(tCm_lef) - Cm0); PUT(Cy_lef - Cy); PUT(Cn_lef - Cn0);   // hifi_C_lef
  PUT(Cl_lef - Cl0);
  PUT(T1L(tdCXq_lef)); PUT(T1L(tdCYr_lef)); PUT(T1L(tdCYp_lef)); PUT(T1L(tdCZq_lef)); PUT(T1L(tdCLr_lef));          // hifi_damping_lef
  PUT(T1L(tdCLp_lef)); PUT(T1L(tdCMq_lef)); PUT(T1L(tdCNr_lef)); PUT(T1L(tdCNp_lef));
  PUT(T2(tCy_r30) - Cy); PUT(T2(tCn_r30) - Cn0); PUT(T2(tCl_r30) - Cl0);                                            // hifi_rudder
  PUT(dCy_a20); PUT(dCn_a20); PUT(dCl_a20);                                                                         // hifi_ailerons
  PUT(T2L(tCy_a20_lef) - Cy_lef - dCy_a20); PUT(T2L(tCn_a20_lef) - Cn_lef - dCn_a20); PUT(T2L(tCl_a20_lef) - Cl_lef - dCl_a20);
  PUT(T1(tdCNbeta)); PUT(T1(tdCLbeta)); PUT(T1(tdCm)); PUT(tab1(T + table_offset(tEta), d1)); PUT(0.0f);             // hifi_other_coeffs
#undef PUT
#undef T3
#undef T2
#undef T2L
#undef T1
#undef T1L
}

// The coefficients the env's nlplant consumes (F16_dynamics.py:167-175), written to the step kernel's slots: c[k] for
// the Coef slots k < kNumSlots (f16_layout.h) and a1[k - kFirstA1] for the alpha-only ones -- the same expressions as
// table_coefficients() above, minus the two rows nlplant never reads (delta_Czq_lef, delta_Cm_ds).  z1 / z2 are the
// el = 0 cells of the two elevator grids (hifi_C_lef etc. subtract the el = 0 value), computed once per thread.
struct ZeroCells {
  Cell z1, z2;
};
__device__ __forceinline__ ZeroCells zero_cells(const float* T) {
  ZeroCells z;
  z.z1 = find_cell(T + kBpD1, kND1, 0.0f);
  z.z2 = find_cell(T + kBpD2, kND2, 0.0f);
  return z;
}
// full = false: only what the force equations read (Cx_tot, Cy_tot, Cz_tot: the Overload evaluation, F16_model.py:132-148).
__device__ __forceinline__ void table_env_coefs(const float* T, const ZeroCells& z, float alpha, float beta, float el, bool full,
                                                float* __restrict__ c, float* __restrict__ a1) {
#ifdef NPLANE_TAB_SCAN
  const Cell ca = find_cell_u<kNA1>(T + kBpA1, alpha), cl = find_cell_u<kNA2>(T + kBpA2, alpha), cb = find_cell_u<kNB1>(T + kBpB1, beta);
#else
  const Cell ca = find_cell_alpha1(T + kBpA1, alpha), cl = cell_alpha2_from_alpha1(T + kBpA2, ca, alpha), cb = find_cell_beta1(T + kBpB1, beta);
#endif
  const Cell d1 = find_cell_u<kND1>(T + kBpD1, el);
#define T3(t, d) tab3(T + table_offset(t), kNA1, kNB1, ca, cb, d)
#define T2(t) tab2(T + table_offset(t), kNA1, ca, cb)
#define T2L(t) tab2(T + table_offset(t), kNA2, cl, cb)
#define T1(t) tab1(T + table_offset(t), ca)
#define T1L(t) tab1(T + table_offset(t), cl)
#define A1(k) a1[(k) - kFirstA1]
  // ---- force coefficients (force_totals, F16_dynamics.py:197-207) ----
  const float Cy = T2(tCy);
  const float Cx0 = T3(tCx, z.z1), Cz0 = T3(tCz, z.z1);
  const float Cy_lef = T2L(tCy_lef);
  const float dCy_a20 = T2(tCy_a20) - Cy;
  c[kCx] = T3(tCx, d1); c[kCz] = T3(tCz, d1);
  c[kCy] = Cy; c[kdCx_lef] = T2L(tCx_lef) - Cx0; c[kdCz_lef] = T2L(tCz_lef) - Cz0; c[kdCy_lef] = Cy_lef - Cy;
  c[kdCy_r30] = T2(tCy_r30) - Cy; c[kdCy_a20] = dCy_a20; c[kdCy_a20_lef] = T2L(tCy_a20_lef) - Cy_lef - dCy_a20;
  A1(kCxq) = T1(tCXq); A1(kCzq) = T1(tCZq); A1(kCyp) = T1(tCYp); A1(kCyr) = T1(tCYr);
  A1(kdCxq_lef) = T1L(tdCXq_lef); A1(kdCyr_lef) = T1L(tdCYr_lef); A1(kdCyp_lef) = T1L(tdCYp_lef);
  if (!full) return;
  // ---- moment coefficients (nlplant_kin_moments, F16_dynamics.py:208-214) ----
  const Cell d2 = find_cell_u<kND2>(T + kBpD2, el);
  const float Cm0 = T3(tCm, z.z1), Cn0 = T3(tCn, z.z2), Cl0 = T3(tCl, z.z2);
  const float Cn_lef = T2L(tCn_lef), Cl_lef = T2L(tCl_lef);
  const float dCn_a20 = T2(tCn_a20) - Cn0, dCl_a20 = T2(tCl_a20) - Cl0;
  c[kCm] = T3(tCm, d1); c[kCn] = T3(tCn, d2); c[kCl] = T3(tCl, d2);
  c[kEtaEl] = tab1(T + table_offset(tEta), d1);
  c[kdCl_a20] = dCl_a20; c[kdCl_lef] = Cl_lef - Cl0; c[kdCm_lef] = T2L(tCm_lef) - Cm0; c[kdCn_lef] = Cn_lef - Cn0;
  c[kdCn_r30] = T2(tCn_r30) - Cn0; c[kdCl_r30] = T2(tCl_r30) - Cl0; c[kdCn_a20] = dCn_a20;
  c[kdCn_a20_lef] = T2L(tCn_a20_lef) - Cn_lef - dCn_a20; c[kdCl_a20_lef] = T2L(tCl_a20_lef) - Cl_lef - dCl_a20;
  A1(kCmq) = T1(tCMq); A1(kCnr) = T1(tCNr); A1(kCnp) = T1(tCNp); A1(kClp) = T1(tCLp); A1(kClr) = T1(tCLr);
  A1(kdCnbeta) = T1(tdCNbeta); A1(kdClbeta) = T1(tdCLbeta); A1(kdCm) = T1(tdCm);
  A1(kdClr_lef) = T1L(tdCLr_lef); A1(kdClp_lef) = T1L(tdCLp_lef);
  A1(kdCmq_lef) = T1L(tdCMq_lef); A1(kdCnr_lef) = T1L(tdCNr_lef); A1(kdCnp_lef) = T1L(tdCNp_lef);
#undef A1
#undef T3
#undef T2
#undef T2L
#undef T1
#undef T1L
}

}  // namespace npl
