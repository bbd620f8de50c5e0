/* Test infrastructure (not shipped code): enumerates the identity the kernels' fmod_twopi() relies on
 * (neuralplane_b200/csrc/f16_device.cuh): for EVERY float a with |a| < 6.0e6 the three-instruction remainder
 *     q = trunc(a * RN(1/2pi));  r = fma(-q, 2pi, a);  one corrective step when r has the wrong sign or |r| >= 2pi
 * returns the bits of fmodf(a, 2pi) (the reference's torch `%` is fmod + a sign fix: envs/utils/utils.py:144-154).
 *   gcc -O2 -o _build/fmod_check fmod_check.c -lm ;  ./_build/fmod_check [stride]      (stride 1 = exhaustive, ~2.4e9 floats) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const float kTwoPi = 6.283185307179586f;

static float fmod_twopi(float a) {
  float q = truncf(a * (float)(1.0 / 6.283185307179586));
  float r = fmaf(-q, kTwoPi, a);
  const float sgn = copysignf(1.0f, a);
  const float rs = r * sgn;
  if (rs < 0.0f) q -= sgn;
  else if (rs >= kTwoPi) q += sgn;
  else return r == 0.0f ? copysignf(0.0f, a) : r;     // fmod keeps the dividend's sign on an exact zero
  r = fmaf(-q, kTwoPi, a);
  return r == 0.0f ? copysignf(0.0f, a) : r;
}

int main(int argc, char** argv) {
  const uint32_t stride = argc > 1 ? (uint32_t)strtoul(argv[1], 0, 10) : 1;
  uint32_t lim;
  const float limf = 6.0e6f;
  memcpy(&lim, &limf, 4);
  unsigned long long n = 0, bad = 0;
  for (int sign = 0; sign < 2; ++sign)
    for (uint32_t u = 0; u < lim; u += stride) {
      const uint32_t bits = u | ((uint32_t)sign << 31);
      float a, got, want;
      memcpy(&a, &bits, 4);
      got = fmod_twopi(a);
      want = fmodf(a, kTwoPi);
      uint32_t g, w;
      memcpy(&g, &got, 4);
      memcpy(&w, &want, 4);
      ++n;
      if (g != w && ++bad < 10) printf("MISMATCH a=%a got=%a want=%a\n", a, got, want);
    }
  printf("fmod_twopi vs fmodf: %llu floats (|a| < 6.0e6, stride %u), %llu mismatches\n", n, stride, bad);
  return bad != 0;
}
