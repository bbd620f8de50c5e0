/* TEST INFRASTRUCTURE (not product code).  Exhaustive proof-by-enumeration that the 3-operation constant division used by
 * the CUDA kernels (csrc/f16_device.cuh: divc) returns the correctly rounded IEEE quotient x / d -- i.e. the very bits
 * of the reference's `tensor / constant` -- for EVERY float x with 2^-100 <= |x| < 2^100 and every divisor listed.
 *     q = RN(x * RN(1/d));  e = fma(-q, d, x) (exact);  y = fma(e, RN(1/d), q)
 * usage: divc_check [stride]   (stride 1 = all 2 x 200 x 2^23 floats per divisor; the CPU test suite uses a large odd stride)
 * build: gcc -O2 -fopenmp -o _build/divc_check divc_check.c -lm                                                          */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const float kDivisors[] = {340.0f, 1000.0f, 3.14159265358979323846f, 45.0f, 0.3048f, 3.0f, 5000.0f, 300.0f,
                                  0.225f, 76300.0f, 180.0f, 2.0f * 3.14159265358979323846f, 36.0f, 6.0f, 9.0f, 18.0f};

int main(int argc, char** argv) {
  const uint32_t stride = argc > 1 ? (uint32_t)strtoul(argv[1], 0, 10) : 1u;
  const int nd = (int)(sizeof(kDivisors) / sizeof(kDivisors[0]));
  long long total_bad = 0, total = 0;
  for (int k = 0; k < nd; ++k) {
    const float d = kDivisors[k];
    const volatile float rv = 1.0f / d;
    const float r = rv;
    long long bad = 0, cnt = 0;
    const uint32_t lo = (uint32_t)(127 - 100) << 23, hi = (uint32_t)(127 + 100) << 23;
#pragma omp parallel for reduction(+ : bad, cnt) schedule(static)
    for (int64_t b = lo; b < (int64_t)hi; b += stride) {
      for (int sgn = 0; sgn < 2; ++sgn) {
        const uint32_t bits = (uint32_t)b | ((uint32_t)sgn << 31);
        float x;
        memcpy(&x, &bits, 4);
        const volatile float want = x / d;
        const volatile float q = x * r;
        const float e = fmaf(-q, d, x);
        const float y = fmaf(e, r, q);
        uint32_t yb, wb;
        const float w = want;
        memcpy(&yb, &y, 4);
        memcpy(&wb, &w, 4);
        bad += yb != wb;
        ++cnt;
      }
    }
    printf("d = %-12.9g  checked %lld  mismatches %lld\n", (double)d, cnt, bad);
    total_bad += bad;
    total += cnt;
  }
  printf("TOTAL checked %lld mismatches %lld\n", total, total_bad);
  return total_bad != 0;
}
