// Micro-benchmark: do the 20 -> 10 layers of the (alpha, beta) coefficient nets belong on the tensor cores?
//
// Workload: NNET nets of the commonest architecture (2, 20, 10, 1) -- ReLU MLPs, random weights -- evaluated for n aircraft.
//   ffma2 : the production path of K1 (f16_device.cuh: dense2 / mlp2): two aircraft per thread packed in FFMA2, weights as
//           warp-broadcast LDS.128 from shared memory.
//   mma3  : layer 2 (20 -> 10, bias folded in as a 21st input fixed to 1) on the tensor cores as an error-compensated 3xTF32
//           product (A_hi B_hi + A_lo B_hi + A_hi B_lo, mma.sync.m16n8k8 tf32 with fp32 accumulation), 16-aircraft M tiles, two
//           tiles per warp pass.  Layer 1 (2 -> 20) is computed on the FMA pipe DIRECTLY IN THE A-FRAGMENT LAYOUT (lane (g, t)
//           owns hidden units t, t+4, t+8, ... of aircraft g and g+8), so no shuffle or shared-memory hop feeds the MMA; layer 3
//           (10 -> 1) is a dot product over the C fragment plus a 4-lane butterfly.
// Reports time, aircraft-net evaluations per second, executed instructions are read from ncu / cuobjdump (HMMA count), and the
// maximum deviation of both paths from the float64 evaluation of the same nets, in units of the output range.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/mlp_tc_bench tools/mlp_tc_bench.cu
// (tcgen05 was not benchmarked: see DESIGN.md section 3 -- TMEM capacity (512 columns per SM for all resident tiles) and its
//  64 B/clk read port bound that design below this one's instruction budget for nets this small.)
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

constexpr int NNET = 16, H1 = 20, H2 = 10;
// ---- shared-memory images ---------------------------------------------------------------------------------------------
// ffma2: per net [b1(20) W1^T(2x20) | b2(10) W2^T(20x10) pad | b3 W3(10) pad] in 4-float groups (production layout)
constexpr int pad4(int x) { return (x + 3) & ~3; }
constexpr int kL1 = pad4(H1 + 2 * H1), kL2 = pad4(H2 + H1 * H2), kL3 = pad4(1 + H2);
constexpr int kNetF = kL1 + kL2 + kL3;
// mma3: per net  l1[24] float4 {w0, w1, b, 0} | bfrag[3][2][32] float4 {b0_hi, b1_hi, b0_lo, b1_lo} | l3[4] float4 | b3
constexpr int kMmaNetF = 24 * 4 + 3 * 2 * 32 * 4 + 4 * 4 + 4;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 q;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(a));
  return q;
}

// ---- production path ----------------------------------------------------------------------------------------------------
template <int IN, int OUT, bool RELU>
__device__ __forceinline__ void dense2(uint32_t w, const float2 (&x)[IN], float2 (&y)[OUT]) {
  constexpr int NF = OUT + IN * OUT, NV = (NF + 3) / 4;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 q = lds128(w + 16 * v);
    const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int f = 4 * v + k;
      if (f < OUT) y[f] = make_float2(e[k], e[k]);
      else if (f < NF) y[(f - OUT) % OUT] = __ffma2_rn(make_float2(e[k], e[k]), x[(f - OUT) / OUT], y[(f - OUT) % OUT]);
    }
  }
  if (RELU) {
#pragma unroll
    for (int j = 0; j < OUT; ++j) y[j] = make_float2(fmaxf(y[j].x, 0.f), fmaxf(y[j].y, 0.f));
  }
}

__global__ void __launch_bounds__(256) ffma2_kernel(const float* __restrict__ img, const float2* __restrict__ z0, const float2* __restrict__ z1,
                                                    float* __restrict__ out, int npairs, int ld) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < NNET * kNetF; i += blockDim.x) sm[i] = img[i];
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += gridDim.x * blockDim.x) {
    const float2 x[2] = {z0[p], z1[p]};
#pragma unroll 1
    for (int k = 0; k < NNET; ++k) {
      const uint32_t w = base + 4 * k * kNetF;
      float2 a[H1], b[H2], y[1];
      dense2<2, H1, true>(w, x, a);
      dense2<H1, H2, true>(w + 4 * kL1, a, b);
      dense2<H2, 1, false>(w + 4 * (kL1 + kL2), b, y);
      reinterpret_cast<float2*>(out + (size_t)k * ld)[p] = y[0];
    }
  }
}

// ---- tensor-core path ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int TILES = 2;   // 16-aircraft M tiles per warp pass: the B fragments of a net are loaded once for both

__global__ void __launch_bounds__(256) mma3_kernel(const float* __restrict__ img, const float* __restrict__ z0, const float* __restrict__ z1,
                                                   float* __restrict__ out, int n, int ld) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < NNET * kMmaNetF; i += blockDim.x) sm[i] = img[i];
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warps = (gridDim.x * blockDim.x) >> 5, wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int i0 = wid * 16 * TILES; i0 + 16 * TILES <= n; i0 += warps * 16 * TILES) {   // n is a multiple of 32 in this benchmark
    float za[TILES][2], zb[TILES][2];     // this lane's two aircraft (rows g, g + 8) of every tile
#pragma unroll
    for (int m = 0; m < TILES; ++m) {
      za[m][0] = z0[i0 + 16 * m + g]; za[m][1] = z0[i0 + 16 * m + g + 8];
      zb[m][0] = z1[i0 + 16 * m + g]; zb[m][1] = z1[i0 + 16 * m + g + 8];
    }
#pragma unroll 1
    for (int k = 0; k < NNET; ++k) {
      const uint32_t wn = base + 4 * k * kMmaNetF;
      // ---- layer 1 on the FMA pipe, straight into the A fragments: units 8 ks + t (a0, a1) and 8 ks + t + 4 (a2, a3) ----
      uint32_t ahi[TILES][3][4], alo[TILES][3][4];
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 w = lds128(wn + 16 * (8 * ks + t + 4 * h));       // {w0, w1, b, 0} of this lane's hidden unit
#pragma unroll
          for (int m = 0; m < TILES; ++m) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const float v = fmaxf(fmaf(w.y, zb[m][r], fmaf(w.x, za[m][r], w.z)), 0.0f);
              const uint32_t hi = tf32_rna(v);
              ahi[m][ks][2 * h + r] = hi;
              alo[m][ks][2 * h + r] = tf32_rna(v - __uint_as_float(hi));
            }
          }
        }
      }
      // ---- layer 2 on the tensor cores: C[16 x 16] = A[16 x 24] B[24 x 16], 3xTF32 ----
      float c[TILES][2][4];
#pragma unroll
      for (int m = 0; m < TILES; ++m)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int j = 0; j < 4; ++j) c[m][nt][j] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const float4 b = lds128(wn + 16 * (24 + (ks * 2 + nt) * 32 + lane));   // {b0_hi, b1_hi, b0_lo, b1_lo}
          const uint32_t bh0 = __float_as_uint(b.x), bh1 = __float_as_uint(b.y), bl0 = __float_as_uint(b.z), bl1 = __float_as_uint(b.w);
#pragma unroll
          for (int m = 0; m < TILES; ++m) {
            mma_tf32(c[m][nt], alo[m][ks], bh0, bh1);     // small terms first
            mma_tf32(c[m][nt], ahi[m][ks], bl0, bl1);
            mma_tf32(c[m][nt], ahi[m][ks], bh0, bh1);
          }
        }
      }
      // ---- layer 3: relu, dot with w3 over this lane's columns {2t, 2t+1, 8+2t, 9+2t}, butterfly over the quad ----
      const float4 w3 = lds128(wn + 16 * (24 + 192 + t));
      const float b3 = *reinterpret_cast<const float*>(sm + k * kMmaNetF + 4 * (24 + 192 + 4));
#pragma unroll
      for (int m = 0; m < TILES; ++m) {
        float s0 = fmaxf(c[m][0][0], 0.f) * w3.x, s1 = fmaxf(c[m][0][2], 0.f) * w3.x;      // rows g / g + 8
        s0 = fmaf(fmaxf(c[m][0][1], 0.f), w3.y, s0); s1 = fmaf(fmaxf(c[m][0][3], 0.f), w3.y, s1);
        s0 = fmaf(fmaxf(c[m][1][0], 0.f), w3.z, s0); s1 = fmaf(fmaxf(c[m][1][2], 0.f), w3.z, s1);
        s0 = fmaf(fmaxf(c[m][1][1], 0.f), w3.w, s0); s1 = fmaf(fmaxf(c[m][1][3], 0.f), w3.w, s1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        if (t == 0) out[(size_t)k * ld + i0 + 16 * m + g] = s0 + b3;
        if (t == 1) out[(size_t)k * ld + i0 + 16 * m + g + 8] = s1 + b3;
      }
    }
  }
}

// ---- host -----------------------------------------------------------------------------------------------------------------
static float tf32_host(float x) {   // cvt.rna.tf32.f32: round to nearest, ties away, on the 13 dropped mantissa bits
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 4 * 1024 * 1024;   // multiple of 32
  const int iters = argc > 2 ? atoi(argv[2]) : 20;
  srand(12345);
  auto rnd = [] { return (float)rand() / RAND_MAX * 2.0f - 1.0f; };
  std::vector<float> W1(NNET * H1 * 2), B1(NNET * H1), W2(NNET * H2 * H1), B2(NNET * H2), W3(NNET * H2), B3(NNET);
  for (auto& v : W1) v = rnd() * 1.2f;
  for (auto& v : B1) v = rnd() * 0.8f;
  for (auto& v : W2) v = rnd() * 0.5f;
  for (auto& v : B2) v = rnd() * 0.5f;
  for (auto& v : W3) v = rnd() * 0.7f;
  for (auto& v : B3) v = rnd() * 0.3f;
  std::vector<float> imgA(NNET * kNetF, 0.f), imgB(NNET * kMmaNetF, 0.f);
  for (int k = 0; k < NNET; ++k) {
    float* a = imgA.data() + k * kNetF;
    for (int j = 0; j < H1; ++j) a[j] = B1[k * H1 + j];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < H1; ++j) a[H1 + i * H1 + j] = W1[(k * H1 + j) * 2 + i];
    a += kL1;
    for (int j = 0; j < H2; ++j) a[j] = B2[k * H2 + j];
    for (int i = 0; i < H1; ++i) for (int j = 0; j < H2; ++j) a[H2 + i * H2 + j] = W2[(k * H2 + j) * H1 + i];
    a += kL2;
    a[0] = B3[k];
    for (int i = 0; i < H2; ++i) a[1 + i] = W3[k * H2 + i];
    float* b = imgB.data() + k * kMmaNetF;
    for (int u = 0; u < 24; ++u) {
      b[4 * u + 0] = u < H1 ? W1[(k * H1 + u) * 2 + 0] : 0.f;
      b[4 * u + 1] = u < H1 ? W1[(k * H1 + u) * 2 + 1] : 0.f;
      b[4 * u + 2] = u < H1 ? B1[k * H1 + u] : (u == H1 ? 1.f : 0.f);     // unit 20 is the constant 1 that carries layer 2's bias
    }
    auto Bm = [&](int kk, int nn) -> float { return nn >= H2 ? 0.f : kk < H1 ? W2[(k * H2 + nn) * H1 + kk] : kk == H1 ? B2[k * H2 + nn] : 0.f; };
    for (int ks = 0; ks < 3; ++ks) for (int nt = 0; nt < 2; ++nt) for (int l = 0; l < 32; ++l) {
      const int g = l >> 2, t = l & 3;
      const float v0 = Bm(8 * ks + t, 8 * nt + g), v1 = Bm(8 * ks + t + 4, 8 * nt + g);
      float* f = b + 4 * (24 + (ks * 2 + nt) * 32 + l);
      f[0] = tf32_host(v0); f[1] = tf32_host(v1); f[2] = tf32_host(v0 - f[0]); f[3] = tf32_host(v1 - f[1]);
    }
    for (int t = 0; t < 4; ++t) {
      float* f = b + 4 * (24 + 192 + t);
      auto w3 = [&](int c) { return c < H2 ? W3[k * H2 + c] : 0.f; };
      f[0] = w3(2 * t); f[1] = w3(2 * t + 1); f[2] = w3(8 + 2 * t); f[3] = w3(9 + 2 * t);
    }
    b[4 * (24 + 192 + 4)] = B3[k];
  }
  std::vector<float> hz0(n), hz1(n);
  for (int i = 0; i < n; ++i) { hz0[i] = rnd() * 2.5f; hz1[i] = rnd() * 2.5f; }
  float *dA, *dB, *dz0, *dz1, *dout;
  CK(cudaMalloc(&dA, imgA.size() * 4)); CK(cudaMalloc(&dB, imgB.size() * 4));
  CK(cudaMalloc(&dz0, n * 4)); CK(cudaMalloc(&dz1, n * 4)); CK(cudaMalloc(&dout, (size_t)NNET * n * 4));
  CK(cudaMemcpy(dA, imgA.data(), imgA.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, imgB.data(), imgB.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dz0, hz0.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dz1, hz1.data(), n * 4, cudaMemcpyHostToDevice));
  // interleaved copies for the pair kernel: z as float2 per aircraft pair
  std::vector<float> hp0(n), hp1(n);
  for (int i = 0; i < n; ++i) { hp0[i] = hz0[i]; hp1[i] = hz1[i]; }
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int smA = NNET * kNetF * 4, smB = NNET * kMmaNetF * 4;
  CK(cudaFuncSetAttribute(ffma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smA));
  CK(cudaFuncSetAttribute(mma3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smB));
  // float64 reference on a sample
  const int ns = 4096;
  std::vector<double> ref((size_t)NNET * ns);
  double lo = 1e30, hi = -1e30;
  for (int k = 0; k < NNET; ++k) for (int i = 0; i < ns; ++i) {
    double a[H1], b[H2];
    for (int j = 0; j < H1; ++j) a[j] = fmax(0.0, (double)W1[(k * H1 + j) * 2] * hz0[i] + (double)W1[(k * H1 + j) * 2 + 1] * hz1[i] + B1[k * H1 + j]);
    for (int j = 0; j < H2; ++j) { double s = B2[k * H2 + j]; for (int q = 0; q < H1; ++q) s += (double)W2[(k * H2 + j) * H1 + q] * a[q]; b[j] = fmax(0.0, s); }
    double y = B3[k];
    for (int q = 0; q < H2; ++q) y += (double)W3[k * H2 + q] * b[q];
    ref[(size_t)k * ns + i] = y; lo = fmin(lo, y); hi = fmax(hi, y);
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<float> got((size_t)NNET * ns);
  for (int variant = 0; variant < 2; ++variant) {
    for (int blocks_per_sm = 1; blocks_per_sm <= 4; ++blocks_per_sm) {
      const int grid = sms * blocks_per_sm;
      auto launch = [&] {
        if (variant == 0) ffma2_kernel<<<grid, 256, smA>>>(dA, reinterpret_cast<const float2*>(dz0), reinterpret_cast<const float2*>(dz1), dout, n / 2, n);
        else mma3_kernel<<<grid, 256, smB>>>(dB, dz0, dz1, dout, n, n);
      };
      CK(cudaMemset(dout, 0, (size_t)NNET * n * 4));
      for (int w = 0; w < 3; ++w) launch();
      CK(cudaGetLastError());
      CK(cudaEventRecord(e0));
      for (int it = 0; it < iters; ++it) launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      ms /= iters;
      double maxd = 0;
      for (int k = 0; k < NNET; ++k) {
        CK(cudaMemcpy(got.data() + (size_t)k * ns, dout + (size_t)k * n, ns * 4, cudaMemcpyDeviceToHost));
        for (int i = 0; i < ns; ++i) maxd = fmax(maxd, fabs((double)got[(size_t)k * ns + i] - ref[(size_t)k * ns + i]));
      }
      printf("%-6s CTAs/SM %d  %8.3f ms/launch  %.3e aircraft-net evals/s  (%.2f TMAC/s of useful MACs)  max|d| vs fp64 = %.2e of range\n",
             variant == 0 ? "ffma2" : "mma3", blocks_per_sm, ms, (double)NNET * n / (ms * 1e-3), (double)NNET * n * 250.0 / (ms * 1e-3) / 1e12,
             maxd / (hi - lo));
    }
  }
  return 0;
}
