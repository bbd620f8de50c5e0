// Micro-benchmark: how should warp-uniform MLP weights reach the FMA pipe on sm_100a?
//
// Workload = the 36 (alpha,beta)-only nets of the F-16 aero model (10 305 MACs per aircraft), random weights.
// Variants (weights path x aircraft per thread):
//   lds1   weights LDS.128 broadcast from shared memory, 1 aircraft/thread (round-1 production path)
//   lds2   same, 2 aircraft/thread (each weight register feeds 2 FFMA)
//   lds2p  same, 2 aircraft/thread packed: FFMA2 acc.xy += w(scalar) * h.xy
//   lds4p  4 aircraft/thread, two FFMA2 per weight
//   ur1    weights in __constant__ memory, uniform loads (LDCU) -> FFMA with UR operand, 1 aircraft/thread
//   ur2    same, 2 aircraft/thread
//   ur2p   same, packed FFMA2
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mlp_bench tools/mlp_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int pad4(int x) { return (x + 3) & ~3; }
constexpr int layer_floats(int in, int out) { return pad4(out + in * out); }
struct Arch { int nin, h1, h2, h3, count; };
// the nine (alpha,beta) groups: arch + number of nets
constexpr Arch kGroups[] = {{2, 20, 10, 0, 2}, {2, 20, 10, 0, 2}, {2, 20, 10, 5, 4}, {2, 20, 10, 5, 4}, {2, 20, 10, 10, 1},
                            {2, 20, 20, 10, 3}, {1, 20, 10, 0, 12}, {1, 20, 10, 0, 7}, {1, 20, 10, 5, 1}};
constexpr int kNumGroups = 9;
constexpr int net_floats(Arch a) {
  return layer_floats(a.nin, a.h1) + layer_floats(a.h1, a.h2) +
         (a.h3 ? layer_floats(a.h2, a.h3) + layer_floats(a.h3, 1) : layer_floats(a.h2, 1));
}
constexpr int net_macs(Arch a) {
  return a.nin * a.h1 + a.h1 * a.h2 + (a.h3 ? a.h2 * a.h3 + a.h3 : a.h2);
}
constexpr int group_offset(int g) {
  int off = 0;
  for (int i = 0; i < g; ++i) off += net_floats(kGroups[i]) * kGroups[i].count;
  return off;
}
constexpr int kBlobFloats = group_offset(kNumGroups);
constexpr int total_macs() {
  int m = 0;
  for (int i = 0; i < kNumGroups; ++i) m += net_macs(kGroups[i]) * kGroups[i].count;
  return m;
}
constexpr int kMacs = total_macs();

__constant__ float c_blob[kBlobFloats];

// ---- weight sources ------------------------------------------------------------------------------
struct SrcLds {
  uint32_t base;  // shared address of the net
  __device__ __forceinline__ float4 ld4(int off_floats) const {
    float4 q;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(base + 4 * off_floats));
    return q;
  }
  __device__ __forceinline__ SrcLds advance(int floats) const { return SrcLds{base + 4 * (uint32_t)floats}; }
};
struct SrcConst {
  int base;  // float index into c_blob (warp-uniform)
  __device__ __forceinline__ float4 ld4(int off_floats) const {
    return *reinterpret_cast<const float4*>(&c_blob[base + off_floats]);
  }
  __device__ __forceinline__ SrcConst advance(int floats) const { return SrcConst{base + floats}; }
};

// ---- dense layers --------------------------------------------------------------------------------
// scalar FFMA, K aircraft per thread
template <int K, int IN, int OUT, bool RELU, class S>
__device__ __forceinline__ void dense(S w, const float (&x)[K][IN], float (&y)[K][OUT]) {
  constexpr int NF = OUT + IN * OUT, NV = (NF + 3) / 4;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 q = w.ld4(4 * v);
    const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int f = 4 * v + k;
      if (f < OUT) {
#pragma unroll
        for (int a = 0; a < K; ++a) y[a][f] = e[k];
      } else if (f < NF) {
        const int g = f - OUT;
#pragma unroll
        for (int a = 0; a < K; ++a) y[a][g % OUT] = fmaf(x[a][g / OUT], e[k], y[a][g % OUT]);
      }
    }
  }
  if (RELU) {
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
      for (int j = 0; j < OUT; ++j) y[a][j] = fmaxf(y[a][j], 0.0f);
  }
}
// packed FFMA2, KP aircraft PAIRS per thread: acc.xy += w * h.xy
template <int KP, int IN, int OUT, bool RELU, class S>
__device__ __forceinline__ void dense_p(S w, const float2 (&x)[KP][IN], float2 (&y)[KP][OUT]) {
  constexpr int NF = OUT + IN * OUT, NV = (NF + 3) / 4;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 q = w.ld4(4 * v);
    const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int f = 4 * v + k;
      if (f < OUT) {
#pragma unroll
        for (int a = 0; a < KP; ++a) y[a][f] = make_float2(e[k], e[k]);
      } else if (f < NF) {
        const int g = f - OUT;
#pragma unroll
        for (int a = 0; a < KP; ++a) y[a][g % OUT] = __ffma2_rn(make_float2(e[k], e[k]), x[a][g / OUT], y[a][g % OUT]);
      }
    }
  }
  if (RELU) {
#pragma unroll
    for (int a = 0; a < KP; ++a)
#pragma unroll
      for (int j = 0; j < OUT; ++j) y[a][j] = make_float2(fmaxf(y[a][j].x, 0.0f), fmaxf(y[a][j].y, 0.0f));
  }
}

template <int K, int NIN, int H1, int H2, int H3, class S>
__device__ __forceinline__ void mlp(S w, const float (&z)[K][2], float (&out)[K]) {
  float x[K][NIN];
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int j = 0; j < NIN; ++j) x[a][j] = z[a][j];
  float h1[K][H1];
  dense<K, NIN, H1, true>(w, x, h1);
  w = w.advance(layer_floats(NIN, H1));
  float h2[K][H2];
  dense<K, H1, H2, true>(w, h1, h2);
  w = w.advance(layer_floats(H1, H2));
  float y[K][1];
  if constexpr (H3 > 0) {
    float h3[K][H3];
    dense<K, H2, H3, true>(w, h2, h3);
    w = w.advance(layer_floats(H2, H3));
    dense<K, H3, 1, false>(w, h3, y);
  } else {
    dense<K, H2, 1, false>(w, h2, y);
  }
#pragma unroll
  for (int a = 0; a < K; ++a) out[a] = y[a][0];
}
template <int KP, int NIN, int H1, int H2, int H3, class S>
__device__ __forceinline__ void mlp_p(S w, const float2 (&z)[KP][2], float2 (&out)[KP]) {
  float2 x[KP][NIN];
#pragma unroll
  for (int a = 0; a < KP; ++a)
#pragma unroll
    for (int j = 0; j < NIN; ++j) x[a][j] = z[a][j];
  float2 h1[KP][H1];
  dense_p<KP, NIN, H1, true>(w, x, h1);
  w = w.advance(layer_floats(NIN, H1));
  float2 h2[KP][H2];
  dense_p<KP, H1, H2, true>(w, h1, h2);
  w = w.advance(layer_floats(H1, H2));
  float2 y[KP][1];
  if constexpr (H3 > 0) {
    float2 h3[KP][H3];
    dense_p<KP, H2, H3, true>(w, h2, h3);
    w = w.advance(layer_floats(H2, H3));
    dense_p<KP, H3, 1, false>(w, h3, y);
  } else {
    dense_p<KP, H2, 1, false>(w, h2, y);
  }
#pragma unroll
  for (int a = 0; a < KP; ++a) out[a] = y[a][0];
}

template <int G, int K, bool PACKED, class S>
__device__ __forceinline__ void eval_group(S base, const float (&z)[K][2], float (&acc)[K]) {
  constexpr Arch A = kGroups[G];
  S w = base.advance(group_offset(G));
#pragma unroll 1
  for (int k = 0; k < A.count; ++k, w = w.advance(net_floats(A))) {
    if constexpr (!PACKED) {
      float o[K];
      mlp<K, A.nin, A.h1, A.h2, A.h3>(w, z, o);
#pragma unroll
      for (int a = 0; a < K; ++a) acc[a] += o[a];
    } else {
      constexpr int KP = K / 2;
      float2 zp[KP][2], o[KP];
#pragma unroll
      for (int a = 0; a < KP; ++a)
#pragma unroll
        for (int j = 0; j < 2; ++j) zp[a][j] = make_float2(z[2 * a][j], z[2 * a + 1][j]);
      mlp_p<KP, A.nin, A.h1, A.h2, A.h3>(w, zp, o);
#pragma unroll
      for (int a = 0; a < KP; ++a) { acc[2 * a] += o[a].x; acc[2 * a + 1] += o[a].y; }
    }
  }
}

template <int K, bool PACKED, bool CONST, int BS, int MINB>
__global__ void __launch_bounds__(BS, MINB) bench_kernel(const float* __restrict__ blob_g, const float* __restrict__ in,
                                                         float* __restrict__ out, int n) {
  extern __shared__ __align__(16) float sm[];
  if (!CONST) {
    for (int j = threadIdx.x; j < kBlobFloats; j += BS) sm[j] = blob_g[j];
    __syncthreads();
  }
  for (int base = (blockIdx.x * BS + threadIdx.x) * K; base < n; base += gridDim.x * BS * K) {
    float z[K][2], acc[K];
#pragma unroll
    for (int a = 0; a < K; ++a) { z[a][0] = in[base + a]; z[a][1] = in[n + base + a]; acc[a] = 0.f; }
    if constexpr (CONST) {
      int b0;
      asm volatile("mov.u32 %0, 0;" : "=r"(b0));
      SrcConst s{b0};
      eval_group<0, K, PACKED>(s, z, acc); eval_group<1, K, PACKED>(s, z, acc); eval_group<2, K, PACKED>(s, z, acc);
      eval_group<3, K, PACKED>(s, z, acc); eval_group<4, K, PACKED>(s, z, acc); eval_group<5, K, PACKED>(s, z, acc);
      eval_group<6, K, PACKED>(s, z, acc); eval_group<7, K, PACKED>(s, z, acc); eval_group<8, K, PACKED>(s, z, acc);
    } else {
      uint32_t b0;
      asm volatile("mov.u32 %0, %1;" : "=r"(b0) : "r"((uint32_t)__cvta_generic_to_shared(sm)));
      SrcLds s{b0};
      eval_group<0, K, PACKED>(s, z, acc); eval_group<1, K, PACKED>(s, z, acc); eval_group<2, K, PACKED>(s, z, acc);
      eval_group<3, K, PACKED>(s, z, acc); eval_group<4, K, PACKED>(s, z, acc); eval_group<5, K, PACKED>(s, z, acc);
      eval_group<6, K, PACKED>(s, z, acc); eval_group<7, K, PACKED>(s, z, acc); eval_group<8, K, PACKED>(s, z, acc);
    }
#pragma unroll
    for (int a = 0; a < K; ++a) out[base + a] = acc[a];
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int K, bool PACKED, bool CONST, int BS, int MINB>
void run(const char* name, const float* blob_d, const float* in_d, float* out_d, int n, int sms, std::vector<float>& ref) {
  auto kern = bench_kernel<K, PACKED, CONST, BS, MINB>;
  const int smem = CONST ? 0 : kBlobFloats * 4;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  const int grid = sms * MINB;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) kern<<<grid, BS, smem>>>(blob_d, in_d, out_d, n);
  CK(cudaDeviceSynchronize());
  const int reps = 10;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) kern<<<grid, BS, smem>>>(blob_d, in_d, out_d, n);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<float> got(n);
  CK(cudaMemcpy(got.data(), out_d, n * 4, cudaMemcpyDeviceToHost));
  double maxd = 0;
  if (ref.empty()) ref = got;
  for (int i = 0; i < n; ++i) maxd = fmax(maxd, fabs((double)got[i] - ref[i]));
  const double macs = (double)kMacs * n * reps / (ms * 1e-3);
  printf("%-8s K=%d packed=%d const=%d BS=%d minb=%d regs=%3d  %8.3f ms/launch  %7.2f TMAC/s  (%.1f%% of 37.2)  %.3e aircraft-evals/s  maxdiff=%.2e\n",
         name, K, (int)PACKED, (int)CONST, BS, MINB, fa.numRegs, ms / reps, macs / 1e12, 100.0 * macs / 37.2e12,
         (double)n * reps / (ms * 1e-3), maxd);
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int n = 1 << 20;
  std::vector<float> blob(kBlobFloats), in(2 * n);
  srand(1);
  for (auto& w : blob) w = (rand() / (float)RAND_MAX - 0.5f) * 0.6f;
  for (auto& x : in) x = (rand() / (float)RAND_MAX - 0.5f) * 3.0f;
  float *blob_d, *in_d, *out_d;
  CK(cudaMalloc(&blob_d, kBlobFloats * 4)); CK(cudaMalloc(&in_d, 2 * n * 4)); CK(cudaMalloc(&out_d, n * 4));
  CK(cudaMemcpy(blob_d, blob.data(), kBlobFloats * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(in_d, in.data(), 2 * n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpyToSymbol(c_blob, blob.data(), kBlobFloats * 4));
  printf("SMs=%d blob=%d floats (%d B) macs/aircraft=%d\n", sms, kBlobFloats, kBlobFloats * 4, kMacs);
  std::vector<float> ref;
  run<1, false, false, 256, 1>("lds1", blob_d, in_d, out_d, n, sms, ref);
  run<1, false, false, 512, 1>("lds1", blob_d, in_d, out_d, n, sms, ref);
  run<1, false, false, 256, 3>("lds1", blob_d, in_d, out_d, n, sms, ref);
  run<2, false, false, 256, 1>("lds2", blob_d, in_d, out_d, n, sms, ref);
  run<2, false, false, 256, 2>("lds2", blob_d, in_d, out_d, n, sms, ref);
  run<2, true, false, 256, 1>("lds2p", blob_d, in_d, out_d, n, sms, ref);
  run<2, true, false, 256, 2>("lds2p", blob_d, in_d, out_d, n, sms, ref);
  run<2, true, false, 256, 3>("lds2p", blob_d, in_d, out_d, n, sms, ref);
  run<4, true, false, 256, 1>("lds4p", blob_d, in_d, out_d, n, sms, ref);
  run<4, true, false, 256, 2>("lds4p", blob_d, in_d, out_d, n, sms, ref);
  run<4, false, false, 256, 1>("lds4", blob_d, in_d, out_d, n, sms, ref);
  run<1, false, true, 256, 1>("ur1", blob_d, in_d, out_d, n, sms, ref);
  run<1, false, true, 256, 2>("ur1", blob_d, in_d, out_d, n, sms, ref);
  run<1, false, true, 256, 4>("ur1", blob_d, in_d, out_d, n, sms, ref);
  run<1, false, true, 128, 8>("ur1", blob_d, in_d, out_d, n, sms, ref);
  run<2, false, true, 256, 1>("ur2", blob_d, in_d, out_d, n, sms, ref);
  run<2, false, true, 256, 2>("ur2", blob_d, in_d, out_d, n, sms, ref);
  run<2, false, true, 256, 3>("ur2", blob_d, in_d, out_d, n, sms, ref);
  run<2, true, true, 256, 2>("ur2p", blob_d, in_d, out_d, n, sms, ref);
  run<4, true, true, 256, 1>("ur4p", blob_d, in_d, out_d, n, sms, ref);
  run<4, true, true, 256, 2>("ur4p", blob_d, in_d, out_d, n, sms, ref);
  return 0;
}
