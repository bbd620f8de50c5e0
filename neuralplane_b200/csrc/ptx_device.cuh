// PTX wrappers used by every kernel of libnplane.so: mbarrier, TMA 1-D bulk copies (global <-> shared), the async-proxy fence,
// and the once-per-CTA staging of a constant image (aero nets / tables) into shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"   // %3: suspend-time hint
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// A waiting warp backs off with nanosleep between polls: a tight try_wait loop was 21 % of the UAV slab kernel's executed
// instructions -- issue slots taken from the warps that had work.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  do {
    __nanosleep(100);
  } while (!mbar_try_wait(bar, parity));
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Stage the aero image into shared memory once per CTA: one elected thread issues TMA bulk copies that
// complete on an mbarrier; everyone waits on it.
// stage_aero_issue() only starts the copies (the caller overlaps its first global loads with them and waits with
// mbar_wait(bar, 0) before the first use of the image).
__device__ __forceinline__ void stage_aero_issue(void* blob_s, const void* aero_g, uint32_t bytes, uint64_t* bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, bytes);
    constexpr uint32_t kChunk = 16384;
    for (uint32_t off = 0; off < bytes; off += kChunk) {
      const uint32_t nb = min(kChunk, bytes - off);
      bulk_g2s(reinterpret_cast<char*>(blob_s) + off, reinterpret_cast<const char*>(aero_g) + off, nb, bar);
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void stage_aero(void* blob_s, const void* aero_g, uint32_t bytes, uint64_t* bar) {
  stage_aero_issue(blob_s, aero_g, bytes, bar);
  mbar_wait(bar, 0);
}

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// before the kernel ahead of it in the stream has finished; grid_dependency_wait() blocks until that kernel has completed and
// its writes are visible (a no-op for an ordinary launch), grid_launch_dependents() lets the NEXT kernel of the stream start.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
