// tma_utils.cuh -- minimal TMA (cp.async.bulk.tensor) + mbarrier wrappers for the tile loaders (sm_100a).
//
// The source image is described once per launch by a CUtensorMap (2-D, one 32-bit element per RGBA8 texel);
// one elected thread per CTA issues a single bulk-tensor copy of the (tile + halo) box into shared memory and
// the CTA waits on an mbarrier.  Out-of-bounds texels arrive as zeros, which IS Texture2D.Load's behaviour
// (RCAS, fsr_rcas.hlsl:18); EASU re-maps border texels to clamp-to-edge while decoding.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ovrfsr {
inline namespace OVRFSR_MODE_NS {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barrier visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 2-D tiled bulk tensor load global -> shared, completion signalled on `bar` (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
// order this thread's prior generic-proxy accesses to shared memory before subsequent async-proxy (TMA) ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// ---- cluster launch control (sm_100): a running CTA cancels a not-yet-launched CTA of its own grid and takes over its
// block index -- hardware work stealing, i.e. a persistent kernel whose work list is balanced dynamically without a
// global counter.  try_cancel is asynchronous: the 16-byte response lands in shared memory and completes `bar`.
__device__ __forceinline__ void clc_try_cancel(void *response16, uint64_t *bar) {
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(
                   smem_u32(response16)),
               "r"(smem_u32(bar))
               : "memory");
}
// blockIdx.x of the cancelled CTA, or -1 when nothing was left to cancel (no further try_cancel may be issued then)
__device__ __forceinline__ int clc_cancelled_block_x(const void *response16) {
  uint32_t x, valid;
  asm volatile(
      "{\n"
      ".reg .pred p1;\n"
      ".reg .b128 r;\n"
      "ld.shared.b128 r, [%2];\n"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n"
      "selp.u32 %1, 1, 0, p1;\n"
      "mov.u32 %0, 0;\n"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid::x.b32.b128 %0, r;\n"
      "}\n"
      : "=r"(x), "=r"(valid)
      : "r"(smem_u32(response16))
      : "memory");
  return valid ? (int)x : -1;
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

} // inline namespace OVRFSR_MODE_NS
} // namespace ovrfsr
