/* mmq_device.cuh — small device helpers shared by the allocation kernels. */
#ifndef MMQ_DEVICE_CUH
#define MMQ_DEVICE_CUH

#include <stdint.h>

#include "../../include/mmq_sampler.h"

/* counts[c] += 1 for every lane with c >= 0, one reduction per distinct column in the warp */
__device__ __forceinline__ void cat_red(int32_t* __restrict__ counts, int32_t c, int lane) {
  const unsigned act = __ballot_sync(0xffffffffu, c >= 0);
  if (c >= 0) {
    const unsigned grp = __match_any_sync(act, c);
    if (lane == __ffs(grp) - 1) atomicAdd(counts + c, __popc(grp));
  }
}

/* Position of this lane's item in a shared-memory queue (counter in shared memory), -1 when pred is false:
 * one atomic per warp.  Every lane of the warp must call it. */
__device__ __forceinline__ int queue_slot(int* counter, bool pred, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m == 0u) return -1;
  const int leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return pred ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}

/* ---- mbarrier + TMA 1-D bulk copy (SASS: SYNCS.*, UBLKCP) ---- */
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
/* TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP) */
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
/* the same with an L2 evict-first policy: for streams that are read once per sweep and are larger than the L2 */
__device__ __forceinline__ void bulk_g2s_stream(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile(
      "{\n\t"
      ".reg .b64 pol;\n\t"
      "createpolicy.fractional.L2::evict_first.b64 pol, 1.0;\n\t"
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], pol;\n\t"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(b))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(smem_u32(b)), "r"(parity)
                 : "memory");
  } while (!ok);
}


#endif
