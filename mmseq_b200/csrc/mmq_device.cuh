/* mmq_device.cuh — small device helpers shared by the allocation kernels. */
#ifndef MMQ_DEVICE_CUH
#define MMQ_DEVICE_CUH

#include <stdint.h>

#include "../../include/mmq_sampler.h"

/* (j + 1/2) 2^-52 for the 52-bit integer j = (hi:lo) >> 12 — the value of mmq_uniform —
 * built without an int->double conversion: [1,2) mantissa trick, both steps exact. */
__device__ __forceinline__ double cat_u52(uint32_t hi, uint32_t lo) {
  const uint64_t j = (((uint64_t)hi << 32) | (uint64_t)lo) >> 12;
  return (__longlong_as_double((long long)(0x3ff0000000000000ull | j)) - 1.0) + 1.1102230246251565404e-16; /* + 2^-53 */
}

/* counts[c] += 1 for every lane with c >= 0, one reduction per distinct column in the warp */
__device__ __forceinline__ void cat_red(int32_t* __restrict__ counts, int32_t c, int lane) {
  const unsigned act = __ballot_sync(0xffffffffu, c >= 0);
  if (c >= 0) {
    const unsigned grp = __match_any_sync(act, c);
    if (lane == __ffs(grp) - 1) atomicAdd(counts + c, __popc(grp));
  }
}

#endif
