// SASS experiment (no GPU needed): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin -o t.cubin elect_issue.cu;
// cuobjdump -sass t.cubin | grep -E 'Function|ELECT|UTCHMMA|PLOP3|BRA'
// kA (branch on threadIdx.x == 0): every UTCHMMA sits in its own election loop (ELECT / UTCHMMA / PLOP3 / PLOP3 / BRA.U.ANY) and the
// descriptors are rebuilt in between: 12 issue slots per MMA.  kB (warp-uniform branch + elect.sync): 8 UTCHMMA back to back.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../ttt-video-dit_b200/csrc/ptx.cuh"
using namespace tb;
#ifndef TB_HAS_ELECT_ONE  // the helper lives in ptx.cuh on the elect-issue branch
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
#endif
// variant A: divergent branch on tid (what the kernels do today)
__global__ void kA(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint64_t* bar) {
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) umma_ss(tmem, desc_advance(da, 32 * k), desc_advance(db, 32 * k), idesc, k > 0);
    tc_commit(bar);
  }
}
// variant B: warp-uniform branch + elect.sync
__global__ void kB(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint64_t* bar) {
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp == 0) {
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_ss(tmem, desc_advance(da, 32 * k), desc_advance(db, 32 * k), idesc, k > 0);
      tc_commit(bar);
    }
  }
}
