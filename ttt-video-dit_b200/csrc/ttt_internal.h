// Internal (non-ABI) declarations shared by the .cu files of libttt_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tb {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_token_tmap(CUtensorMap* tm, const void* base, uint64_t rows);

cudaError_t launch_mlp_forward(const void* XQ, const void* XK, const void* XV, const void* last_eta, const float* ln_w,
                               const float* ln_b, const float* W1, const float* b1, const float* W2, const float* b2,
                               float* W1c, float* b1c, float* W2c, float* b2c, float* W1o, float* b1o, float* W2o,
                               float* b2o, void* Out, int B, int H, int NC, int ckpt_group, cudaStream_t stream);

cudaError_t launch_umma_selftest(int mode, const void* A, const void* Bm, float* D, int N, int K, cudaStream_t stream);

}  // namespace tb
