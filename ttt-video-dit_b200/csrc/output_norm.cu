// Output side of the TTT layer before the wo Linear, one HBM pass: the op output [B,H,NC,CS,64] is
// transposed to [B,L,H*64] (TTTMLP.ttt, ttt/models/ssm/ttt_layer.py:456,472), normalised by post_norm = LayerNorm(H*64,
// eps 1e-6) (ttt_layer.py:71,324) and put back into the caller's token order (undo_interleave, ttt_layer.py:191-217,
// :329-331 -- a token permutation, which commutes with the per-token wo Linear and is therefore applied here, before wo).
// The reference does this as permute + reshape copy, LayerNorm, and chunk/cat copies.
//
// One CTA per destination token: warp w reads the 128-byte rows of heads w, w+8, ... of the source scan position (lane =
// 2 features), two block reductions (mean, then centred sum of squares -- the row stays in registers), 128-byte stores.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"
#include "ttt_internal.h"

namespace tb {

constexpr int kMaxHeadsPerWarp = 8;  // H <= 64

__global__ void __launch_bounds__(256)
ttt_output_norm_kernel(const uint32_t* __restrict__ O, const float* __restrict__ gamma, const float* __restrict__ beta,
                       const int* __restrict__ index, uint32_t* __restrict__ out, int L, int H, float eps) {
  __shared__ float red[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x, b = blockIdx.y;
  const int src = index ? index[m] : m;
  float x0[kMaxHeadsPerWarp], x1[kMaxHeadsPerWarp];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxHeadsPerWarp; ++i) {
    const int h = warp + 8 * i;
    x0[i] = 0.f; x1[i] = 0.f;
    if (h < H) {
      const uint32_t v = O[(((size_t)b * H + h) * L + src) * 32 + lane];
      x0[i] = bf16_lo(v); x1[i] = bf16_hi(v);
      sum += x0[i] + x1[i];
    }
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  if (lane == 0) red[0][warp] = sum;
  __syncthreads();
  const float E = (float)(H * 64);
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[0][w];
  const float mean = tot / E;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxHeadsPerWarp; ++i) {
    if (warp + 8 * i < H) {
      x0[i] -= mean; x1[i] -= mean;
      sq = fmaf(x0[i], x0[i], fmaf(x1[i], x1[i], sq));
    }
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, s);
  if (lane == 0) red[1][warp] = sq;
  __syncthreads();
  tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[1][w];
  const float rstd = rsqrtf(tot / E + eps);
#pragma unroll
  for (int i = 0; i < kMaxHeadsPerWarp; ++i) {
    const int h = warp + 8 * i;
    if (h < H) {
      const int f = h * 64 + 2 * lane;
      out[((size_t)b * L + m) * (H * 32) + h * 32 + lane] =
          pack_bf16(fmaf(x0[i] * rstd, gamma[f], beta[f]), fmaf(x1[i] * rstd, gamma[f + 1], beta[f + 1]));
    }
  }
}

cudaError_t launch_output_norm(const void* O, const float* gamma, const float* beta, const int* index, void* out, int B, int L,
                               int H, float eps, cudaStream_t stream) {
  if (B <= 0 || L <= 0 || H <= 0 || H > 8 * kMaxHeadsPerWarp) { g_where = "bad sizes (H must be <= 64)"; return cudaErrorInvalidValue; }
  g_where = "output norm launch";
  dim3 grid(L, B);
  ttt_output_norm_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(O), gamma, beta, index,
                                                   reinterpret_cast<uint32_t*>(out), L, H, eps);
  return cudaGetLastError();
}

}  // namespace tb
