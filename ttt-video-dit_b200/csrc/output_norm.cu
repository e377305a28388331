// Output side of the TTT layer before the wo Linear, one HBM pass: the op output [B,H,NC,CS,64] is
// transposed to [B,L,H*64] (TTTMLP.ttt, ttt/models/ssm/ttt_layer.py:456,472), normalised by post_norm = LayerNorm(H*64,
// eps 1e-6) (ttt_layer.py:71,324) and put back into the caller's token order (undo_interleave, ttt_layer.py:191-217,
// :329-331 -- a token permutation, which commutes with the per-token wo Linear and is therefore applied here, before wo).
// The reference does this as permute + reshape copy, LayerNorm, and chunk/cat copies.
//
// One CTA = two destination tokens (4 warps each): a thread reads 16-byte pieces (8 features; 8 lanes cover the 128-byte
// row of one head, a warp 4 heads, the 4 warps 16 heads per round), two block reductions (mean, then centred sum of
// squares -- the row stays in registers), 16-byte stores.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"
#include "ttt_internal.h"

namespace tb {

constexpr int kRounds = 4;  // H <= 64

__global__ void __launch_bounds__(256)
ttt_output_norm_kernel(const uint4* __restrict__ O, const float* __restrict__ gamma, const float* __restrict__ beta,
                       const int* __restrict__ index, uint4* __restrict__ out, int L, int H, float eps) {
  __shared__ float red[2][2][4];  // [pass][token of the CTA][warp]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tk = warp >> 2, w4 = warp & 3;            // which of the CTA's two tokens, warp within the token
  const int m = min(blockIdx.x * 2 + tk, L - 1), b = blockIdx.y;
  const bool store = blockIdx.x * 2 + tk < L;
  const int src = index ? index[m] : m;
  const int hl = lane >> 3, piece = lane & 7;          // head within the warp's group of 4, 16-byte piece of its row
  float x[kRounds][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kRounds; ++i) {
    const int h = 16 * i + 4 * w4 + hl;
#pragma unroll
    for (int e = 0; e < 8; ++e) x[i][e] = 0.f;
    if (h < H) {
      const uint4 v = O[(((size_t)b * H + h) * L + src) * 8 + piece];
      x[i][0] = bf16_lo(v.x); x[i][1] = bf16_hi(v.x); x[i][2] = bf16_lo(v.y); x[i][3] = bf16_hi(v.y);
      x[i][4] = bf16_lo(v.z); x[i][5] = bf16_hi(v.z); x[i][6] = bf16_lo(v.w); x[i][7] = bf16_hi(v.w);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += x[i][e];
    }
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  if (lane == 0) red[0][tk][w4] = sum;
  __syncthreads();
  const float E = (float)(H * 64);
  const float mean = ((red[0][tk][0] + red[0][tk][1]) + (red[0][tk][2] + red[0][tk][3])) / E;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kRounds; ++i) {
    if (16 * i + 4 * w4 + hl < H) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { x[i][e] -= mean; sq = fmaf(x[i][e], x[i][e], sq); }
    }
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, s);
  if (lane == 0) red[1][tk][w4] = sq;
  __syncthreads();
  const float rstd = rsqrtf(((red[1][tk][0] + red[1][tk][1]) + (red[1][tk][2] + red[1][tk][3])) / E + eps);
  if (!store) return;
#pragma unroll
  for (int i = 0; i < kRounds; ++i) {
    const int h = 16 * i + 4 * w4 + hl;
    if (h < H) {
      const int f = h * 64 + 8 * piece;
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + f), g1 = *reinterpret_cast<const float4*>(gamma + f + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + f), b1 = *reinterpret_cast<const float4*>(beta + f + 4);
      out[((size_t)b * L + m) * (H * 8) + h * 8 + piece] =
          make_uint4(pack_bf16(fmaf(x[i][0] * rstd, g0.x, b0.x), fmaf(x[i][1] * rstd, g0.y, b0.y)),
                     pack_bf16(fmaf(x[i][2] * rstd, g0.z, b0.z), fmaf(x[i][3] * rstd, g0.w, b0.w)),
                     pack_bf16(fmaf(x[i][4] * rstd, g1.x, b1.x), fmaf(x[i][5] * rstd, g1.y, b1.y)),
                     pack_bf16(fmaf(x[i][6] * rstd, g1.z, b1.z), fmaf(x[i][7] * rstd, g1.w, b1.w)));
    }
  }
}

cudaError_t launch_output_norm(const void* O, const float* gamma, const float* beta, const int* index, void* out, int B, int L,
                               int H, float eps, cudaStream_t stream) {
  if (B <= 0 || L <= 0 || H <= 0 || H > 16 * kRounds) { g_where = "bad sizes (H must be <= 64)"; return cudaErrorInvalidValue; }
  g_where = "output norm launch";
  dim3 grid((L + 1) / 2, B);
  ttt_output_norm_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(O), gamma, beta, index,
                                                   reinterpret_cast<uint4*>(out), L, H, eps);
  return cudaGetLastError();
}

}  // namespace tb
