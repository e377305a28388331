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

// Backward of the kernel above: g_out bf16 [B,L,H*64] (gradient of the normalised, un-interleaved rows) -> g_op bf16
// [B,H,L,64] at the source scan position, d gamma / d beta f32 [H*64] (atomic, pre-zeroed).  Same thread mapping; a CTA
// walks `tokens_per_cta` destination tokens (two at a time) and keeps its d gamma / d beta partial sums in registers, so
// the global atomics are one per feature per CTA.
__global__ void __launch_bounds__(256)
ttt_output_norm_bwd_kernel(const uint4* __restrict__ O, const float* __restrict__ gamma, const int* __restrict__ index,
                           const uint4* __restrict__ gout, uint4* __restrict__ gO, float* __restrict__ dgamma,
                           float* __restrict__ dbeta, int L, int H, float eps, int tokens_per_cta) {
  __shared__ float red[3][2][4];
  __shared__ float acc[2][64 * 64];  // d gamma / d beta of this CTA: the two token halves are folded here at the end
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tk = warp >> 2, w4 = warp & 3;
  const int b = blockIdx.y;
  const int hl = lane >> 3, piece = lane & 7;
  const float E = (float)(H * 64);
  float dg[kRounds][8], db[kRounds][8];
#pragma unroll
  for (int i = 0; i < kRounds; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) { dg[i][e] = 0.f; db[i][e] = 0.f; }
  for (int i = threadIdx.x; i < 2 * 64 * 64; i += 256) (&acc[0][0])[i] = 0.f;

  const int m_begin = blockIdx.x * tokens_per_cta;
  for (int it = 0; it < tokens_per_cta; it += 2) {
    const int mm = m_begin + it + tk;
    const bool valid = mm < L && it + tk < tokens_per_cta;
    const int m = min(mm, L - 1);
    const int src = index ? index[m] : m;
    float x[kRounds][8], g[kRounds][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kRounds; ++i) {
      const int h = 16 * i + 4 * w4 + hl;
#pragma unroll
      for (int e = 0; e < 8; ++e) { x[i][e] = 0.f; g[i][e] = 0.f; }
      if (h < H) {
        const uint4 v = O[(((size_t)b * H + h) * L + src) * 8 + piece];
        const uint4 u = gout[((size_t)b * L + m) * (H * 8) + h * 8 + piece];
        x[i][0] = bf16_lo(v.x); x[i][1] = bf16_hi(v.x); x[i][2] = bf16_lo(v.y); x[i][3] = bf16_hi(v.y);
        x[i][4] = bf16_lo(v.z); x[i][5] = bf16_hi(v.z); x[i][6] = bf16_lo(v.w); x[i][7] = bf16_hi(v.w);
        g[i][0] = bf16_lo(u.x); g[i][1] = bf16_hi(u.x); g[i][2] = bf16_lo(u.y); g[i][3] = bf16_hi(u.y);
        g[i][4] = bf16_lo(u.z); g[i][5] = bf16_hi(u.z); g[i][6] = bf16_lo(u.w); g[i][7] = bf16_hi(u.w);
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += x[i][e];
      }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    if (lane == 0) red[0][tk][w4] = sum;
    __syncthreads();
    const float mean = ((red[0][tk][0] + red[0][tk][1]) + (red[0][tk][2] + red[0][tk][3])) / E;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kRounds; ++i)
      if (16 * i + 4 * w4 + hl < H)
#pragma unroll
        for (int e = 0; e < 8; ++e) { x[i][e] -= mean; sq = fmaf(x[i][e], x[i][e], sq); }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, s);
    if (lane == 0) red[1][tk][w4] = sq;
    __syncthreads();
    const float rstd = rsqrtf(((red[1][tk][0] + red[1][tk][1]) + (red[1][tk][2] + red[1][tk][3])) / E + eps);
    // x <- x_hat ; g <- g_xhat = g_y * gamma ; row sums of g_xhat and g_xhat * x_hat
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kRounds; ++i) {
      const int h = 16 * i + 4 * w4 + hl;
      if (h < H) {
        const int f = h * 64 + 8 * piece;
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + f), g1 = *reinterpret_cast<const float4*>(gamma + f + 4);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          x[i][e] *= rstd;
          if (valid) { dg[i][e] = fmaf(g[i][e], x[i][e], dg[i][e]); db[i][e] += g[i][e]; }
          g[i][e] *= gm[e];
          s1 += g[i][e];
          s2 = fmaf(g[i][e], x[i][e], s2);
        }
      }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, s);
      s2 += __shfl_xor_sync(0xffffffffu, s2, s);
    }
    __syncthreads();  // red[0] of this round has been read by everyone
    if (lane == 0) { red[0][tk][w4] = s1; red[2][tk][w4] = s2; }
    __syncthreads();
    const float m1 = ((red[0][tk][0] + red[0][tk][1]) + (red[0][tk][2] + red[0][tk][3])) / E;
    const float m2 = ((red[2][tk][0] + red[2][tk][1]) + (red[2][tk][2] + red[2][tk][3])) / E;
    if (valid) {
#pragma unroll
      for (int i = 0; i < kRounds; ++i) {
        const int h = 16 * i + 4 * w4 + hl;
        if (h < H) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = rstd * (g[i][e] - m1 - x[i][e] * m2);
          gO[(((size_t)b * H + h) * L + src) * 8 + piece] =
              make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        }
      }
    }
    __syncthreads();  // red[*] reused by the next pair of tokens
  }
  // fold the CTA's d gamma / d beta: smem (two token halves share features), then one global atomic per feature
#pragma unroll
  for (int i = 0; i < kRounds; ++i) {
    const int h = 16 * i + 4 * w4 + hl;
    if (h < H)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&acc[0][h * 64 + 8 * piece + e], dg[i][e]);
        atomicAdd(&acc[1][h * 64 + 8 * piece + e], db[i][e]);
      }
  }
  __syncthreads();
  for (int f = threadIdx.x; f < H * 64; f += 256) {
    atomicAdd(dgamma + f, acc[0][f]);
    atomicAdd(dbeta + f, acc[1][f]);
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

cudaError_t launch_output_norm_backward(const void* O, const float* gamma, const int* index, const void* gout, void* gO,
                                        float* dgamma, float* dbeta, int B, int L, int H, float eps, cudaStream_t stream) {
  if (B <= 0 || L <= 0 || H <= 0 || H > 16 * kRounds) { g_where = "bad sizes (H must be <= 64)"; return cudaErrorInvalidValue; }
  TB_TRY(cudaMemsetAsync(dgamma, 0, (size_t)H * 64 * sizeof(float), stream), "memset dgamma");
  TB_TRY(cudaMemsetAsync(dbeta, 0, (size_t)H * 64 * sizeof(float), stream), "memset dbeta");
  const int tpc = 32;  // destination tokens per CTA
  g_where = "output norm backward launch";
  dim3 grid((L + tpc - 1) / tpc, B);
  ttt_output_norm_bwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(O), gamma, index,
                                                       reinterpret_cast<const uint4*>(gout), reinterpret_cast<uint4*>(gO),
                                                       dgamma, dbeta, L, H, eps, tpc);
  return cudaGetLastError();
}

}  // namespace tb
