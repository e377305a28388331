// Prologue of the per-segment local attention (HBM-bound, one pass each way): per-head LayerNorm of q and k followed by the
// 3-D rotary embedding of the video tokens with segment-local positions.
// Reference: ttt/models/cogvideo/dit.py:188-194 -- cur_q = self.q_norm(cur_q); cur_k = self.k_norm(cur_k) (nn.LayerNorm(64,
// eps 1e-6) over the head dim), then self.rotary(x[:, :, text_length:]) = x * cos + rotate_half(x) * sin with interleaved
// pairs (cogvideo/utils.py:93-99,432-437), positions counted from the first video token of the segment.  The reference
// runs this as ~10 elementwise torch ops per tensor inside a torch.compile'd closure; here q and k are read once and written
// once ([B, T, H, 64] bf16, the layout of the q / k Linear outputs that the attention kernel consumes).
//   y = x_hat * gamma + beta ;  out[2i] = y[2i] c_i - y[2i+1] s_i ;  out[2i+1] = y[2i+1] c_i + y[2i] s_i      (video rows)
// Backward: d y from the transposed rotation, then the LayerNorm backward (x_hat recomputed from the saved input) and the
// parameter gradients (register partials per thread, folded through shared memory, one global atomic per feature and CTA).
// Thread mapping: 8 threads x 8 features = one (token, head) row; a warp = 4 rows; row statistics by 3 xor-shuffles.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"
#include "ttt_internal.h"

namespace tb {
namespace {

__device__ __forceinline__ void unpack8p(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8p(const float* o) {
  return make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
}
__device__ __forceinline__ float sum8(float v) {  // over the 8 lanes of a row
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ void load8f(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

}  // namespace

// grid-stride over rows (b, t, h); tensor index 0 = q, 1 = k selected by blockIdx.y; params [2][64]
__global__ void __launch_bounds__(256)
qk_norm_rope_fwd_kernel(const uint4* __restrict__ q, const uint4* __restrict__ k, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const float* __restrict__ cosT, const float* __restrict__ sinT,
                        uint4* __restrict__ qo, uint4* __restrict__ ko, long long rows, int T, int H, int text_len, float eps) {
  const int which = blockIdx.y;
  const uint4* x = which ? k : q;
  uint4* y = which ? ko : qo;
  const int piece = threadIdx.x & 7;
  float g[8], bt[8];
  load8f(gamma + which * 64 + 8 * piece, g);
  load8f(beta + which * 64 + 8 * piece, bt);
  // the row statistics use full-mask warp shuffles: every lane of a warp must run the same number of iterations, so the loop
  // bound is per CTA (32 rows at a time) and rows past the end are computed on a clamped index and not stored
  for (long long row0 = (long long)blockIdx.x * 32; row0 < rows; row0 += (long long)gridDim.x * 32) {
    const long long rr = row0 + (threadIdx.x >> 3);
    const bool valid = rr < rows;
    const long long row = valid ? rr : rows - 1;
    const int t = (int)((row / H) % T);
    float v[8];
    unpack8p(x[row * 8 + piece], v);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[e];
    const float mean = sum8(s) * (1.f / 64.f);
    float sq = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { v[e] -= mean; sq = fmaf(v[e], v[e], sq); }
    const float rstd = rsqrtf(sum8(sq) * (1.f / 64.f) + eps);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e] * rstd, g[e], bt[e]);
    if (t >= text_len) {
      float c[8], sn[8];
      load8f(cosT + (size_t)(t - text_len) * 64 + 8 * piece, c);
      load8f(sinT + (size_t)(t - text_len) * 64 + 8 * piece, sn);
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const float a = v[e], b = v[e + 1];
        v[e] = a * c[e] - b * sn[e];
        v[e + 1] = b * c[e + 1] + a * sn[e + 1];
      }
    }
    if (valid) y[row * 8 + piece] = pack8p(v);
  }
}

__global__ void __launch_bounds__(256)
qk_norm_rope_bwd_kernel(const uint4* __restrict__ q, const uint4* __restrict__ k, const float* __restrict__ gamma,
                        const float* __restrict__ cosT, const float* __restrict__ sinT, const uint4* __restrict__ dqo,
                        const uint4* __restrict__ dko, uint4* __restrict__ dq, uint4* __restrict__ dk, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, long long rows, int T, int H, int text_len, float eps) {
  __shared__ float acc[2][64];
  const int which = blockIdx.y;
  const uint4* x = which ? k : q;
  const uint4* dy = which ? dko : dqo;
  uint4* dx = which ? dk : dq;
  const int piece = threadIdx.x & 7;
  float g[8], dg[8], db[8];
  load8f(gamma + which * 64 + 8 * piece, g);
#pragma unroll
  for (int e = 0; e < 8; ++e) { dg[e] = 0.f; db[e] = 0.f; }
  if (threadIdx.x < 128) (&acc[0][0])[threadIdx.x] = 0.f;
  __syncthreads();
  for (long long row0 = (long long)blockIdx.x * 32; row0 < rows; row0 += (long long)gridDim.x * 32) {  // warp-uniform trip count
    const long long rr = row0 + (threadIdx.x >> 3);
    const bool valid = rr < rows;
    const long long row = valid ? rr : rows - 1;
    const int t = (int)((row / H) % T);
    float v[8], d[8];
    unpack8p(x[row * 8 + piece], v);
    unpack8p(dy[row * 8 + piece], d);
    if (!valid) {
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] = 0.f;  // contributes nothing to the parameter gradients
    }
    if (t >= text_len) {  // transposed rotation: d y from d out
      float c[8], sn[8];
      load8f(cosT + (size_t)(t - text_len) * 64 + 8 * piece, c);
      load8f(sinT + (size_t)(t - text_len) * 64 + 8 * piece, sn);
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const float a = d[e], b = d[e + 1];
        d[e] = a * c[e] + b * sn[e + 1];
        d[e + 1] = b * c[e + 1] - a * sn[e];
      }
    }
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[e];
    const float mean = sum8(s) * (1.f / 64.f);
    float sq = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { v[e] -= mean; sq = fmaf(v[e], v[e], sq); }
    const float rstd = rsqrtf(sum8(sq) * (1.f / 64.f) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] *= rstd;  // x_hat
      dg[e] = fmaf(d[e], v[e], dg[e]);
      db[e] += d[e];
      d[e] *= g[e];  // d x_hat
      s1 += d[e];
      s2 = fmaf(d[e], v[e], s2);
    }
    const float m1 = sum8(s1) * (1.f / 64.f), m2 = sum8(s2) * (1.f / 64.f);
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] = rstd * (d[e] - m1 - v[e] * m2);
    if (valid) dx[row * 8 + piece] = pack8p(d);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    atomicAdd(&acc[0][8 * piece + e], dg[e]);
    atomicAdd(&acc[1][8 * piece + e], db[e]);
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    atomicAdd(dgamma + which * 64 + threadIdx.x, acc[0][threadIdx.x]);
    atomicAdd(dbeta + which * 64 + threadIdx.x, acc[1][threadIdx.x]);
  }
}

static bool prologue_args_ok(int B, int T, int H, int text_len) { return B > 0 && T > 0 && H > 0 && text_len >= 0 && text_len <= T; }

cudaError_t launch_qk_norm_rope(const void* q, const void* k, const float* gamma, const float* beta, const float* cosT,
                                const float* sinT, void* q_out, void* k_out, int B, int T, int H, int text_len, float eps,
                                cudaStream_t stream) {
  if (!prologue_args_ok(B, T, H, text_len)) { g_where = "bad sizes"; return cudaErrorInvalidValue; }
  g_where = "qk_norm_rope launch";
  const long long rows = (long long)B * T * H;
  long long blocks = (rows + 31) / 32;
  if (blocks > 148 * 8) blocks = 148 * 8;
  dim3 grid((unsigned)blocks, 2);
  qk_norm_rope_fwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(q), reinterpret_cast<const uint4*>(k), gamma, beta,
                                                    cosT, sinT, reinterpret_cast<uint4*>(q_out), reinterpret_cast<uint4*>(k_out), rows, T,
                                                    H, text_len, eps);
  return cudaGetLastError();
}

cudaError_t launch_qk_norm_rope_backward(const void* q, const void* k, const float* gamma, const float* cosT, const float* sinT,
                                         const void* dq_out, const void* dk_out, void* dq, void* dk, float* dgamma, float* dbeta,
                                         int B, int T, int H, int text_len, float eps, cudaStream_t stream) {
  if (!prologue_args_ok(B, T, H, text_len)) { g_where = "bad sizes"; return cudaErrorInvalidValue; }
  TB_TRY(cudaMemsetAsync(dgamma, 0, 128 * sizeof(float), stream), "memset dgamma");
  TB_TRY(cudaMemsetAsync(dbeta, 0, 128 * sizeof(float), stream), "memset dbeta");
  g_where = "qk_norm_rope backward launch";
  const long long rows = (long long)B * T * H;
  long long blocks = (rows + 31) / 32;
  if (blocks > 148 * 4) blocks = 148 * 4;
  dim3 grid((unsigned)blocks, 2);
  qk_norm_rope_bwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(q), reinterpret_cast<const uint4*>(k), gamma, cosT,
                                                    sinT, reinterpret_cast<const uint4*>(dq_out), reinterpret_cast<const uint4*>(dk_out),
                                                    reinterpret_cast<uint4*>(dq), reinterpret_cast<uint4*>(dk), dgamma, dbeta, rows, T, H,
                                                    text_len, eps);
  return cudaGetLastError();
}

}  // namespace tb
