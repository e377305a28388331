// adaLN shell of the DiT TransformerLayer around the hot path, HBM-bound, one pass each way.
// Reference: ttt/models/cogvideo/dit.py:321-382 -- pre_seq_layernorm / pre_mlp_layernorm (nn.LayerNorm(model_dim, eps)) +
// modulate(x, shift, scale) = x * (1 + scale) + shift with separate (shift, scale) for text and video tokens, and the gated
// residuals  emb + gate * block_out  (again one gate vector for text, one for video, per batch element).
//
//   ln_affine:  out[b,l,:] = x_hat[b,l,:] * A[b,s(l),:] + C[b,s(l),:]       s(l) = l < text_len ? 0 (text) : 1 (video)
//               with A = gamma * (1 + scale), C = beta * (1 + scale) + shift folded by the caller ([B,2,E] fp32, a few KB);
//   gate_add:   out[b,l,:] = x[b,l,:] + G[b,s(l),:] * y[b,l,:]
// and their backward passes (d x, d A, d C / d x, d y, d G; the parameter gradients are summed over the rows of each
// (batch, segment) in registers, folded through shared memory, one global atomic per feature and CTA).
// Layout: [B, L, E] bf16, text tokens first; E a multiple of 64 and <= 4096 (a CTA handles two rows; a thread holds up
// to 4 x 8 features of its row, as in output_norm.cu).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"
#include "ttt_internal.h"

namespace tb {

namespace {
constexpr int kR = 4;  // rounds of 16 chunks x 64 features: E <= 4096

__device__ __forceinline__ void unpack8f(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8f(const float* o) {
  return make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
}
__device__ __forceinline__ void load8(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// sum over the 4 warps that share one row (tk = which of the CTA's two rows)
__device__ __forceinline__ float row_sum(float v, float (*red)[4], int tk, int w4, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  if (lane == 0) red[tk][w4] = v;
  __syncthreads();
  const float r = (red[tk][0] + red[tk][1]) + (red[tk][2] + red[tk][3]);
  __syncthreads();
  return r;
}
}  // namespace

__global__ void __launch_bounds__(256)
ln_affine_fwd_kernel(const uint4* __restrict__ X, const float* __restrict__ A, const float* __restrict__ C,
                     uint4* __restrict__ out, int L, int E, int text_len, float eps) {
  __shared__ float red[2][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tk = warp >> 2, w4 = warp & 3;
  const int b = blockIdx.y, C64 = E / 64, E8 = E / 8;
  const int l = min(blockIdx.x * 2 + tk, L - 1);
  const bool store = blockIdx.x * 2 + tk < L;
  const int cl = lane >> 3, piece = lane & 7;
  const uint4* xr = X + ((size_t)b * L + l) * E8;
  float x[kR][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kR; ++i) {
    const int c = 16 * i + 4 * w4 + cl;
#pragma unroll
    for (int e = 0; e < 8; ++e) x[i][e] = 0.f;
    if (c < C64) {
      unpack8f(xr[c * 8 + piece], x[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += x[i][e];
    }
  }
  const float mean = row_sum(sum, red, tk, w4, lane) / (float)E;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kR; ++i)
    if (16 * i + 4 * w4 + cl < C64)
#pragma unroll
      for (int e = 0; e < 8; ++e) { x[i][e] -= mean; sq = fmaf(x[i][e], x[i][e], sq); }
  const float rstd = rsqrtf(row_sum(sq, red, tk, w4, lane) / (float)E + eps);
  if (!store) return;
  const size_t seg = ((size_t)b * 2 + (l < text_len ? 0 : 1)) * E;
#pragma unroll
  for (int i = 0; i < kR; ++i) {
    const int c = 16 * i + 4 * w4 + cl;
    if (c < C64) {
      float a[8], cc[8], o[8];
      load8(A + seg + c * 64 + 8 * piece, a);
      load8(C + seg + c * 64 + 8 * piece, cc);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf(x[i][e] * rstd, a[e], cc[e]);
      out[((size_t)b * L + l) * E8 + c * 8 + piece] = pack8f(o);
    }
  }
}

// One CTA walks `rows_per_cta` consecutive rows of ONE (batch, segment) -- rows [r0, r1) are clipped to the segment -- two at
// a time, keeping d A / d C partials in registers.
__global__ void __launch_bounds__(256)
ln_affine_bwd_kernel(const uint4* __restrict__ X, const float* __restrict__ A, const uint4* __restrict__ gout,
                     uint4* __restrict__ gX, float* __restrict__ dA, float* __restrict__ dC, int L, int E, int text_len,
                     float eps, int rows_per_cta, int text_ctas) {
  __shared__ float red[2][4];
  __shared__ float acc[2][4096];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tk = warp >> 2, w4 = warp & 3;
  const int b = blockIdx.y, C64 = E / 64, E8 = E / 8;
  const int cl = lane >> 3, piece = lane & 7;
  const bool text = (int)blockIdx.x < text_ctas;
  const int seg_lo = text ? 0 : text_len, seg_hi = text ? text_len : L;
  const int r0 = seg_lo + (text ? blockIdx.x : blockIdx.x - text_ctas) * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, seg_hi);
  const size_t seg = ((size_t)b * 2 + (text ? 0 : 1)) * E;
  float a[kR][8], da[kR][8], dc[kR][8];
#pragma unroll
  for (int i = 0; i < kR; ++i) {
    const int c = 16 * i + 4 * w4 + cl;
#pragma unroll
    for (int e = 0; e < 8; ++e) { a[i][e] = 0.f; da[i][e] = 0.f; dc[i][e] = 0.f; }
    if (c < C64) load8(A + seg + c * 64 + 8 * piece, a[i]);
  }
  for (int i = threadIdx.x; i < 2 * 4096; i += 256) (&acc[0][0])[i] = 0.f;
  for (int r = r0; r < r1; r += 2) {
    const bool valid = r + tk < r1;
    const int l = min(r + tk, r1 - 1);
    const uint4* xr = X + ((size_t)b * L + l) * E8;
    const uint4* gr = gout + ((size_t)b * L + l) * E8;
    float x[kR][8], g[kR][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      const int c = 16 * i + 4 * w4 + cl;
#pragma unroll
      for (int e = 0; e < 8; ++e) { x[i][e] = 0.f; g[i][e] = 0.f; }
      if (c < C64) {
        unpack8f(xr[c * 8 + piece], x[i]);
        unpack8f(gr[c * 8 + piece], g[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += x[i][e];
      }
    }
    const float mean = row_sum(sum, red, tk, w4, lane) / (float)E;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kR; ++i)
      if (16 * i + 4 * w4 + cl < C64)
#pragma unroll
        for (int e = 0; e < 8; ++e) { x[i][e] -= mean; sq = fmaf(x[i][e], x[i][e], sq); }
    const float rstd = rsqrtf(row_sum(sq, red, tk, w4, lane) / (float)E + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kR; ++i)
      if (16 * i + 4 * w4 + cl < C64)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          x[i][e] *= rstd;  // x_hat
          if (valid) { da[i][e] = fmaf(g[i][e], x[i][e], da[i][e]); dc[i][e] += g[i][e]; }
          g[i][e] *= a[i][e];  // d x_hat
          s1 += g[i][e];
          s2 = fmaf(g[i][e], x[i][e], s2);
        }
    const float m1 = row_sum(s1, red, tk, w4, lane) / (float)E;
    const float m2 = row_sum(s2, red, tk, w4, lane) / (float)E;
    if (valid) {
#pragma unroll
      for (int i = 0; i < kR; ++i) {
        const int c = 16 * i + 4 * w4 + cl;
        if (c < C64) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = rstd * (g[i][e] - m1 - x[i][e] * m2);
          gX[((size_t)b * L + l) * E8 + c * 8 + piece] = pack8f(o);
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kR; ++i) {
    const int c = 16 * i + 4 * w4 + cl;
    if (c < C64)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&acc[0][c * 64 + 8 * piece + e], da[i][e]);
        atomicAdd(&acc[1][c * 64 + 8 * piece + e], dc[i][e]);
      }
  }
  __syncthreads();
  if (r0 < r1)
    for (int f = threadIdx.x; f < E; f += 256) {
      atomicAdd(dA + seg + f, acc[0][f]);
      atomicAdd(dC + seg + f, acc[1][f]);
    }
}

// out = x + G[b, seg] * y ; one thread = one 16-byte chunk, grid-stride
__global__ void __launch_bounds__(256)
gate_add_fwd_kernel(const uint4* __restrict__ X, const uint4* __restrict__ Y, const float* __restrict__ G,
                    uint4* __restrict__ out, int B, int L, int E8, int text_len) {
  const size_t total = (size_t)B * L * E8;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int e8 = (int)(idx % E8);
    const size_t row = idx / E8;
    const int l = (int)(row % L);
    const size_t b = row / L;
    float x[8], y[8], g[8];
    unpack8f(X[idx], x);
    unpack8f(Y[idx], y);
    load8(G + ((b * 2 + (l < text_len ? 0 : 1)) * (size_t)E8 + e8) * 8, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(g[i], y[i], x[i]);
    out[idx] = pack8f(x);
  }
}

// d y = G * g ; d G[b,seg] += sum_rows g * y   (d x = g is the caller's tensor itself).  grid (E8 / 128, row blocks, B)
__global__ void __launch_bounds__(128)
gate_add_bwd_kernel(const uint4* __restrict__ gout, const uint4* __restrict__ Y, const float* __restrict__ G,
                    uint4* __restrict__ dY, float* __restrict__ dG, int L, int E8, int text_len, int rows_per_block) {
  const int e8 = blockIdx.x * blockDim.x + threadIdx.x;
  if (e8 >= E8) return;
  const size_t b = blockIdx.z;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, L);
  float gt[8], gv[8], at[8], av[8];
  load8(G + ((b * 2 + 0) * (size_t)E8 + e8) * 8, gt);
  load8(G + ((b * 2 + 1) * (size_t)E8 + e8) * 8, gv);
#pragma unroll
  for (int i = 0; i < 8; ++i) { at[i] = 0.f; av[i] = 0.f; }
  for (int l = r0; l < r1; ++l) {
    const size_t idx = (b * L + l) * E8 + e8;
    float g[8], y[8], d[8];
    unpack8f(gout[idx], g);
    unpack8f(Y[idx], y);
    const bool is_text = l < text_len;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      d[i] = (is_text ? gt[i] : gv[i]) * g[i];
      if (is_text) at[i] = fmaf(g[i], y[i], at[i]);
      else         av[i] = fmaf(g[i], y[i], av[i]);
    }
    dY[idx] = pack8f(d);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (r0 < text_len) atomicAdd(dG + ((b * 2 + 0) * (size_t)E8 + e8) * 8 + i, at[i]);
    if (r1 > text_len) atomicAdd(dG + ((b * 2 + 1) * (size_t)E8 + e8) * 8 + i, av[i]);
  }
}

static bool adaln_args_ok(int B, int L, int E, int text_len) {
  return B > 0 && L > 0 && E > 0 && E % 64 == 0 && E <= 4096 && text_len >= 0 && text_len <= L;
}

cudaError_t launch_ln_affine(const void* x, const float* A, const float* C, void* out, int B, int L, int E, int text_len,
                             float eps, cudaStream_t stream) {
  if (!adaln_args_ok(B, L, E, text_len)) { g_where = "bad sizes (E % 64 == 0, E <= 4096)"; return cudaErrorInvalidValue; }
  g_where = "ln_affine launch";
  dim3 grid((L + 1) / 2, B);
  ln_affine_fwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), A, C, reinterpret_cast<uint4*>(out), L, E,
                                                 text_len, eps);
  return cudaGetLastError();
}

cudaError_t launch_ln_affine_backward(const void* x, const float* A, const void* gout, void* gx, float* dA, float* dC, int B,
                                      int L, int E, int text_len, float eps, cudaStream_t stream) {
  if (!adaln_args_ok(B, L, E, text_len)) { g_where = "bad sizes (E % 64 == 0, E <= 4096)"; return cudaErrorInvalidValue; }
  TB_TRY(cudaMemsetAsync(dA, 0, (size_t)B * 2 * E * sizeof(float), stream), "memset dA");
  TB_TRY(cudaMemsetAsync(dC, 0, (size_t)B * 2 * E * sizeof(float), stream), "memset dC");
  const int rpc = 32;
  const int text_ctas = (text_len + rpc - 1) / rpc, vid_ctas = (L - text_len + rpc - 1) / rpc;
  g_where = "ln_affine backward launch";
  dim3 grid(text_ctas + vid_ctas, B);
  ln_affine_bwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), A, reinterpret_cast<const uint4*>(gout),
                                                 reinterpret_cast<uint4*>(gx), dA, dC, L, E, text_len, eps, rpc, text_ctas);
  return cudaGetLastError();
}

cudaError_t launch_gate_add(const void* x, const void* y, const float* G, void* out, int B, int L, int E, int text_len,
                            cudaStream_t stream) {
  if (!adaln_args_ok(B, L, E, text_len)) { g_where = "bad sizes"; return cudaErrorInvalidValue; }
  g_where = "gate_add launch";
  const size_t total = (size_t)B * L * (E / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gate_add_fwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(y), G,
                                                            reinterpret_cast<uint4*>(out), B, L, E / 8, text_len);
  return cudaGetLastError();
}

cudaError_t launch_gate_add_backward(const void* gout, const void* y, const float* G, void* dy, float* dG, int B, int L, int E,
                                     int text_len, cudaStream_t stream) {
  if (!adaln_args_ok(B, L, E, text_len)) { g_where = "bad sizes"; return cudaErrorInvalidValue; }
  TB_TRY(cudaMemsetAsync(dG, 0, (size_t)B * 2 * E * sizeof(float), stream), "memset dG");
  g_where = "gate_add backward launch";
  const int E8 = E / 8, bx = (E8 + 127) / 128;
  int by = (148 * 8) / (bx * B);
  if (by < 1) by = 1;
  if (by > L) by = L;
  const int rpb = (L + by - 1) / by;
  dim3 grid(bx, (L + rpb - 1) / rpb, B);
  gate_add_bwd_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const uint4*>(gout), reinterpret_cast<const uint4*>(y), G,
                                                reinterpret_cast<uint4*>(dy), dG, L, E8, text_len, rpb);
  return cudaGetLastError();
}

}  // namespace tb
