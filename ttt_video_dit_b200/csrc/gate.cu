// Learned residual gate + sequence reversal around the bidirectional TTT pass (HBM-bound elementwise kernels).
// Reference: ttt/models/cogvideo/dit.py:90-103 (SSMGating: tanh(alpha) * x), :213-217 (_reverse_text_chunks),
// :219-222 (_gate), :224-266 (_ssm_forward).  The reference makes 2 clones + 4 flips + 3 cats per layer; here each
// direction is one pass:   out[l] = res[l] + tanh(alpha(l)) * s[src(l)]   (+ optional second store rev[perm(l)] = out[l])
// with perm = the involution "text chunks in reverse order, video tokens flipped" (perm(perm(l)) = l).
// Layout: [B, L, E] bf16, text tokens first (L_text = seq_text_length), E % 8 == 0; alpha fp32 [E].
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ttt_internal.h"

namespace tb {

__device__ __forceinline__ int perm_index(int l, int L, int text_len, int num_chunks) {
  if (l < text_len) {
    const int cl = text_len / num_chunks;
    const int c = l / cl;
    return (num_chunks - 1 - c) * cl + (l - c * cl);
  }
  return text_len + (L - 1 - l);  // video part: flip
}

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// one thread = one 16-byte chunk (8 bf16) of one token row; grid-stride over B*L*(E/8) chunks
template <bool kPermS, bool kWriteRev>
__global__ void __launch_bounds__(256)
gate_fwd_kernel(const uint4* __restrict__ res, const uint4* __restrict__ s, const float* __restrict__ a_text,
                const float* __restrict__ a_video, uint4* __restrict__ out, uint4* __restrict__ rev, int B, int L, int E8,
                int text_len, int num_chunks) {
  const size_t total = (size_t)B * L * E8;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int e8 = (int)(idx % E8);
    const size_t row = idx / E8;
    const int l = (int)(row % L);
    const size_t b = row / L;
    const int pl = perm_index(l, L, text_len, num_chunks);
    const float* al = (l < text_len ? a_text : a_video) + 8 * e8;
    const uint4 rv = res[idx];
    const uint4 sv = s[kPermS ? ((b * L + pl) * E8 + e8) : idx];
    float r[8], x[8];
    unpack8(rv, r);
    unpack8(sv, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = fmaf(tanhf(al[i]), x[i], r[i]);
    const uint4 o = pack8(r);
    out[idx] = o;
    if (kWriteRev) rev[(b * L + pl) * E8 + e8] = o;
  }
}

// backward: g = dout[l] (+ drev[perm(l)]) ; dres[l] = g ; ds[src(l)] = tanh(alpha) g ; dalpha += (1-tanh^2) sum g*s[src(l)]
// grid: (E8 chunks / 32?, row blocks); each thread owns one e8 chunk and loops over a slice of rows -> register partials
template <bool kPermS, bool kHasRev>
__global__ void __launch_bounds__(256)
gate_bwd_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ drev, const uint4* __restrict__ s,
                const float* __restrict__ a_text, const float* __restrict__ a_video, uint4* __restrict__ dres,
                uint4* __restrict__ ds, float* __restrict__ da_text, float* __restrict__ da_video, int B, int L, int E8,
                int text_len, int num_chunks, int rows_per_block) {
  const int e8 = blockIdx.x * blockDim.x + threadIdx.x;
  if (e8 >= E8) return;
  const size_t rows = (size_t)B * L;
  const size_t r0 = (size_t)blockIdx.y * rows_per_block;
  const size_t r1 = (r0 + rows_per_block < rows) ? r0 + rows_per_block : rows;
  float tt[8], tv[8], acc_t[8], acc_v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    tt[i] = tanhf(a_text[8 * e8 + i]);
    tv[i] = tanhf(a_video[8 * e8 + i]);
    acc_t[i] = 0.f;
    acc_v[i] = 0.f;
  }
  for (size_t row = r0; row < r1; ++row) {
    const int l = (int)(row % L);
    const size_t b = row / L;
    const int pl = perm_index(l, L, text_len, num_chunks);
    const size_t idx = row * E8 + e8, pidx = (b * L + pl) * E8 + e8;
    float g[8], x[8];
    unpack8(dout[idx], g);
    if (kHasRev) {
      float g2[8];
      unpack8(drev[pidx], g2);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += g2[i];
    }
    unpack8(s[kPermS ? pidx : idx], x);
    dres[idx] = pack8(g);
    const bool is_text = l < text_len;
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      d[i] = (is_text ? tt[i] : tv[i]) * g[i];
      if (is_text) acc_t[i] = fmaf(g[i], x[i], acc_t[i]);
      else         acc_v[i] = fmaf(g[i], x[i], acc_v[i]);
    }
    ds[kPermS ? pidx : idx] = pack8(d);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    atomicAdd(&da_text[8 * e8 + i], acc_t[i] * (1.f - tt[i] * tt[i]));
    atomicAdd(&da_video[8 * e8 + i], acc_v[i] * (1.f - tv[i] * tv[i]));
  }
}

static int gate_grid(size_t total) {
  size_t blocks = (total + 255) / 256;
  const size_t cap = 148 * 16;  // a multiple of the SM count; grid-stride covers the rest
  return (int)(blocks < cap ? blocks : cap);
}

cudaError_t launch_gate_forward(const void* res, const void* s, const float* a_text, const float* a_video, void* out,
                                void* rev, int B, int L, int E, int text_len, int num_chunks, int perm_s,
                                cudaStream_t stream) {
  if (B <= 0 || L <= 0 || E <= 0 || E % 8 || text_len < 0 || text_len > L || num_chunks <= 0 || text_len % num_chunks)
    return cudaErrorInvalidValue;
  const int E8 = E / 8;
  const size_t total = (size_t)B * L * E8;
  const int grid = gate_grid(total);
  const uint4 *r4 = (const uint4*)res, *s4 = (const uint4*)s;
  uint4 *o4 = (uint4*)out, *v4 = (uint4*)rev;
  if (perm_s) {
    if (rev) gate_fwd_kernel<true, true><<<grid, 256, 0, stream>>>(r4, s4, a_text, a_video, o4, v4, B, L, E8, text_len, num_chunks);
    else     gate_fwd_kernel<true, false><<<grid, 256, 0, stream>>>(r4, s4, a_text, a_video, o4, v4, B, L, E8, text_len, num_chunks);
  } else {
    if (rev) gate_fwd_kernel<false, true><<<grid, 256, 0, stream>>>(r4, s4, a_text, a_video, o4, v4, B, L, E8, text_len, num_chunks);
    else     gate_fwd_kernel<false, false><<<grid, 256, 0, stream>>>(r4, s4, a_text, a_video, o4, v4, B, L, E8, text_len, num_chunks);
  }
  return cudaGetLastError();
}

cudaError_t launch_gate_backward(const void* dout, const void* drev, const void* s, const float* a_text,
                                 const float* a_video, void* dres, void* ds, float* da_text, float* da_video, int B,
                                 int L, int E, int text_len, int num_chunks, int perm_s, cudaStream_t stream) {
  if (B <= 0 || L <= 0 || E <= 0 || E % 8 || text_len < 0 || text_len > L || num_chunks <= 0 || text_len % num_chunks)
    return cudaErrorInvalidValue;
  const int E8 = E / 8;
  cudaError_t e = cudaMemsetAsync(da_text, 0, E * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(da_video, 0, E * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  const size_t rows = (size_t)B * L;
  const int bx = (E8 + 127) / 128;
  int by = (148 * 8) / bx;
  if ((size_t)by > rows) by = (int)rows;
  const int rpb = (int)((rows + by - 1) / by);
  dim3 grid(bx, (unsigned)((rows + rpb - 1) / rpb));
  const uint4 *g4 = (const uint4*)dout, *r4 = (const uint4*)drev, *s4 = (const uint4*)s;
  uint4 *dr = (uint4*)dres, *dsv = (uint4*)ds;
  if (perm_s) {
    if (drev) gate_bwd_kernel<true, true><<<grid, 128, 0, stream>>>(g4, r4, s4, a_text, a_video, dr, dsv, da_text, da_video, B, L, E8, text_len, num_chunks, rpb);
    else      gate_bwd_kernel<true, false><<<grid, 128, 0, stream>>>(g4, r4, s4, a_text, a_video, dr, dsv, da_text, da_video, B, L, E8, text_len, num_chunks, rpb);
  } else {
    if (drev) gate_bwd_kernel<false, true><<<grid, 128, 0, stream>>>(g4, r4, s4, a_text, a_video, dr, dsv, da_text, da_video, B, L, E8, text_len, num_chunks, rpb);
    else      gate_bwd_kernel<false, false><<<grid, 128, 0, stream>>>(g4, r4, s4, a_text, a_video, dr, dsv, da_text, da_video, B, L, E8, text_len, num_chunks, rpb);
  }
  return cudaGetLastError();
}

}  // namespace tb
