// Device helpers shared by the backward kernels (ttt_mlp_bwd.cu = sequential K-side kernel, ttt_mlp_bwd_q.cu = parallel
// Q-side kernel): GELU derivatives, SW128 row <-> register helpers, warp column sums, single-thread MMA issue helpers.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "ptx.cuh"

namespace tb {
namespace bwd {

__device__ __forceinline__ void gelu3(float z, float& g0, float& g1, float& g2) {
  // gelu, gelu' (ops/utils.py:51-54) and gelu'' (ttt_backward/matching.py:47-55)
  const float c0 = 0.79788456f, c1 = 0.79788456f * 0.044715f;
  const float z2 = z * z;
  const float a = fmaf(3.0f * c1, z2, c0);
  const float t = tanh_fast(z * fmaf(c1, z2, c0));
  const float s = fmaf(-t, t, 1.0f);
  const float hz = 0.5f * z;
  g0 = fmaf(hz, t, hz);
  g1 = fmaf(hz * s, a, fmaf(0.5f, t, 0.5f));
  g2 = s * (fmaf(2.0f, a, -c0) - z * t * a * a);
}
// gelu3 on a packed pair: the same operations in the same order, two lanes per instruction (tanh stays scalar: MUFU)
__device__ __forceinline__ void gelu3x2(f32x2 z, f32x2& g0, f32x2& g1, f32x2& g2) {
  const float c0 = 0.79788456f, c1 = 0.79788456f * 0.044715f;
  const f32x2 C0 = pk2(c0), C1 = pk2(c1), C3 = pk2(3.0f * c1), HALF = pk2(0.5f), ONE = pk2(1.0f), NEG1 = pk2(-1.0f);
  const f32x2 z2 = mul2(z, z);
  const f32x2 a = fma2(C3, z2, C0);
  const f32x2 u = mul2(z, fma2(C1, z2, C0));
  float u0, u1;
  up2(u, u0, u1);
  const f32x2 t = pk2(tanh_fast(u0), tanh_fast(u1));
  const f32x2 s = fma2(mul2(t, NEG1), t, ONE);
  const f32x2 hz = mul2(HALF, z);
  g0 = fma2(hz, t, hz);
  g1 = fma2(mul2(hz, s), a, fma2(HALF, t, HALF));
  g2 = mul2(s, sub2(fma2(pk2(2.0f), a, pk2(-c0)), mul2(mul2(mul2(z, t), a), a)));
}
__device__ __forceinline__ float gelu1(float z, float& g1) {
  const float c0 = 0.79788456f, c1 = 0.79788456f * 0.044715f;
  const float z2 = z * z;
  const float t = tanh_fast(z * fmaf(c1, z2, c0));
  const float hz = 0.5f * z;
  g1 = fmaf(hz * fmaf(-t, t, 1.0f), fmaf(3.0f * c1, z2, c0), fmaf(0.5f, t, 0.5f));
  return fmaf(hz, t, hz);
}

// 32 fp32 -> bf16 -> 4 consecutive chunks of one SW128 row
__device__ __forceinline__ void st_row32(uint32_t tile, int row, int chunk0, const float* v) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    st_shared_v4(tile + sw128_off(row, chunk0 + c), pack_bf16(v[8 * c], v[8 * c + 1]), pack_bf16(v[8 * c + 2], v[8 * c + 3]),
                 pack_bf16(v[8 * c + 4], v[8 * c + 5]), pack_bf16(v[8 * c + 6], v[8 * c + 7]));
}
// read one 64-wide bf16 row of a SW128 tile into fp32
__device__ __forceinline__ void ld_row64(uint32_t tile, int row, float* v) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t a, b, cc, d;
    ld_shared_v4(tile + sw128_off(row, c), a, b, cc, d);
    v[8 * c + 0] = bf16_lo(a); v[8 * c + 1] = bf16_hi(a); v[8 * c + 2] = bf16_lo(b); v[8 * c + 3] = bf16_hi(b);
    v[8 * c + 4] = bf16_lo(cc); v[8 * c + 5] = bf16_hi(cc); v[8 * c + 6] = bf16_lo(d); v[8 * c + 7] = bf16_hi(d);
  }
}
// column sums over the 32 lanes of a warp of a per-lane vector v[N] (N = 32 or 64) by recursive halving.
// On return lane l holds in v[0] the sum of element l (N=32) or in v[0], v[1] the sums of elements l and l+32 (N=64).
template <int N>
__device__ __forceinline__ void warp_colsum(float* v, int lane) {
  if (N == 64) {  // first fold 64 -> 32 pairs kept as (v[q], v[q+32]) handled by two independent 32-wide reductions
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      const bool up = (lane & m) != 0;
#pragma unroll
      for (int q = 0; q < m; ++q) {
        float a0 = v[q], b0 = v[q + m], a1 = v[32 + q], b1 = v[32 + q + m];
        float s0 = up ? a0 : b0, k0 = up ? b0 : a0, s1 = up ? a1 : b1, k1 = up ? b1 : a1;
        v[q] = k0 + __shfl_xor_sync(0xffffffffu, s0, m);
        v[32 + q] = k1 + __shfl_xor_sync(0xffffffffu, s1, m);
      }
    }
    v[1] = v[32];
  } else {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      const bool up = (lane & m) != 0;
#pragma unroll
      for (int q = 0; q < m; ++q) {
        float a0 = v[q], b0 = v[q + m];
        float s0 = up ? a0 : b0, k0 = up ? b0 : a0;
        v[q] = k0 + __shfl_xor_sync(0xffffffffu, s0, m);
      }
    }
  }
}

// column sums of a per-lane vector v[16] over the 32 lanes: afterwards v[0] = sum over lanes of element ((lane >> 1) & 15)
__device__ __forceinline__ void warp_colsum16(float* v, int lane) {
#pragma unroll
  for (int m = 8; m >= 1; m >>= 1) {
    const bool up = (lane & (2 * m)) != 0;  // lane bits 4..1 select the element, bit 0 is folded last
#pragma unroll
    for (int q = 0; q < m; ++q) {
      const float a0 = v[q], b0 = v[q + m];
      const float snd = up ? a0 : b0, kp = up ? b0 : a0;
      v[q] = kp + __shfl_xor_sync(0xffffffffu, snd, 2 * m);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}
// load 16 consecutive bf16 (two 16-B chunks c0, c0+1 of row r) from a SW128 tile
__device__ __forceinline__ void ld_row16(uint32_t tile, int row, int chunk0, float* v) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t a, b, cc, d;
    ld_shared_v4(tile + sw128_off(row, chunk0 + c), a, b, cc, d);
    v[8 * c + 0] = bf16_lo(a); v[8 * c + 1] = bf16_hi(a); v[8 * c + 2] = bf16_lo(b); v[8 * c + 3] = bf16_hi(b);
    v[8 * c + 4] = bf16_lo(cc); v[8 * c + 5] = bf16_hi(cc); v[8 * c + 6] = bf16_lo(d); v[8 * c + 7] = bf16_hi(d);
  }
}
__device__ __forceinline__ void st_row16(uint32_t tile, int row, int chunk0, const float* v) {
#pragma unroll
  for (int c = 0; c < 2; ++c)
    st_shared_v4(tile + sw128_off(row, chunk0 + c), pack_bf16(v[8 * c], v[8 * c + 1]), pack_bf16(v[8 * c + 2], v[8 * c + 3]),
                 pack_bf16(v[8 * c + 4], v[8 * c + 5]), pack_bf16(v[8 * c + 6], v[8 * c + 7]));
}
__device__ __forceinline__ void st_global16(__nv_bfloat16* g, const float* v) {
#pragma unroll
  for (int c = 0; c < 2; ++c)
    *reinterpret_cast<uint4*>(g + 8 * c) = make_uint4(pack_bf16(v[8 * c], v[8 * c + 1]), pack_bf16(v[8 * c + 2], v[8 * c + 3]),
                                                      pack_bf16(v[8 * c + 4], v[8 * c + 5]), pack_bf16(v[8 * c + 6], v[8 * c + 7]));
}

// ---- MMA issue helpers (single thread) -----------------------------------------------------------------------------
// hidden-lane output: D[h] (128 lanes x N cols) = A_tile[h] (K-major, [256][64]) . B  ; 4 k-steps
__device__ __forceinline__ void mma_hid(uint32_t d0, uint32_t d1, uint32_t a_tile, uint32_t b_tile, bool b_mn, int n,
                                        bool acc) {
  const uint32_t idesc = make_idesc_bf16(128, n, false, b_mn);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint64_t da = make_desc_sw128(a_tile + h * 16384, 16, 1024);
    const uint64_t db = b_mn ? make_desc_sw128(b_tile, 1024, 1024) : make_desc_sw128(b_tile, 16, 1024);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_ss(h ? d1 : d0, desc_advance(da, 32 * k), desc_advance(db, b_mn ? 2048 * k : 32 * k), idesc, acc || k > 0);
  }
}
// token-lane output: D = A_tile (hidden-lane tile viewed MN-major, [256 j][64 tok]) . B_tile ([256 j][64 f], MN-major);
// 16 k-steps over the hidden dim.  LBO = 0 makes the second 64-row block of A alias the first, so rows 64-127 of D are a
// copy of rows 0-63: every token row is then readable from two TMEM lane halves and all 8 warps share the token phases
// (thread <-> (row, 16-column quarter)); pinned by umma self-test mode 6.
__device__ __forceinline__ void mma_tok(uint32_t d, uint32_t a_tile, uint32_t b_tile, bool acc) {
  const uint32_t idesc = make_idesc_bf16(128, 64, true, true);
  const uint64_t da = make_desc_sw128(a_tile, 0, 1024);
  const uint64_t db = make_desc_sw128(b_tile, 1024, 1024);
#pragma unroll
  for (int k = 0; k < 16; ++k) umma_ss(d, desc_advance(da, 2048 * k), desc_advance(db, 2048 * k), idesc, acc || k > 0);
}

#define MMA_WAIT()                 \
  do {                             \
    mbar_wait(mma_bar, mma_phase); \
    mma_phase ^= 1;                \
    tc_fence_after();              \
  } while (0)
#define PHASE_SYNC()     \
  do {                   \
    fence_proxy_async(); \
    tc_fence_before();   \
    __syncthreads();     \
  } while (0)

}  // namespace bwd
}  // namespace tb
