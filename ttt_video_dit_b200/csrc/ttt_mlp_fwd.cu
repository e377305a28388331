// TTT-MLP forward scan for sm_100a (tcgen05 + TMEM + TMA).  One CTA per (batch, head) sequence.
//
// Replaces ttt-tk/kernels/ttt/ttt.cu:85-721 (fwd_ttt_mlp_ker / ttt_forward, Hopper wgmma) with a Blackwell design;
// the arithmetic it must reproduce is the reference's eager step ttt/models/ssm/ops/ttt_mlp.py:9-67 in its primal
// form, LayerNorm eps 1e-8 (ttt/models/ssm/ops/utils.py:4,21).
//
// Layout idea ("hidden units on TMEM lanes"): every 256-wide tensor is kept transposed so that the MMA M dimension is
// the 4F=256 hidden dimension (2 x M=128, full-rate tcgen05) and each thread owns one hidden unit:
//   W1^T [256 x 64] fp32 and W2 [256 x 64] fp32 are *persistent TMEM accumulators*; the TTT update
//   W -= (eta*X)^T grad is an accumulate-MMA, the master copy never leaves TMEM.
//   D1   = W1b^T . [K_t | Q_{t-1}]^T   -> [256 hidden x 128 tokens]   (Z1_t^T and Zbar1_{t-1}^T in ONE MMA batch)
//   D2   = [X2_t ; X2bar_{t-1}] . W2b  -> [128 tokens x 64]           (Z2_t and Zbar2_{t-1}, M=128 full rate)
//   D3   = W2b . G2^T                  -> [256 x 64 tokens]           (-eta * gradZ2 W2^T)
//   W2  += X2^T . G2 ;  W1^T += G1^T . K                               (G = -eta*grad, bf16)
// The Q side of mini-batch t-1 rides along with the K side of mini-batch t (software pipelining by one step), which
// makes both "half-size" GEMMs of the recurrence M=128.  bf16 operand copies of W (W1b^T, W2b) are re-materialised
// from TMEM once per step.  All smem operand tiles use the SW128 row-tile convention of ptx.cuh.
//
// Operand format: every MMA operand tile is fp16 by default (template kF16): the state images, X2 / G tiles are produced
// here and simply packed as f16; the K and Q token tiles arrive as bf16 by TMA and are converted in place (exact) right
// after landing.  bf16 -> f16 operands cut the operand rounding 8x, which matters because the first mini-batch of a
// sequence (b2 = 0: LayerNorm std ~3e-3) amplifies it by 1/std into the whole state trajectory, and the gradient of some
// heads is ill-conditioned w.r.t. that trajectory (measured: 8 % gradient deviation on the worst of 48 heads with bf16
// operands, profiles/r02_diag_bwd*.log).  Mixed f16 / bf16 operands in one MMA trap, so V (never an MMA operand) stays
// bf16.  TTT_B200_OPERANDS=bf16 selects the bf16 instantiation (A/B runs).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "ptx.cuh"
#include "ttt_internal.h"

namespace tb {

constexpr int CS = 64, F = 64, HID = 256;
constexpr int NT = 256;

// shared memory map (bytes)
constexpr uint32_t SM_W1B = 0;                    // W1^T bf16 [256][64]   (A, K-major, K = F)
constexpr uint32_t SM_W2B = 32768;                // W2   bf16 [256][64]   (A K-major for D3; B MN-major for D2)
constexpr uint32_t SM_X2 = 65536;                 // block0: X2^T [256][64 tok] ; block1: X2bar^T (later reused for G1^T)
constexpr uint32_t SM_KQ = 131072;                // 2 slots x { K_t [64][64] ; Q_{t-1} [64][64] }
constexpr uint32_t SM_V = SM_KQ + 2 * 16384;      // 2 slots x V_t [64][64]
constexpr uint32_t SM_G2 = SM_V + 2 * 8192;       // G2 = -eta*gradZ2 bf16 [64 tok][64]
constexpr uint32_t SM_MISC = SM_G2 + 8192;        // b2[64] f32, ln_w[64], ln_b[64], barriers, tmem ptr
constexpr uint32_t SM_XB = SM_MISC + 1024;        // LN exchange: float2[2][128] + float2[2][64] + db2 accumulator[64]
constexpr uint32_t SM_TOTAL = SM_XB + 4096;

// TMEM column map (512 columns allocated)
constexpr uint32_t TM_W1 = 0;    // + 64*h
constexpr uint32_t TM_W2 = 128;  // + 64*h
constexpr uint32_t TM_D1 = 256;  // + 128*h ; cols [0,64) K side, [64,128) Q side
constexpr uint32_t TM_D3 = 256;  // + 128*h ; aliases the K side of D1
constexpr uint32_t TM_D2 = 320;  // aliases the Q side of D1 half 0

// Scalar definitions of the two activations (tanh-GELU, ops/utils.py:45-54).  The kernel runs the packed-pair forms below
// (same operations in the same order on two tokens per instruction); these stay as the readable statement of the math.
__device__ __forceinline__ float gelu_and_grad(float z, float& grad) {
  const float c0 = 0.79788456f, c1 = 0.79788456f * 0.044715f;
  float z2 = z * z;
  float u = z * fmaf(c1, z2, c0);
  float t = tanh_fast(u);
  float hz = 0.5f * z;
  // ops/utils.py:51-54: 0.5*x*((1-t^2)*(0.79788456+0.1070322243*x^2)) + 0.5*(1+t)
  grad = fmaf(hz * fmaf(-t, t, 1.0f), fmaf(3.0f * c1, z2, c0), fmaf(0.5f, t, 0.5f));
  return fmaf(hz, t, hz);
}
__device__ __forceinline__ float gelu_only(float z) {
  const float c0 = 0.79788456f, c1 = 0.79788456f * 0.044715f;
  float u = z * fmaf(c1, z * z, c0);
  float t = tanh_fast(u);
  float hz = 0.5f * z;
  return fmaf(hz, t, hz);
}

// the same two functions on packed pairs (same operations, same order, two tokens per instruction; tanh stays scalar)
__device__ __forceinline__ f32x2 tanh2(f32x2 u) {
  float u0, u1;
  up2(u, u0, u1);
  return pk2(tanh_fast(u0), tanh_fast(u1));
}
__device__ __forceinline__ f32x2 gelu_and_grad2(f32x2 z, f32x2& grad) {
  const float c0 = 0.79788456f, c1 = 0.79788456f * 0.044715f;
  const f32x2 C0 = pk2(c0), HALF = pk2(0.5f);
  const f32x2 z2 = mul2(z, z);
  const f32x2 t = tanh2(mul2(z, fma2(pk2(c1), z2, C0)));
  const f32x2 hz = mul2(HALF, z);
  grad = fma2(mul2(hz, fma2(mul2(t, pk2(-1.0f)), t, pk2(1.0f))), fma2(pk2(3.0f * c1), z2, C0), fma2(HALF, t, HALF));
  return fma2(hz, t, hz);
}
__device__ __forceinline__ f32x2 gelu_only2(f32x2 z) {
  const float c0 = 0.79788456f, c1 = 0.79788456f * 0.044715f;
  const f32x2 t = tanh2(mul2(z, fma2(pk2(c1), mul2(z, z), pk2(c0))));
  const f32x2 hz = mul2(pk2(0.5f), z);
  return fma2(hz, t, hz);
}

struct FwdParams {
  const __nv_bfloat16* last_eta;  // [B,H,NC,64]
  const float* ln_w;              // [H,64]
  const float* ln_b;
  const float *W1, *b1, *W2, *b2;      // initial state [B,H,64,256],[B,H,256],[B,H,256,64],[B,H,64]
  float *W1c, *b1c, *W2c, *b2c;        // checkpoints [B,H,K,...] (may be null)
  float *W1o, *b1o, *W2o, *b2o;        // final state (may be null)
  __nv_bfloat16* Out;                  // [B,H,NC,64,64]
  int B, H, NC, ckpt_group, K;
  unsigned* dbg;  // phase-timing buffer (debug builds)
};

// store one thread's state rows (fp32, still in registers) to a [64][256] W1 image and a [256][64] W2 image
__device__ __forceinline__ void store_w1_col(float* W1g, int j, const uint32_t* v, int f0) {
#pragma unroll
  for (int i = 0; i < 32; ++i) W1g[(size_t)(f0 + i) * HID + j] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void store_w2_row(float* W2g, int j, const uint32_t* v, int f0) {
  float4* dst = reinterpret_cast<float4*>(W2g + (size_t)j * F + f0);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                         __uint_as_float(v[4 * i + 3]));
}
// 32 fp32 (registers) -> operand format -> 4 chunks of a SW128 row
template <bool kF16>
__device__ __forceinline__ void store_row_op(uint32_t tile_saddr, int row, int chunk0, const uint32_t* v) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t p0 = pack_op<kF16>(__uint_as_float(v[8 * c + 0]), __uint_as_float(v[8 * c + 1]));
    uint32_t p1 = pack_op<kF16>(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3]));
    uint32_t p2 = pack_op<kF16>(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5]));
    uint32_t p3 = pack_op<kF16>(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7]));
    st_shared_v4(tile_saddr + sw128_off(row, chunk0 + c), p0, p1, p2, p3);
  }
}
// in-place bf16 -> fp16 conversion of one [64][64] token tile (8 KB = 512 chunks of 16 bytes; 2 per thread; element-wise,
// so the swizzle does not matter)
__device__ __forceinline__ void tile_bf16_to_f16(uint32_t tile_saddr, int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint32_t a = tile_saddr + (uint32_t)(tid + 256 * i) * 16u;
    uint32_t w0, w1, w2, w3;
    ld_shared_v4(a, w0, w1, w2, w3);
    st_shared_v4(a, bf16x2_to_f16x2(w0), bf16x2_to_f16x2(w1), bf16x2_to_f16x2(w2), bf16x2_to_f16x2(w3));
  }
}

template <bool kF16>
__global__ void __launch_bounds__(NT, 1)
ttt_mlp_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = uniform_warp_id();  // == warp, provably warp-uniform: single-thread issue blocks branch on it
  const int bh = blockIdx.x;
  const int head = bh % p.H;
  const int NC = p.NC;

  float* b2s = reinterpret_cast<float*>(smem + SM_MISC);
  float* lnw = b2s + 64;
  float* lnb = lnw + 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_MISC + 768);
  uint64_t* tma_bar = bars;      // [2]
  uint64_t* mma_bar = bars + 2;  // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  float2* xs1 = reinterpret_cast<float2*>(smem + SM_XB);         // [2][128]  (sum z, sum z^2) per column half
  float2* xs2 = reinterpret_cast<float2*>(smem + SM_XB + 2048);  // [2][64]   (s1, s2) of the K side
  float* db2acc = reinterpret_cast<float*>(smem + SM_XB + 3072);  // [64] column sums of G2, folded into b2 in P6

  if (warp_u == 0 && elect_one()) {
    mbar_init(&tma_bar[0], 1);
    mbar_init(&tma_bar[1], 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  if (tid < 64) {
    lnw[tid] = p.ln_w[head * 64 + tid];
    lnb[tid] = p.ln_b[head * 64 + tid];
    b2s[tid] = p.b2[(size_t)bh * 64 + tid];
    db2acc[tid] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const int half = warp >> 2;                                    // which 128-row half of the hidden dim
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;  // this warp's TMEM lane quarter
  const int j = tid;                                             // hidden unit owned by this thread (P2/P6/P8)
  const size_t row_base = (size_t)bh * p.NC * CS;                // first token row of this sequence, in [B*H*NC*CS, 64]

  // prologue TMA: K_0, V_0 into slot 0
  if (warp_u == 0 && (NC > 0) && elect_one()) {
    mbar_expect_tx(&tma_bar[0], 16384);
    tma_load_2d(smem + SM_KQ, &tmK, 0, (int)row_base, &tma_bar[0]);
    tma_load_2d(smem + SM_V, &tmV, 0, (int)row_base, &tma_bar[0]);
  }

  // ---- initial state: global fp32 -> TMEM accumulators + operand copies (+ checkpoint 0)
  float b1r = p.b1[(size_t)bh * HID + j];
  {
    const float* W1g = p.W1 + (size_t)bh * F * HID;
    const float* W2g = p.W2 + (size_t)bh * HID * F;
    uint32_t v[32];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(W1g[(size_t)(32 * c + i) * HID + j]);
      tmem_st32(tmem + lane_addr + TM_W1 + 64 * half + 32 * c, v);
      store_row_op<kF16>(sbase + SM_W1B, j, 4 * c, v);
      if (p.W1c) store_w1_col(p.W1c + ((size_t)bh * p.K) * F * HID, j, v, 32 * c);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(W2g[(size_t)j * F + 32 * c + i]);
      tmem_st32(tmem + lane_addr + TM_W2 + 64 * half + 32 * c, v);
      store_row_op<kF16>(sbase + SM_W2B, j, 4 * c, v);
      if (p.W2c) store_w2_row(p.W2c + ((size_t)bh * p.K) * HID * F, j, v, 32 * c);
    }
    if (p.b1c) p.b1c[((size_t)bh * p.K) * HID + j] = b1r;
    if (p.b2c && tid < 64) p.b2c[((size_t)bh * p.K) * F + tid] = p.b2[(size_t)bh * 64 + tid];
    tc_wait_st();
  }
  if (kF16 && NC > 0) {  // K_0 has landed long ago (the state staging above took microseconds): bf16 -> f16 in place
    mbar_wait(&tma_bar[0], 0);
    tile_bf16_to_f16(sbase + SM_KQ, tid);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();

  constexpr uint32_t IDESC_A = make_idesc_bf16(128, 128, false, false, false, kF16, kF16);  // D1: A K-major, B K-major
  constexpr uint32_t IDESC_B = make_idesc_bf16(128, 64, true, true, false, kF16, kF16);     // D2: A (X2) MN-major, B MN-major
  constexpr uint32_t IDESC_U2 = make_idesc_bf16(128, 64, false, true, false, kF16, kF16);   // W2 update: A = X2^T
  constexpr uint32_t IDESC_C = make_idesc_bf16(128, 64, false, false, false, kF16, kF16);   // D3
  constexpr uint32_t IDESC_U = make_idesc_bf16(128, 64, false, true, false, kF16, kF16);    // state updates: A K-major, B MN-major

  uint32_t mma_phase = 0;
  auto issue_p1 = [&](int it_next) {  // one thread: wait for the tiles of iteration it_next, then D1 = W1b^T . [K|Q]^T
    const int sl = it_next & 1;
    mbar_wait(&tma_bar[sl], (it_next >> 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint64_t da = make_desc_sw128(sbase + SM_W1B + h * 16384, 16, 1024);
      const uint64_t db = make_desc_sw128(sbase + SM_KQ + sl * 16384, 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ss(tmem + TM_D1 + 128 * h, desc_advance(da, 32 * k), desc_advance(db, 32 * k), IDESC_A, k > 0);
    }
    tc_commit(mma_bar);
  };
  if (warp_u == 0 && (NC > 0) && elect_one()) issue_p1(0);
  TICK_DECL(12, 224)
  uint32_t gp[32];  // gelu'(Z1) for this thread's hidden unit, 64 tokens, packed bf16x2

  for (int it = 0; it < NC + 1; ++it) {
    const int slot = it & 1;
    const bool has_k = it < NC, has_q = it > 0;
    const uint32_t kq = sbase + SM_KQ + slot * 16384;
    const uint32_t vt = sbase + SM_V + slot * 8192;

    // eta of this step for the K-side LN rows: fetched as raw bf16 bits here, converted at the point of use (P4) so the
    // global-load latency hides behind P1-P3 instead of stalling the warp at the top of the iteration
    unsigned short eta_raw = 0;
    if (has_k && (warp & 3) < 2)
      eta_raw = reinterpret_cast<const unsigned short*>(p.last_eta)[row_base + (size_t)it * CS + 32 * (warp & 3) + lane];

    TICK(0);
    mbar_wait(&tma_bar[slot], (it >> 1) & 1);
    TICK(1);
    if (warp_u == 0 && (it < NC) && elect_one()) {  // next iteration's tiles: K_{it+1}, V_{it+1} (if any) and Q_{it}
      const int ns = slot ^ 1;
      const bool nk = (it + 1) < NC;
      mbar_expect_tx(&tma_bar[ns], (nk ? 16384 : 0) + 8192);
      if (nk) {
        tma_load_2d(smem + SM_KQ + ns * 16384, &tmK, 0, (int)(row_base + (size_t)(it + 1) * CS), &tma_bar[ns]);
        tma_load_2d(smem + SM_V + ns * 8192, &tmV, 0, (int)(row_base + (size_t)(it + 1) * CS), &tma_bar[ns]);
      }
      tma_load_2d(smem + SM_KQ + ns * 16384 + 8192, &tmQ, 0, (int)(row_base + (size_t)it * CS), &tma_bar[ns]);
    }

    // ---------------- P1: D1[h] = W1b^T[h] . [K | Q]^T  (M=128, N=128, K=64) -- issued early, at the end of the
    //                  previous iteration's P8 (or in the prologue), so it overlaps the W2 re-materialisation
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    TICK(2);

    // ---------------- P2: gelu on D1 row j -> X2^T / X2bar^T (operand format, SW128 rows); keep gelu'(Z1)
    {
      const uint32_t tsrc = tmem + lane_addr + TM_D1 + 128 * half;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if ((c < 2 && !has_k) || (c >= 2 && !has_q)) continue;
        uint32_t v[32];
        tmem_ld32(tsrc + 32 * c, v);
        tc_wait_ld();
        const f32x2 B1 = pk2(b1r);
        if (c < 2) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {  // two tokens per instruction (FFMA2)
            f32x2 g;
            const f32x2 x = gelu_and_grad2(add2(pk2u(v[i], v[i + 1]), B1), g);
            up2u(x, v[i], v[i + 1]);
            gp[16 * c + i / 2] = pack_bf16(g);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) up2u(gelu_only2(add2(pk2u(v[i], v[i + 1]), B1)), v[i], v[i + 1]);
        }
        store_row_op<kF16>(sbase + SM_X2 + (c >> 1) * 32768, j, 4 * (c & 1), v);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    TICK(3);

    // ---------------- P3: D2 = [X2 ; X2bar] . W2b   (M=128 tokens, N=64, K=256 hidden; both operands MN-major)
    if (warp_u == 0 && elect_one()) {
      tc_fence_after();
      const uint64_t da = make_desc_sw128(sbase + SM_X2, 32768, 1024);
      const uint64_t db = make_desc_sw128(sbase + SM_W2B, 1024, 1024);
#pragma unroll
      for (int k = 0; k < 16; ++k)
        umma_ss(tmem + TM_D2, desc_advance(da, 2048 * k), desc_advance(db, 2048 * k), IDESC_B, k > 0);
      tc_commit(mma_bar);
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    TICK(4);

    // ---------------- P4: LayerNorm stage on all 8 warps: thread = (token row, 32-column half); warps w and w+4 share
    //                  TMEM lanes, the row statistics are exchanged through smem (2 floats per exchange)
    {
      const int row = 32 * (warp & 3) + lane;  // 0..63 K side, 64..127 Q side
      const int ch = warp >> 2;                // column half
      const bool kside = row < 64;
      const bool active = kside ? has_k : has_q;
      // packed pairs of adjacent columns throughout (FFMA2); row sums are accumulated per lane of the pair (even / odd
      // columns) and the two lanes added at the end
      f32x2 z[16];
      const f32x2* lw2 = reinterpret_cast<const f32x2*>(lnw + 32 * ch);
      const f32x2* lb2 = reinterpret_cast<const f32x2*>(lnb + 32 * ch);
      float mu = 0.f, rstd = 0.f;
      if (active) {
        uint32_t zr[32];
        tmem_ld32(tmem + lane_addr + TM_D2 + 32 * ch, zr);
        tc_wait_ld();
        const f32x2* b2p = reinterpret_cast<const f32x2*>(b2s + 32 * ch);
        f32x2 a1 = pk2(0.f), a2 = pk2(0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          z[i] = add2(pk2u(zr[2 * i], zr[2 * i + 1]), b2p[i]);
          a1 = add2(a1, z[i]);
          a2 = fma2(z[i], z[i], a2);
        }
        xs1[ch * 128 + row] = make_float2(lo2(a1) + hi2(a1), lo2(a2) + hi2(a2));
      }
      __syncthreads();
      if (active) {
        const float2 p0 = xs1[row], p1 = xs1[128 + row];
        mu = (p0.x + p1.x) * (1.0f / 64.0f);
        rstd = rsqrtf(fmaxf((p0.y + p1.y) * (1.0f / 64.0f) - mu * mu, 0.f) + 1e-8f);
        const f32x2 MU = pk2(mu), RS = pk2(rstd);
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = mul2(sub2(z[i], MU), RS);  // x_hat
      }
      if (kside) {
        // grad_out = gamma*xhat + beta - (V - K); gxh = grad_out*gamma        (K tile in operand format, V tile bf16)
        float g[32];
        f32x2 gq[16];
        if (active) {
          f32x2 s1 = pk2(0.f), s2 = pk2(0.f);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t kk[4], vv[4];
            ld_shared_v4(kq + sw128_off(row, 4 * ch + c), kk[0], kk[1], kk[2], kk[3]);
            ld_shared_v4(vt + sw128_off(row, 4 * ch + c), vv[0], vv[1], vv[2], vv[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * c + e;  // pair index: columns 32 ch + 2 i, + 1
              const f32x2 tgt = sub2(bf16x2_to_f32x2(vv[e]), op_to_f32x2<kF16>(kk[e]));
              gq[i] = mul2(sub2(fma2(lw2[i], z[i], lb2[i]), tgt), lw2[i]);
              s1 = add2(s1, gq[i]);
              s2 = fma2(gq[i], z[i], s2);
            }
          }
          xs2[ch * 64 + row] = make_float2(lo2(s1) + hi2(s1), lo2(s2) + hi2(s2));
        }
        __syncthreads();
        if (active) {
          const float2 p0 = xs2[row], p1 = xs2[64 + row];
          const float s1 = p0.x + p1.x, s2 = p0.y + p1.y;
          // gradZ2 = (64*g - s1 - xhat*s2) / (64*std);  G2 = -eta * gradZ2
          const float eta_i = __uint_as_float((uint32_t)eta_raw << 16);
          const f32x2 SC = pk2(-eta_i * rstd * (1.0f / 64.0f)), NS1 = pk2(-s1), S2 = pk2(s2), C64 = pk2(64.0f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            gq[i] = mul2(sub2(fma2(C64, gq[i], NS1), mul2(z[i], S2)), SC);
            up2(gq[i], g[2 * i], g[2 * i + 1]);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
            st_shared_v4(sbase + SM_G2 + sw128_off(row, 4 * ch + c), pack_op<kF16>(gq[4 * c]), pack_op<kF16>(gq[4 * c + 1]),
                         pack_op<kF16>(gq[4 * c + 2]), pack_op<kF16>(gq[4 * c + 3]));
          // b2 update = column sums of G2 over the 64 token rows (butterfly over the warp, then one shared add per row half:
          // db2acc starts at zero every step and takes exactly two addends, so the sum does not depend on their order)
#pragma unroll
          for (int m = 16; m >= 1; m >>= 1) {
            const bool up = (lane & m) != 0;
#pragma unroll
            for (int q = 0; q < m; ++q) {
              const float a0 = g[q], b0 = g[q + m];
              g[q] = (up ? b0 : a0) + __shfl_xor_sync(0xffffffffu, up ? a0 : b0, m);
            }
          }
          atomicAdd(&db2acc[32 * ch + lane], g[0]);
        }
      } else {
        __syncthreads();  // pairs with the K-side exchange barrier
        if (active) {
          // output: O = Q + gamma*xhat + beta   (mini-batch it-1)
          const int r = row - 64;
          __nv_bfloat16* og = p.Out + (row_base + (size_t)(it - 1) * CS + r) * F + 32 * ch;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t qq[4], o[4];
            ld_shared_v4(kq + 8192 + sw128_off(r, 4 * ch + c), qq[0], qq[1], qq[2], qq[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * c + e;
              o[e] = pack_bf16(add2(op_to_f32x2<kF16>(qq[e]), fma2(lw2[i], z[i], lb2[i])));
            }
            *reinterpret_cast<uint4*>(og + 8 * c) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
    if (!has_k) break;  // last iteration only drains the Q side
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    TICK(5);

    // ---------------- P5: D3[h] = W2b[h] . G2^T  (critical) ;  W2[h] += X2^T[h] . G2  (off the critical path)
    if (warp_u == 0 && elect_one()) {
      tc_fence_after();
      const uint64_t dg_k = make_desc_sw128(sbase + SM_G2, 16, 1024);     // B K-major view  (N = token, K = F)
      const uint64_t dg_mn = make_desc_sw128(sbase + SM_G2, 1024, 1024);  // B MN-major view (K = token, N = F)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t da = make_desc_sw128(sbase + SM_W2B + h * 16384, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem + TM_D3 + 128 * h, desc_advance(da, 32 * k), desc_advance(dg_k, 32 * k), IDESC_C, k > 0);
      }
      tc_commit(mma_bar);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t da = make_desc_sw128(sbase + SM_X2 + h * 16384, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem + TM_W2 + 64 * h, desc_advance(da, 32 * k), desc_advance(dg_mn, 2048 * k), IDESC_U2, 1);
      }
    }
    if (kF16) {  // in the shadow of the D3 batch: next iteration's token tiles (TMA issued at the top of this iteration) bf16 ->
                 // f16 in place; the proxy fences + barriers of P6 / P8 order these writes before the D1 issue that reads them
      const int ns = slot ^ 1;
      mbar_wait(&tma_bar[ns], ((it + 1) >> 1) & 1);
      if (it + 1 < NC) tile_bf16_to_f16(sbase + SM_KQ + ns * 16384, tid);  // K_{it+1}
      tile_bf16_to_f16(sbase + SM_KQ + ns * 16384 + 8192, tid);             // Q_{it}
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    TICK(6);

    // ---------------- P6: G1^T row j = D3 row j * gelu'(Z1) ; b1 += sum ; threads 0-63: b2 += column sums of G2
    {
      const uint32_t tsrc = tmem + lane_addr + TM_D3 + 128 * half;
      f32x2 acc = pk2(0.f);  // even / odd tokens
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tsrc + 32 * c, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const f32x2 a = mul2(pk2u(v[i], v[i + 1]), bf16x2_to_f32x2(gp[16 * c + i / 2]));
          acc = add2(acc, a);
          up2u(a, v[i], v[i + 1]);
        }
        store_row_op<kF16>(sbase + SM_X2 + 32768, j, 4 * c, v);
      }
      b1r += lo2(acc) + hi2(acc);
      if (tid < 64) {  // fold the column sums of G2 gathered in P4 into b2
        b2s[tid] += db2acc[tid];
        db2acc[tid] = 0.f;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    TICK(7);

    // ---------------- P7: W1^T[h] += G1^T[h] . K
    if (warp_u == 0 && elect_one()) {
      tc_fence_after();
      const uint64_t db = make_desc_sw128(kq, 1024, 1024);  // K tile, MN-major view (K = token, N = F)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t da = make_desc_sw128(sbase + SM_X2 + 32768 + h * 16384, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem + TM_W1 + 64 * h, desc_advance(da, 32 * k), desc_advance(db, 2048 * k), IDESC_U, 1);
      }
      tc_commit(mma_bar);  // also covers the W2 update issued in P5
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    TICK(8);

    // ---------------- P8: re-materialise the operand copies of the new state (+ checkpoint / final state)
    {
      const int nstep = it + 1;  // state now equals the state entering mini-batch nstep
      const bool ck = (p.W1c != nullptr) && (nstep < NC) && (nstep % p.ckpt_group == 0);
      const bool fin = (p.W1o != nullptr) && (nstep == NC);
      const size_t kidx = ck ? ((size_t)bh * p.K + nstep / p.ckpt_group) : 0;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + TM_W1 + 64 * half + 32 * c, v);
        tc_wait_ld();
        store_row_op<kF16>(sbase + SM_W1B, j, 4 * c, v);
        if (ck) store_w1_col(p.W1c + kidx * F * HID, j, v, 32 * c);
        if (fin) store_w1_col(p.W1o + (size_t)bh * F * HID, j, v, 32 * c);
      }
      // W1b is complete: start the next iteration's D1 now, it runs under the W2 conversion below
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (warp_u == 0 && elect_one()) issue_p1(it + 1);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + TM_W2 + 64 * half + 32 * c, v);
        tc_wait_ld();
        store_row_op<kF16>(sbase + SM_W2B, j, 4 * c, v);
        if (ck) store_w2_row(p.W2c + kidx * HID * F, j, v, 32 * c);
        if (fin) store_w2_row(p.W2o + (size_t)bh * HID * F, j, v, 32 * c);
      }
      if (ck) {
        p.b1c[kidx * HID + j] = b1r;
        if (tid < 64) p.b2c[kidx * F + tid] = b2s[tid];  // b2s was updated by the same thread in P6
      }
      if (fin) {
        p.b1o[(size_t)bh * HID + j] = b1r;
        if (tid < 64) p.b2o[(size_t)bh * F + tid] = b2s[tid];
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    TICK(9);
  }
  TICK_DUMP(12, p.dbg);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------ host
static int g_bf16_operands = -1;  // -1: read TTT_B200_OPERANDS from the environment once ("bf16" selects the A/B variant)
static bool use_bf16_operands() {
  if (g_bf16_operands < 0) {
    const char* e = getenv("TTT_B200_OPERANDS");
    g_bf16_operands = (e && e[0] == 'b') ? 1 : 0;
  }
  return g_bf16_operands != 0;
}
bool mlp_operands_bf16() { return use_bf16_operands(); }

cudaError_t launch_mlp_forward(const void* XQ, const void* XK, const void* XV, const void* last_eta, const float* ln_w,
                               const float* ln_b, const float* W1, const float* b1, const float* W2, const float* b2,
                               float* W1c, float* b1c, float* W2c, float* b2c, float* W1o, float* b1o, float* W2o,
                               float* b2o, void* Out, int B, int H, int NC, int ckpt_group, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || NC <= 0 || ckpt_group <= 0) return cudaErrorInvalidValue;
  FwdParams p{};
  p.last_eta = reinterpret_cast<const __nv_bfloat16*>(last_eta);
  p.ln_w = ln_w; p.ln_b = ln_b;
  p.W1 = W1; p.b1 = b1; p.W2 = W2; p.b2 = b2;
  p.W1c = W1c; p.b1c = b1c; p.W2c = W2c; p.b2c = b2c;
  p.W1o = W1o; p.b1o = b1o; p.W2o = W2o; p.b2o = b2o;
  p.Out = reinterpret_cast<__nv_bfloat16*>(Out);
  p.B = B; p.H = H; p.NC = NC; p.ckpt_group = ckpt_group;
  p.K = (NC + ckpt_group - 1) / ckpt_group;
  CUtensorMap tq, tk, tv;
  const uint64_t rows = (uint64_t)B * H * NC * CS;
  if (rows > 0x7FFFFFFFull) return cudaErrorInvalidValue;
  if (make_token_tmap(&tq, XQ, rows) || make_token_tmap(&tk, XK, rows) || make_token_tmap(&tv, XV, rows))
    return cudaErrorInvalidValue;
  g_where = "forward launch";
  static bool attr_done_dev[64] = {};  // function attributes are per device
  bool& attr_done = *device_once(attr_done_dev);
  if (!attr_done) {
    TB_TRY(cudaFuncSetAttribute(ttt_mlp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL), "smem attr");
    TB_TRY(cudaFuncSetAttribute(ttt_mlp_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL), "smem attr");
    attr_done = true;
  }
  p.dbg = g_timing_buf;
  if (use_bf16_operands()) ttt_mlp_fwd_kernel<false><<<B * H, NT, SM_TOTAL, stream>>>(tq, tk, tv, p);
  else                     ttt_mlp_fwd_kernel<true><<<B * H, NT, SM_TOTAL, stream>>>(tq, tk, tv, p);
  return cudaGetLastError();
}

}  // namespace tb
