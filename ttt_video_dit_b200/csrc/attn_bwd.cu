// Non-causal local self-attention backward for sm_100a, head_dim 64: the gradient of attn_fwd.cu, i.e. of the
// F.scaled_dot_product_attention(q, k, v, is_causal=False) call at ttt/models/cogvideo/dit.py:196-198 that autograd
// differentiates through the library's flash-attention backward in the reference.  Tensors stay in the Linear-output
// layout [B, T, H, 64] (4-D TMA maps).
//
// Standard recompute formulation with the forward's row statistics (lse2 = log2-domain log-sum-exp per query row):
//   S = Q K^T ; P = exp2(S*scale*log2e - lse2) ; dP = dO V^T ; dS = P * (dP - delta) * scale, delta = rowsum(dO * O)
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K
// Two passes of ONE kernel template, no atomics and no fp32 dQ buffer (7 GEMMs instead of 5, deterministic):
//   mode 0: CTA owns 128 keys (K_j, V_j stationary), streams 64-query sub-tiles, accumulates dK_j, dV_j in TMEM.
//           It works on the TRANSPOSED score tile  S^T = K_j Q_u^T  [128 keys x 64 queries]  so that the key dimension
//           is the MMA M = 128 and P^T / dS^T come out of the element-wise pass already in the [key][query] layout the
//           dV / dK GEMMs want as a K-major A operand; lse2 / delta are then per COLUMN (small smem arrays).
//   mode 1: CTA owns 128 queries (Q_i, dO_i stationary), streams 64-key sub-tiles, accumulates dQ_i in TMEM
//           (S [128 queries x 64 keys], lse2 / delta per row = per thread).
// Both modes: X = OWN0 . STR0^T, Y = OWN1 . STR1^T (N = 64), element-wise pass on [128 x 64] by 256 threads
// (thread = TMEM lane x 32 columns, no reductions); P / dS are written back to TMEM as packed bf16 over the X / Y columns
// the thread has just consumed and feed the accumulate GEMMs as the A operand FROM TMEM ("TS" form; the streamed
// sub-tile is the MN-major B operand) -- no smem round trip for P / dS.  192 / 256 TMEM columns and 65 KB smem per CTA ->
// TWO CTAs per SM overlap each other's MMA batches and element-wise passes.
// Tail sub-tiles are shifted back to end at T (no out-of-bounds TMA boxes, as in the forward); streamed columns they
// share with the previous sub-tile are masked, rows shared between two stationary tiles are computed twice, equally.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ptx.cuh"
#include "ttt_internal.h"

namespace tb {
namespace attnb {

constexpr int D = 64, BT = 128, BS = 64, NT = 256;
constexpr uint32_t SM_OWN0 = 0;                   // stationary tile 0: K_j (mode 0) / Q_i (mode 1)        [128][64] 16 KB
constexpr uint32_t SM_OWN1 = 16384;               // stationary tile 1: V_j (mode 0) / dO_i (mode 1)       16 KB
constexpr uint32_t SM_STR0 = 32768;               // 2 x streamed sub-tile 0: Q_u (mode 0) / K_u (mode 1)  [64][64] 8 KB
constexpr uint32_t SM_STR1 = SM_STR0 + 16384;     // 2 x streamed sub-tile 1: dO_u (mode 0) / V_u (mode 1)
constexpr uint32_t SM_MISC = SM_STR1 + 16384;     // barriers, tmem ptr ; + mode 0: lse2 / delta of the sub-tile [2][2][64]
constexpr uint32_t SM_TOTAL0 = SM_MISC + 256 + 1024;
constexpr uint32_t SM_TOTAL1 = SM_MISC + 256;
constexpr uint32_t TM_X = 0, TM_Y = 64, TM_ACC0 = 128, TM_ACC1 = 192;  // acc0: dK / dQ, acc1: dV

__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// delta[b,h,t] = sum_d dO[b,t,h,d] * O[b,t,h,d]   (one 8-lane group per row, 16-byte loads)
__global__ void attn_delta_kernel(const uint4* __restrict__ dO, const uint4* __restrict__ O, float* __restrict__ delta,
                                  int T, int H, long long rows) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = gid >> 3;  // row index in [B*T*H): (b*T + t)*H + h
  const int part = (int)(gid & 7);
  float acc = 0.f;
  if (row < rows) {
    const uint4 a = dO[row * 8 + part], b = O[row * 8 + part];
    acc = bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) + bf16_hi(a.y) * bf16_hi(b.y) +
          bf16_lo(a.z) * bf16_lo(b.z) + bf16_hi(a.z) * bf16_hi(b.z) + bf16_lo(a.w) * bf16_lo(b.w) + bf16_hi(a.w) * bf16_hi(b.w);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (row < rows && part == 0) {
    const long long bt = row / H;
    const int h = (int)(row - bt * H);
    const long long b = bt / T;
    const int t = (int)(bt - b * T);
    delta[(b * H + h) * T + t] = acc;
  }
}

template <int kMode>
__global__ void __launch_bounds__(NT, 2)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmOwn0, const __grid_constant__ CUtensorMap tmOwn1,
                const __grid_constant__ CUtensorMap tmStr0, const __grid_constant__ CUtensorMap tmStr1,
                const float* __restrict__ lse2, const float* __restrict__ delta, __nv_bfloat16* __restrict__ out0,
                __nv_bfloat16* __restrict__ out1, int T, int H, float scale_log2, float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = uniform_warp_id();  // == warp, provably warp-uniform: single-thread issue blocks branch on it
  const int own = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row = 32 * (warp & 3) + lane;  // row of the stationary tile == TMEM lane
  const int ch = warp >> 2;                // 32-column half of the 64 streamed columns
  const int nsub = (T + BS - 1) / BS;
  const int own0 = min(own * BT, T - BT);  // first token of the stationary tile
  constexpr uint32_t MISC = SM_MISC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MISC);
  uint64_t* bar_own = bars;
  uint64_t* bar_str = bars + 1;  // [2]
  uint64_t* mma_bar = bars + 3;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  float* colstat = reinterpret_cast<float*>(smem + MISC + 256);  // mode 0: [2 buffers][lse2 | delta][64]

  if (warp_u == 0 && elect_one()) {
    mbar_init(bar_own, 1);
    mbar_init(&bar_str[0], 1);
    mbar_init(&bar_str[1], 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmOwn0); tma_prefetch_desc(&tmOwn1); tma_prefetch_desc(&tmStr0); tma_prefetch_desc(&tmStr1);
  }
  if (warp == 0) tmem_alloc<256>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
  const float* lse_bh = lse2 + ((size_t)b * H + h) * T;
  const float* del_bh = delta + ((size_t)b * H + h) * T;

  auto load_stream = [&](int u) {  // streamed sub-tiles of step u -> slot u & 1
    const int s = u & 1, t0 = min(u * BS, T - BS);
    mbar_expect_tx(&bar_str[s], 16384);
    tma_load_4d(smem + SM_STR0 + s * 8192, &tmStr0, 0, h, t0, b, &bar_str[s]);
    tma_load_4d(smem + SM_STR1 + s * 8192, &tmStr1, 0, h, t0, b, &bar_str[s]);
  };
  constexpr uint32_t IDESC_KK = make_idesc_bf16(128, 64, false, false);  // A K-major, B K-major
  constexpr uint32_t IDESC_KN = make_idesc_bf16(128, 64, false, true);   // A K-major, B MN-major (streamed sub-tile, K = its rows)
  auto issue_xy = [&](int u) {  // X = OWN0 . STR0^T -> TM_X ; Y = OWN1 . STR1^T -> TM_Y
    const int s = u & 1;
    const uint64_t a0 = make_desc_sw128(sbase + SM_OWN0, 16, 1024), b0 = make_desc_sw128(sbase + SM_STR0 + s * 8192, 16, 1024);
    const uint64_t a1 = make_desc_sw128(sbase + SM_OWN1, 16, 1024), b1 = make_desc_sw128(sbase + SM_STR1 + s * 8192, 16, 1024);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ss(tmem + TM_X, desc_advance(a0, 32 * k), desc_advance(b0, 32 * k), IDESC_KK, k > 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ss(tmem + TM_Y, desc_advance(a1, 32 * k), desc_advance(b1, 32 * k), IDESC_KK, k > 0);
  };

  if (warp_u == 0 && elect_one()) {
    mbar_expect_tx(bar_own, 32768);
    tma_load_4d(smem + SM_OWN0, &tmOwn0, 0, h, own0, b, bar_own);
    tma_load_4d(smem + SM_OWN1, &tmOwn1, 0, h, own0, b, bar_own);
    load_stream(0);
    mbar_wait(bar_own, 0);
    mbar_wait(&bar_str[0], 0);
    tc_fence_after();
    issue_xy(0);
    tc_commit(mma_bar);
  }
  float lse_r = 0.f, del_r = 0.f;
  if (kMode == 1) {
    lse_r = -lse_bh[own0 + row];
    del_r = -del_bh[own0 + row];
  } else if (tid < 128) {  // column statistics of sub-tile 0
    const int t0 = min(0, T - BS);
    colstat[tid] = -((tid < 64) ? lse_bh[t0 + tid] : del_bh[t0 + tid - 64]);
  }
  __syncthreads();
  uint32_t mma_phase = 0;

  for (int u = 0; u < nsub; ++u) {
    const int s = u & 1;
    const int t0 = min(u * BS, T - BS);
    const int first_new = u * BS - t0;  // streamed columns below this index were in the previous sub-tile
    if (kMode == 0 && u + 1 < nsub && tid < 128) {  // column statistics of the next sub-tile (visible after this step's sync)
      const int tn = min((u + 1) * BS, T - BS);
      colstat[(s ^ 1) * 128 + tid] = -((tid < 64) ? lse_bh[tn + tid] : del_bh[tn + tid - 64]);
    }
    mbar_wait(mma_bar, mma_phase);  // X, Y of this step (and the accumulate GEMMs of the previous one) are done
    mma_phase ^= 1;
    tc_fence_after();
    if (warp_u == 0 && (u + 1 < nsub) && elect_one()) load_stream(u + 1);  // slot (u+1)&1 was last read by step u-1's GEMMs

    // ---- P and dS for (row, 32 columns); only a shifted-back tail sub-tile needs the column masks
    {
      float xv[32], yv[32];
      tmem_ld32(tmem + lane_addr + TM_X + 32 * ch, reinterpret_cast<uint32_t*>(xv));
      tmem_ld32(tmem + lane_addr + TM_Y + 32 * ch, reinterpret_cast<uint32_t*>(yv));
      tc_wait_ld();
      const float* cs = colstat + s * 128 + 32 * ch;  // mode 0: -lse2[q], (+64) -delta[q] of this thread's columns
      if (first_new == 0) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float nl = (kMode == 0) ? cs[e] : lse_r, nd = (kMode == 0) ? cs[64 + e] : del_r;
          xv[e] = ex2(fmaf(xv[e], scale_log2, nl));
          yv[e] = xv[e] * ((yv[e] + nd) * scale);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float nl = (kMode == 0) ? cs[e] : lse_r, nd = (kMode == 0) ? cs[64 + e] : del_r;
          const float pv = (32 * ch + e >= first_new) ? ex2(fmaf(xv[e], scale_log2, nl)) : 0.f;
          xv[e] = pv;
          yv[e] = pv * ((yv[e] + nd) * scale);
        }
      }
      // P / dS -> TMEM as packed bf16 pairs, the A operand of the accumulate GEMMs ("TS" form, umma self-test mode 8): each
      // thread overwrites the first 16 of the 32 X / Y columns it has just consumed (its own columns only, so no
      // cross-thread hazard; the tensor pipe is in order, so the next step's X / Y GEMMs cannot overtake these reads)
      {
        uint32_t pk[16];
        if (kMode == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(xv[2 * i], xv[2 * i + 1]);
          tmem_st16(tmem + lane_addr + TM_X + 32 * ch, pk);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(yv[2 * i], yv[2 * i + 1]);
        tmem_st16(tmem + lane_addr + TM_Y + 32 * ch, pk);
        tc_wait_st();
      }
    }
    tc_fence_before();
    __syncthreads();

    if (warp_u == 0 && elect_one()) {
      tc_fence_after();
      // accumulate: A = P / dS from TMEM (K = the 64 streamed tokens: k-step kb at column 32 (kb >> 1) + 8 (kb & 1)),
      // B = streamed sub-tile [64 tokens][64 d] MN-major
      const uint64_t bs0 = make_desc_sw128(sbase + SM_STR0 + s * 8192, 1024, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k)  // dK += dS^T Q_u  /  dQ += dS K_u
        umma_ts(tmem + TM_ACC0, tmem + TM_Y + 32 * (k >> 1) + 8 * (k & 1), desc_advance(bs0, 2048 * k), IDESC_KN, (u > 0) || (k > 0));
      if (kMode == 0) {
        const uint64_t bs1 = make_desc_sw128(sbase + SM_STR1 + s * 8192, 1024, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dV += P^T dO_u
          umma_ts(tmem + TM_ACC1, tmem + TM_X + 32 * (k >> 1) + 8 * (k & 1), desc_advance(bs1, 2048 * k), IDESC_KN, (u > 0) || (k > 0));
      }
      if (u + 1 < nsub) {
        mbar_wait(&bar_str[(u + 1) & 1], ((u + 1) >> 1) & 1);
        tc_fence_after();
        issue_xy(u + 1);
      }
      tc_commit(mma_bar);
    }
  }
  mbar_wait(mma_bar, mma_phase);
  tc_fence_after();

  // ---- epilogue: accumulators -> bf16 -> [b, own0 + row, h, :]
  {
    // mode 0: warpgroup 0 stores dK (acc0), warpgroup 1 stores dV (acc1), 64 columns each; mode 1: 32 columns of dQ each
    const bool second = (kMode == 0) && (ch == 1);
    __nv_bfloat16* dst = (second ? out1 : out0) + (((size_t)b * T + own0 + row) * H + h) * D;
    const uint32_t src = tmem + lane_addr + (second ? TM_ACC1 : TM_ACC0);
#pragma unroll
    for (int c = 0; c < (kMode == 0 ? 2 : 1); ++c) {
      const int col = (kMode == 0) ? 32 * c : 32 * ch;
      float o[32];
      tmem_ld32(src + col, reinterpret_cast<uint32_t*>(o));
      tc_wait_ld();
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(dst + col + 8 * q) =
            make_uint4(pack_bf16(o[8 * q], o[8 * q + 1]), pack_bf16(o[8 * q + 2], o[8 * q + 3]),
                       pack_bf16(o[8 * q + 4], o[8 * q + 5]), pack_bf16(o[8 * q + 6], o[8 * q + 7]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace attnb

int make_bthd_tmap(CUtensorMap* tm, const void* base, int B, int T, int H, int box_rows);  // attn_fwd.cu

cudaError_t launch_attention_backward(const void* Q, const void* K, const void* V, const void* Out, const void* dOut,
                                      const float* lse2, float* delta, void* dQ, void* dK, void* dV, int B, int T, int H,
                                      float scale, cudaStream_t stream) {
  if (B <= 0 || T < attnb::BT || H <= 0) { g_where = "bad sizes (T must be >= 128)"; return cudaErrorInvalidValue; }
  CUtensorMap q128, k128, v128, do128, q64, k64, v64, do64;
  if (make_bthd_tmap(&q128, Q, B, T, H, 128) || make_bthd_tmap(&k128, K, B, T, H, 128) || make_bthd_tmap(&v128, V, B, T, H, 128) ||
      make_bthd_tmap(&do128, dOut, B, T, H, 128) || make_bthd_tmap(&q64, Q, B, T, H, 64) || make_bthd_tmap(&k64, K, B, T, H, 64) ||
      make_bthd_tmap(&v64, V, B, T, H, 64) || make_bthd_tmap(&do64, dOut, B, T, H, 64))
    return cudaErrorInvalidValue;
  static bool attr_done_dev[64] = {};  // function attributes (and side streams) are per device
  bool& attr_done = *device_once(attr_done_dev);
  if (!attr_done) {
    TB_TRY(cudaFuncSetAttribute(attnb::attn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, attnb::SM_TOTAL0), "smem attr");
    TB_TRY(cudaFuncSetAttribute(attnb::attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, attnb::SM_TOTAL1), "smem attr");
    attr_done = true;
  }
  const long long rows = (long long)B * T * H;
  g_where = "attention delta launch";
  attnb::attn_delta_kernel<<<(unsigned)((rows * 8 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(dOut), reinterpret_cast<const uint4*>(Out), delta, T, H, rows);
  TB_TRY(cudaGetLastError(), "attention delta launch");
  const float sl2 = scale * 1.4426950408889634f;
  dim3 grid((T + attnb::BT - 1) / attnb::BT, H, B);
  g_where = "attention backward launch";
  // mode 0: own = K, V ; streamed = Q, dO       mode 1: own = Q, dO ; streamed = K, V
  attnb::attn_bwd_kernel<0><<<grid, attnb::NT, attnb::SM_TOTAL0, stream>>>(k128, v128, q64, do64, lse2, delta,
                                                                           reinterpret_cast<__nv_bfloat16*>(dK),
                                                                           reinterpret_cast<__nv_bfloat16*>(dV), T, H, sl2, scale);
  TB_TRY(cudaGetLastError(), "attention backward launch (dK, dV)");
  attnb::attn_bwd_kernel<1><<<grid, attnb::NT, attnb::SM_TOTAL1, stream>>>(q128, do128, k64, v64, lse2, delta,
                                                                           reinterpret_cast<__nv_bfloat16*>(dQ), nullptr, T, H, sl2, scale);
  return cudaGetLastError();
}

}  // namespace tb
