// Non-causal local self-attention backward for sm_100a, head_dim 64: the gradient of attn_fwd.cu, i.e. of the
// F.scaled_dot_product_attention(q, k, v, is_causal=False) call at ttt/models/cogvideo/dit.py:196-198 that autograd
// differentiates through the library's flash-attention backward in the reference.  Tensors stay in the Linear-output
// layout [B, T, H, 64] (4-D TMA maps).
//
// Standard recompute formulation with the forward's row statistics (lse2 = log2-domain log-sum-exp per query row):
//   S = Q K^T ; P = exp2(S*scale*log2e - lse2) ; dP = dO V^T ; dS = P * (dP - delta) * scale, delta = rowsum(dO * O)
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K
// Two passes of ONE kernel template, no atomics and no fp32 dQ buffer (7 GEMMs instead of 5, deterministic):
//   mode 0: CTA = one 128-key tile (stationary K_j, V_j), streams the query tiles, accumulates dK_j, dV_j in TMEM
//   mode 1: CTA = one 128-query tile (stationary Q_i, dO_i), streams the key tiles, accumulates dQ_i in TMEM
// In both modes S / dP are [128 query rows x 128 key columns] fp32 in TMEM; 512 threads = (row, 32-column quarter): no row
// reductions are needed in the backward (lse2, delta are inputs), so the four warpgroups simply split the columns
// (16 warps per SM hide the tcgen05.ld / MUFU latencies of the element-wise pass).  P and dS go
// to shared memory as bf16 [query][key] tiles (two 64-column SW128 blocks): read MN-major they are the A operand of the
// dV / dK GEMMs (M = keys, K = queries), read K-major dS is the A operand of the dQ GEMM -- no transposes anywhere.
// The accumulate GEMMs of iteration i and the S / dP GEMMs of iteration i+1 are issued as one batch.
// Tail tiles are shifted back to end at T (no out-of-bounds TMA boxes, as in the forward); rows / columns they share with
// the previous streamed tile are masked, rows shared between two stationary tiles are computed twice with equal results.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ptx.cuh"
#include "ttt_internal.h"

namespace tb {
namespace attnb {

constexpr int D = 64, BT = 128, NT = 512;
constexpr uint32_t SM_FIX0 = 0;                   // stationary tile 0: K_j (mode 0) / Q_i (mode 1)       16 KB
constexpr uint32_t SM_FIX1 = 16384;               // stationary tile 1: V_j (mode 0) / dO_i (mode 1)      16 KB
constexpr uint32_t SM_STR0 = 32768;               // 2 x streamed tile 0: Q_i (mode 0) / K_j (mode 1)     32 KB
constexpr uint32_t SM_STR1 = SM_STR0 + 32768;     // 2 x streamed tile 1: dO_i (mode 0) / V_j (mode 1)    32 KB
constexpr uint32_t SM_P = SM_STR1 + 32768;        // P  [128 q][128 k] bf16, 2 blocks of 64 columns      32 KB
constexpr uint32_t SM_DS = SM_P + 32768;          // dS                                                    32 KB
constexpr uint32_t SM_MISC = SM_DS + 32768;
constexpr uint32_t SM_TOTAL = SM_MISC + 256;
constexpr uint32_t TM_S = 0, TM_DP = 128, TM_ACC0 = 256, TM_ACC1 = 320;  // acc0: dK / dQ, acc1: dV

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
__global__ void __launch_bounds__(NT, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const float* __restrict__ lse2, const float* __restrict__ delta, __nv_bfloat16* __restrict__ out0,
                __nv_bfloat16* __restrict__ out1, int T, int H, float scale_log2, float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int own = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row = 32 * (warp & 3) + lane;  // query row of the S tile == TMEM lane
  const int ch = warp >> 2;                // 32-column quarter of the S tile
  const int ntiles = (T + BT - 1) / BT;
  const int own0 = min(own * BT, T - BT);  // first token of the stationary tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_MISC);
  uint64_t* bar_fix = bars;
  uint64_t* bar_str = bars + 1;  // [2]
  uint64_t* mma_bar = bars + 3;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

  if (tid == 0) {
    mbar_init(bar_fix, 1);
    mbar_init(&bar_str[0], 1);
    mbar_init(&bar_str[1], 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
  const float* lse_bh = lse2 + ((size_t)b * H + h) * T;
  const float* del_bh = delta + ((size_t)b * H + h) * T;

  auto load_stream = [&](int i) {  // streamed tiles of iteration i -> slot i & 1
    const int s = i & 1, t0 = min(i * BT, T - BT);
    mbar_expect_tx(&bar_str[s], 32768);
    tma_load_4d(smem + SM_STR0 + s * 16384, kMode == 0 ? &tmQ : &tmK, 0, h, t0, b, &bar_str[s]);
    tma_load_4d(smem + SM_STR1 + s * 16384, kMode == 0 ? &tmDO : &tmV, 0, h, t0, b, &bar_str[s]);
  };
  constexpr uint32_t IDESC_S = make_idesc_bf16(128, 128, false, false);  // A K-major, B K-major, N = 128
  constexpr uint32_t IDESC_T = make_idesc_bf16(128, 64, true, true);     // A MN-major ([q][k] tile read as k x q), B MN-major
  constexpr uint32_t IDESC_Q = make_idesc_bf16(128, 64, false, true);    // A K-major (dS), B MN-major (K tile)
  auto issue_s_dp = [&](int i) {  // S = Q K^T -> TM_S ; dP = dO V^T -> TM_DP
    const int s = i & 1;
    const uint32_t q_t = kMode == 0 ? sbase + SM_STR0 + s * 16384 : sbase + SM_FIX0;
    const uint32_t k_t = kMode == 0 ? sbase + SM_FIX0 : sbase + SM_STR0 + s * 16384;
    const uint32_t do_t = kMode == 0 ? sbase + SM_STR1 + s * 16384 : sbase + SM_FIX1;
    const uint32_t v_t = kMode == 0 ? sbase + SM_FIX1 : sbase + SM_STR1 + s * 16384;
    const uint64_t dq = make_desc_sw128(q_t, 16, 1024), dk = make_desc_sw128(k_t, 16, 1024);
    const uint64_t dd = make_desc_sw128(do_t, 16, 1024), dv = make_desc_sw128(v_t, 16, 1024);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ss(tmem + TM_S, desc_advance(dq, 32 * k), desc_advance(dk, 32 * k), IDESC_S, k > 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ss(tmem + TM_DP, desc_advance(dd, 32 * k), desc_advance(dv, 32 * k), IDESC_S, k > 0);
  };

  if (tid == 0) {
    mbar_expect_tx(bar_fix, 32768);
    tma_load_4d(smem + SM_FIX0, kMode == 0 ? &tmK : &tmQ, 0, h, own0, b, bar_fix);
    tma_load_4d(smem + SM_FIX1, kMode == 0 ? &tmV : &tmDO, 0, h, own0, b, bar_fix);
    load_stream(0);
    mbar_wait(bar_fix, 0);
    mbar_wait(&bar_str[0], 0);
    tc_fence_after();
    issue_s_dp(0);
    tc_commit(mma_bar);
  }
  float lse_r = 0.f, del_r = 0.f;
  if (kMode == 1) { lse_r = lse_bh[own0 + row]; del_r = del_bh[own0 + row]; }
  uint32_t mma_phase = 0;

  for (int i = 0; i < ntiles; ++i) {
    const int s = i & 1;
    const int t0 = min(i * BT, T - BT);
    const int first_new = i * BT - t0;  // streamed rows (mode 0) / columns (mode 1) below this were in the previous tile
    if (kMode == 0) { lse_r = lse_bh[t0 + row]; del_r = del_bh[t0 + row]; }
    mbar_wait(mma_bar, mma_phase);  // S, dP of this iteration (and the accumulate GEMMs of the previous one) are done
    mma_phase ^= 1;
    tc_fence_after();
    if (tid == 0 && i + 1 < ntiles) load_stream(i + 1);  // slot (i+1)&1 was last read by iteration i-1's GEMMs

    // ---- P and dS for (row, column half ch); only a shifted-back tail tile needs the element masks
    const bool row_ok = (kMode == 1) || (row >= first_new);
    const float nlse = -lse_r, ndel = -del_r;
    {
      float sv[32], dp[32];
      tmem_ld32(tmem + lane_addr + TM_S + 32 * ch, reinterpret_cast<uint32_t*>(sv));
      tmem_ld32(tmem + lane_addr + TM_DP + 32 * ch, reinterpret_cast<uint32_t*>(dp));
      tc_wait_ld();
      if (first_new == 0) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          sv[e] = ex2(fmaf(sv[e], scale_log2, nlse));
          dp[e] = sv[e] * ((dp[e] + ndel) * scale);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const bool ok = row_ok && ((kMode == 0) || (32 * ch + e >= first_new));
          const float pv = ok ? ex2(fmaf(sv[e], scale_log2, nlse)) : 0.f;
          sv[e] = pv;
          dp[e] = pv * ((dp[e] + ndel) * scale);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t off = (uint32_t)(ch >> 1) * 16384 + sw128_off(row, 4 * (ch & 1) + q);
        if (kMode == 0)
          st_shared_v4(sbase + SM_P + off, pack_bf16(sv[8 * q], sv[8 * q + 1]), pack_bf16(sv[8 * q + 2], sv[8 * q + 3]),
                       pack_bf16(sv[8 * q + 4], sv[8 * q + 5]), pack_bf16(sv[8 * q + 6], sv[8 * q + 7]));
        st_shared_v4(sbase + SM_DS + off, pack_bf16(dp[8 * q], dp[8 * q + 1]), pack_bf16(dp[8 * q + 2], dp[8 * q + 3]),
                     pack_bf16(dp[8 * q + 4], dp[8 * q + 5]), pack_bf16(dp[8 * q + 6], dp[8 * q + 7]));
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
      if (kMode == 0) {
        // dV += P^T dO_i ; dK += dS^T Q_i   (A = [q][k] tile read MN-major: M = 128 keys = 2 blocks, K = 128 query rows)
        const uint64_t ap = make_desc_sw128(sbase + SM_P, 16384, 1024), ads = make_desc_sw128(sbase + SM_DS, 16384, 1024);
        const uint64_t bdo = make_desc_sw128(sbase + SM_STR1 + s * 16384, 1024, 1024);
        const uint64_t bq = make_desc_sw128(sbase + SM_STR0 + s * 16384, 1024, 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem + TM_ACC1, desc_advance(ap, 2048 * k), desc_advance(bdo, 2048 * k), IDESC_T, (i > 0) || (k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem + TM_ACC0, desc_advance(ads, 2048 * k), desc_advance(bq, 2048 * k), IDESC_T, (i > 0) || (k > 0));
      } else {
        // dQ += dS K_j   (A = dS K-major over the 128 keys = 2 blocks x 4 k-steps, B = K tile MN-major)
        const uint64_t bk = make_desc_sw128(sbase + SM_STR0 + s * 16384, 1024, 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t ads = make_desc_sw128(sbase + SM_DS + (k >> 2) * 16384, 16, 1024);
          umma_ss(tmem + TM_ACC0, desc_advance(ads, 32 * (k & 3)), desc_advance(bk, 2048 * k), IDESC_Q, (i > 0) || (k > 0));
        }
      }
      if (i + 1 < ntiles) {
        mbar_wait(&bar_str[(i + 1) & 1], ((i + 1) >> 1) & 1);
        tc_fence_after();
        issue_s_dp(i + 1);
      }
      tc_commit(mma_bar);
    }
  }
  mbar_wait(mma_bar, mma_phase);
  tc_fence_after();

  // ---- epilogue: accumulators -> bf16 -> [b, own0 + row, h, :]
  {
    // mode 0: warpgroups 0,1 store the two 32-column halves of dK (acc0), warpgroups 2,3 those of dV (acc1);
    // mode 1: warpgroups 0,1 store dQ (acc0)
    const bool second = (kMode == 0) && (ch >= 2);
    if (kMode == 0 || ch < 2) {
      __nv_bfloat16* dst = (second ? out1 : out0) + (((size_t)b * T + own0 + row) * H + h) * D;
      const int col = 32 * (ch & 1);
      float o[32];
      tmem_ld32(tmem + lane_addr + (second ? TM_ACC1 : TM_ACC0) + col, reinterpret_cast<uint32_t*>(o));
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
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace attnb

int make_bthd_tmap(CUtensorMap* tm, const void* base, int B, int T, int H);  // attn_fwd.cu

cudaError_t launch_attention_backward(const void* Q, const void* K, const void* V, const void* Out, const void* dOut,
                                      const float* lse2, float* delta, void* dQ, void* dK, void* dV, int B, int T, int H,
                                      float scale, cudaStream_t stream) {
  if (B <= 0 || T < attnb::BT || H <= 0) { g_where = "bad sizes (T must be >= 128)"; return cudaErrorInvalidValue; }
  CUtensorMap tq, tk, tv, tdo;
  if (make_bthd_tmap(&tq, Q, B, T, H) || make_bthd_tmap(&tk, K, B, T, H) || make_bthd_tmap(&tv, V, B, T, H) ||
      make_bthd_tmap(&tdo, dOut, B, T, H))
    return cudaErrorInvalidValue;
  static bool attr_done = false;
  if (!attr_done) {
    TB_TRY(cudaFuncSetAttribute(attnb::attn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, attnb::SM_TOTAL), "smem attr");
    TB_TRY(cudaFuncSetAttribute(attnb::attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, attnb::SM_TOTAL), "smem attr");
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
  attnb::attn_bwd_kernel<0><<<grid, attnb::NT, attnb::SM_TOTAL, stream>>>(tq, tk, tv, tdo, lse2, delta,
                                                                          reinterpret_cast<__nv_bfloat16*>(dK),
                                                                          reinterpret_cast<__nv_bfloat16*>(dV), T, H, sl2, scale);
  TB_TRY(cudaGetLastError(), "attention backward launch (dK, dV)");
  attnb::attn_bwd_kernel<1><<<grid, attnb::NT, attnb::SM_TOTAL, stream>>>(tq, tk, tv, tdo, lse2, delta,
                                                                          reinterpret_cast<__nv_bfloat16*>(dQ), nullptr, T, H, sl2, scale);
  return cudaGetLastError();
}

}  // namespace tb
