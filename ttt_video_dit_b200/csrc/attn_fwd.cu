// Non-causal local (per-segment) self-attention forward for sm_100a, head_dim 64.
// Replaces the library call F.scaled_dot_product_attention(q, k, v, is_causal=False) of
// ttt/models/cogvideo/dit.py:196-198 (inside _attn_forward, dit.py:163-211) for the segment shapes of the DiT
// ([B, 48, ~18 k, 64]).  Layout: q/k/v/out are bf16 [B, T, H, 64] -- i.e. the *un-rearranged* output of the q/k/v
// Linears ("b t (h d)"), read through 4-D TMA maps, so the reference's rearrange "b t (h d) -> b h t d" and its inverse
// cost nothing.
//
// One CTA = 128 query rows of one (b,h); FlashAttention-style online softmax over 128-key tiles:
//   S[128x128] = Q.K^T            tcgen05, A = Q tile (K-major), B = K tile (K-major), fp32 in TMEM (128 columns)
//   softmax row-wise              two threads per query row (64 key columns each; warps w and w+4 share a TMEM lane
//                                 quadrant), row maximum exchanged through smem; exp2 with log2e-prescale
//   O[128x64] += P.V              A = P (bf16, written back to TMEM over the consumed S columns: "TS" MMA form),
//                                 B = V tile (MN-major view of the [key][d] tile)
// O lives in TMEM and is rescaled in place when the running maximum moves.  K/V tiles are double-buffered by TMA.
// 112 KB smem + 256 TMEM columns per CTA -> two CTAs (16 warps) per SM overlap each other's MMA and softmax phases.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ptx.cuh"
#include "ttt_internal.h"

#ifndef ATTN_P_IN_TMEM
#define ATTN_P_IN_TMEM 1  // 1: P stays in TMEM as the A operand of the PV MMA (TS form); 0: P through a smem tile
#endif

namespace tb {
namespace attn {

constexpr int D = 64, BM = 128, BN = 128, NT = 256;
constexpr uint32_t SM_Q = 0;                  // [128][64]           16 KB
constexpr uint32_t SM_K = 16384;              // 2 x [128][64]       32 KB
constexpr uint32_t SM_V = SM_K + 32768;       // 2 x [128][64]       32 KB
constexpr uint32_t SM_P = SM_V + 32768;       // 2 blocks [128][64]  32 KB  (keys 0-63 | 64-127)
constexpr uint32_t SM_MISC = SM_P + 32768;    // barriers
constexpr uint32_t SM_TOTAL = SM_MISC + 256;
constexpr uint32_t TM_S = 0, TM_O = 128, TM_X = 192;  // TM_X + ch: scalar exchange cells of the two threads of a row

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

__global__ void __launch_bounds__(NT, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ Out, float* __restrict__ lse2,
                int T, int H, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = uniform_warp_id();    // same value as warp, provably warp-uniform: single-thread issue blocks branch on it
  const int row = 32 * (warp & 3) + lane;  // query row of this thread == TMEM lane
  const int ch = warp >> 2;                // which 64 key columns of the S tile / which 32 columns of O
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  // No out-of-bounds TMA boxes: the last query / key tile is shifted back to end exactly at T (T >= 128 is required);
  // rows it shares with the previous tile are masked (keys) or recomputed identically (queries).
  const int q0 = min(qt * BM, T - BM);
  const int ntiles = (T + BN - 1) / BN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_MISC);
  uint64_t* bar_q = bars;        // Q tile
  uint64_t* bar_kv = bars + 1;   // [2]
  uint64_t* mma_bar = bars + 3;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

  if (tid == 0) {
    mbar_init(bar_q, 1);
    mbar_init(&bar_kv[0], 1);
    mbar_init(&bar_kv[1], 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
  }
  if (warp == 0) tmem_alloc<256>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;

  if (warp_u == 0 && elect_one()) {
    mbar_expect_tx(bar_q, 16384);
    tma_load_4d(smem + SM_Q, &tmQ, 0, h, q0, b, bar_q);
    mbar_expect_tx(&bar_kv[0], 32768);
    tma_load_4d(smem + SM_K, &tmK, 0, h, 0, b, &bar_kv[0]);
    tma_load_4d(smem + SM_V, &tmV, 0, h, 0, b, &bar_kv[0]);
  }

  constexpr uint32_t IDESC_S = make_idesc_bf16(128, 128, false, false);
  constexpr uint32_t IDESC_O = make_idesc_bf16(128, 64, false, true);
  float m_run = -INFINITY, l_run = 0.f;
  uint32_t mma_phase = 0;
  auto issue_s = [&](int jn) {  // S = Q . K_jn^T -> TM_S (thread 0, after the K tile of jn has landed)
    const uint64_t da = make_desc_sw128(sbase + SM_Q, 16, 1024);
    const uint64_t db = make_desc_sw128(sbase + SM_K + (jn & 1) * 16384, 16, 1024);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ss(tmem + TM_S, desc_advance(da, 32 * k), desc_advance(db, 32 * k), IDESC_S, k > 0);
  };
  if (warp_u == 0 && elect_one()) {
    mbar_wait(bar_q, 0);
    mbar_wait(&bar_kv[0], 0);
    tc_fence_after();
    issue_s(0);
    tc_commit(mma_bar);
  }

  for (int j = 0; j < ntiles; ++j) {
    const int slot = j & 1;
    // one MMA batch per tile: [P.V of tile j-1, S of tile j] (issued at the end of the previous iteration)
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    if (warp_u == 0 && j + 1 < ntiles && elect_one()) {  // the other K/V slot was released by the PV MMA of tile j-1
      mbar_expect_tx(&bar_kv[slot ^ 1], 32768);
      const int kb = min((j + 1) * BN, T - BN);
      tma_load_4d(smem + SM_K + (slot ^ 1) * 16384, &tmK, 0, h, kb, b, &bar_kv[slot ^ 1]);
      tma_load_4d(smem + SM_V + (slot ^ 1) * 16384, &tmV, 0, h, kb, b, &bar_kv[slot ^ 1]);
    }

    // ---- online softmax for this thread's query row
    const int kbase = min(j * BN, T - BN);
    const int first_new = j * BN - kbase;  // keys below this index were already consumed by the previous tile
    // (only a shifted-back tail tile needs the key masks; the row maximum is taken on the raw scores, scale > 0)
    float mraw = -INFINITY;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float s[32];
      tmem_ld32(tmem + lane_addr + TM_S + 64 * ch + 32 * c, reinterpret_cast<uint32_t*>(s));
      tc_wait_ld();
      if (first_new == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mraw = fmaxf(mraw, s[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) mraw = fmaxf(mraw, (64 * ch + 32 * c + i >= first_new) ? s[i] : -INFINITY);
      }
    }
    // the two threads of a row sit on the same TMEM lane: exchange the partial maxima through two spare TMEM cells
    tmem_st1(tmem + lane_addr + TM_X + ch, __float_as_uint(mraw));
    tc_wait_st();
    tc_fence_before();
    asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");  // only the two warps that share these rows
    tc_fence_after();
    {
      const uint32_t other = tmem_ld1(tmem + lane_addr + TM_X + (ch ^ 1));
      tc_wait_ld();
      mraw = fmaxf(mraw, __uint_as_float(other));
    }
    const float mx = fmaxf(m_run, mraw * scale_log2);
    const float alpha = ex2(m_run - mx);  // m_run = -inf on the first tile -> alpha = 0 (O, l start at 0 anyway)
    float lsum = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float s[32];
      tmem_ld32(tmem + lane_addr + TM_S + 64 * ch + 32 * c, reinterpret_cast<uint32_t*>(s));
      tc_wait_ld();
      if (first_new == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s[i] = ex2(fmaf(s[i], scale_log2, -mx));
          lsum += s[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s[i] = (64 * ch + 32 * c + i >= first_new) ? ex2(fmaf(s[i], scale_log2, -mx)) : 0.f;
          lsum += s[i];
        }
      }
#if ATTN_P_IN_TMEM
      // P chunk -> TMEM as the A operand of the PV MMA ("TS" form, pinned by umma self-test mode 8): packed bf16 pairs in
      // the 16 columns this thread has already consumed of its own half of the S tile (TM_S + 64 ch + 16 c)
      {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(s[2 * i], s[2 * i + 1]);
        tmem_st16(tmem + lane_addr + TM_S + 64 * ch + 16 * c, pk);
      }
#else
      // P chunk -> smem (A operand, K-major over keys): block ch, 16-byte chunks 4*c .. +3 of this row
#pragma unroll
      for (int q = 0; q < 4; ++q)
        st_shared_v4(sbase + SM_P + ch * 16384 + sw128_off(row, 4 * c + q), pack_bf16(s[8 * q], s[8 * q + 1]),
                     pack_bf16(s[8 * q + 2], s[8 * q + 3]), pack_bf16(s[8 * q + 4], s[8 * q + 5]), pack_bf16(s[8 * q + 6], s[8 * q + 7]));
#endif
    }
#if ATTN_P_IN_TMEM
    tc_wait_st();
#endif
    l_run = fmaf(l_run, alpha, lsum);  // partial row sum over this thread's key columns (combined in the epilogue)
    // rescale the running output (this thread: 32 of the 64 columns) only when some row of this warp moved its maximum
    // (the PV MMA of tile j-1 has completed: its commit was waited below); after the first tiles the maximum rarely moves
    if (j > 0 && __any_sync(0xffffffffu, mx > m_run)) {
      float o[32];
      tmem_ld32(tmem + lane_addr + TM_O + 32 * ch, reinterpret_cast<uint32_t*>(o));
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] *= alpha;
      tmem_st32(tmem + lane_addr + TM_O + 32 * ch, reinterpret_cast<uint32_t*>(o));
      tc_wait_st();
    }
    m_run = mx;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    // ---- O += P . V_j, then S of the next tile in the same batch (the tensor pipe is in order: the S GEMM overwrites the
    //      S / P columns only after the PV GEMM has read P)
    if (warp_u == 0 && elect_one()) {
      tc_fence_after();
      const uint64_t db = make_desc_sw128(sbase + SM_V + slot * 16384, 1024, 1024);  // MN-major: rows = keys (K), 64 d (N)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
#if ATTN_P_IN_TMEM
        umma_ts(tmem + TM_O, tmem + TM_S + 64 * (k >> 2) + 8 * (k & 3), desc_advance(db, 2048 * k), IDESC_O, (j > 0) || (k > 0));
#else
        const uint64_t da = make_desc_sw128(sbase + SM_P + (k >> 2) * 16384, 16, 1024);
        umma_ss(tmem + TM_O, desc_advance(da, 32 * (k & 3)), desc_advance(db, 2048 * k), IDESC_O, (j > 0) || (k > 0));
#endif
      }
      if (j + 1 < ntiles) {
        mbar_wait(&bar_kv[slot ^ 1], ((j + 1) >> 1) & 1);
        tc_fence_after();
        issue_s(j + 1);
      }
      tc_commit(mma_bar);
    }
  }
  mbar_wait(mma_bar, mma_phase);
  tc_fence_after();

  // ---- epilogue: O / l -> bf16 -> out[b, q0 + row, h, 32 ch .. 32 ch + 31]
  {
    tmem_st1(tmem + lane_addr + TM_X + 2 + ch, __float_as_uint(l_run));  // combine the two partial row sums
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t l_other = tmem_ld1(tmem + lane_addr + TM_X + 2 + (ch ^ 1));
    tc_wait_ld();
    const float l_tot = l_run + __uint_as_float(l_other);
    const float inv = 1.f / l_tot;
    // row statistic for the backward (attn_bwd.cu): log2-domain log-sum-exp, P = exp2(S * scale_log2 - lse2)
    if (lse2 && ch == 0) lse2[((size_t)b * H + h) * T + q0 + row] = m_run + log2f(l_tot);
    __nv_bfloat16* og = Out + (((size_t)b * T + q0 + row) * H + h) * D + 32 * ch;
    float o[32];
    tmem_ld32(tmem + lane_addr + TM_O + 32 * ch, reinterpret_cast<uint32_t*>(o));
    tc_wait_ld();
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(og + 8 * q) =
          make_uint4(pack_bf16(o[8 * q] * inv, o[8 * q + 1] * inv), pack_bf16(o[8 * q + 2] * inv, o[8 * q + 3] * inv),
                     pack_bf16(o[8 * q + 4] * inv, o[8 * q + 5] * inv), pack_bf16(o[8 * q + 6] * inv, o[8 * q + 7] * inv));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace attn

// 4-D bf16 tensor [B][T][H][64] (innermost first: {64, H, T, B}), box {64, 1, box_rows, 1}, 128-B swizzle
int make_bthd_tmap(CUtensorMap* tm, const void* base, int B, int T, int H, int box_rows) {
  static thread_local char detail[160];
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess) { g_where = "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed"; return -1; }
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(ptr);
  cuuint64_t gdim[4] = {64, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstride[3] = {128, (cuuint64_t)H * 128, (cuuint64_t)T * H * 128};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(detail, sizeof(detail), "cuTensorMapEncodeTiled(4d base=%p B=%d T=%d H=%d) -> CUresult %d", base, B, T, H, (int)r);
    g_where = detail;
    return -2;
  }
  return 0;
}

cudaError_t launch_attention_forward(const void* Q, const void* K, const void* V, void* Out, float* lse2, int B, int T,
                                     int H, float scale, cudaStream_t stream) {
  if (B <= 0 || T < attn::BM || H <= 0) { g_where = "bad sizes (T must be >= 128)"; return cudaErrorInvalidValue; }
  CUtensorMap tq, tk, tv;
  if (make_bthd_tmap(&tq, Q, B, T, H, 128) || make_bthd_tmap(&tk, K, B, T, H, 128) || make_bthd_tmap(&tv, V, B, T, H, 128))
    return cudaErrorInvalidValue;
  static bool attr_done_dev[64] = {};  // function attributes (and side streams) are per device
  bool& attr_done = *device_once(attr_done_dev);
  if (!attr_done) {
    TB_TRY(cudaFuncSetAttribute(attn::attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SM_TOTAL), "smem attr");
    attr_done = true;
  }
  g_where = "attention launch";
  dim3 grid((T + attn::BM - 1) / attn::BM, H, B);
  attn::attn_fwd_kernel<<<grid, attn::NT, attn::SM_TOTAL, stream>>>(tq, tk, tv, reinterpret_cast<__nv_bfloat16*>(Out), lse2, T, H,
                                                                     scale * 1.4426950408889634f);
  return cudaGetLastError();
}

}  // namespace tb
