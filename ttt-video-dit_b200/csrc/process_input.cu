// Input preparation of the TTT op after the q/k/v Linears, one HBM pass.  Replaces the chain of torch ops in
// TTTBase.process_input (ttt/models/ssm/ttt_layer.py:252-306): L2-normalise q, k over the head dim
// (:265-266), RoPE with global positions on the video tokens (:271-276, ssm/utils.py:82-108: complex multiply on
// interleaved pairs), XV <- LN_unbiased-std(XV - XK) * gamma + beta + XK (:219-235, eps added to the std), the
// [B,L,H,F] -> [B,H,NC,CS,F] transpose (:237-250), the multi-scene interleave (:157-189, as a gather index) and
// eta = base_lr * sigmoid(X.w_h + b_h) / F / CS (:143-155,287-288).  eta is produced as the single row the scan reads
// ([B,H,NC,CS] instead of the reference's materialised [B,H,NC,CS,CS], 2.2 GB at 63 s): the reference repeats the row
// BEFORE interleaving, so the last row of an interleaved mini-batch n is the lr vector of the SOURCE mini-batch that holds
// the source token of position n*CS + CS-1 -- reproduced here so that the op sees exactly the reference's numbers.
//
// One warp per (token, head) row: lane = one interleaved RoPE pair (2 of the 64 features), row reductions by shuffles;
// a CTA = 8 consecutive destination tokens of one head, so the [B,H,L,F] stores are 1 KB contiguous.  HBM-bound:
// 3 x 128 B in + 3 x 128 B out per row.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"
#include "ttt_internal.h"

namespace tb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

__global__ void __launch_bounds__(256)
ttt_process_input_kernel(const uint32_t* __restrict__ xq, const uint32_t* __restrict__ xk, const uint32_t* __restrict__ xv,
                         const float* __restrict__ lr_logit, const float* __restrict__ cosT, const float* __restrict__ sinT,
                         const float* __restrict__ ln_w, const float* __restrict__ ln_b, const int* __restrict__ index,
                         uint32_t* __restrict__ oq, uint32_t* __restrict__ ok, uint32_t* __restrict__ ov,
                         __nv_bfloat16* __restrict__ oeta, int L, int H, int seq_text, int CS, float eta_scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = blockIdx.x * 8 + warp, h = blockIdx.y, b = blockIdx.z;
  if (l >= L) return;
  const int src = index ? index[l] : l;
  const size_t in = (((size_t)b * L + src) * H + h) * 32 + lane;
  const uint32_t qp = xq[in], kp = xk[in], vp = xv[in];
  float q0 = bf16_lo(qp), q1 = bf16_hi(qp), k0 = bf16_lo(kp), k1 = bf16_hi(kp), v0 = bf16_lo(vp), v1 = bf16_hi(vp);
  const float qn = 1.f / fmaxf(sqrtf(warp_sum(q0 * q0 + q1 * q1)), 1e-12f);  // F.normalize eps
  const float kn = 1.f / fmaxf(sqrtf(warp_sum(k0 * k0 + k1 * k1)), 1e-12f);
  q0 *= qn; q1 *= qn; k0 *= kn; k1 *= kn;
  if (src >= seq_text) {  // video token: rotate the pair by the angle of its global video position
    const float c = cosT[(size_t)(src - seq_text) * 32 + lane], s = sinT[(size_t)(src - seq_text) * 32 + lane];
    const float a = q0 * c - q1 * s, bq = q0 * s + q1 * c;
    q0 = a; q1 = bq;
    const float e = k0 * c - k1 * s, f = k0 * s + k1 * c;
    k0 = e; k1 = f;
  }
  // reconstruction target: LayerNorm with the unbiased std of (XV - XK), eps added to the std
  float d0 = v0 - k0, d1 = v1 - k1;
  const float mean = warp_sum(d0 + d1) * (1.f / 64.f);
  d0 -= mean; d1 -= mean;
  const float inv = 1.f / (sqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.f / 63.f)) + 1e-8f);
  v0 = fmaf(ln_w[h * 64 + 2 * lane], d0 * inv, ln_b[h * 64 + 2 * lane]) + k0;
  v1 = fmaf(ln_w[h * 64 + 2 * lane + 1], d1 * inv, ln_b[h * 64 + 2 * lane + 1]) + k1;
  const size_t out = (((size_t)b * H + h) * L + l) * 32 + lane;
  oq[out] = pack_bf16(q0, q1);
  ok[out] = pack_bf16(k0, k1);
  ov[out] = pack_bf16(v0, v1);
  if (lane == 0) {  // eta of (mini-batch n, column j): lr of token j of the source mini-batch of this mini-batch's last row
    const int n = l / CS, j = l - n * CS;
    const int last = n * CS + CS - 1;
    const int src_mb = (index ? index[last] : last) / CS;
    const float z = lr_logit[((size_t)b * L + (size_t)src_mb * CS + j) * H + h];
    oeta[((size_t)b * H + h) * L + l] = __float2bfloat16(eta_scale / (1.f + __expf(-z)));
  }
}

cudaError_t launch_process_input(const void* xq, const void* xk, const void* xv, const float* lr_logit, const float* cosT,
                                 const float* sinT, const float* ln_w, const float* ln_b, const int* index, void* XQ,
                                 void* XK, void* XV, void* last_eta, int B, int L, int H, int seq_text_length, int mini_batch,
                                 float base_lr, cudaStream_t stream) {
  if (B <= 0 || L <= 0 || H <= 0 || mini_batch <= 0 || L % mini_batch != 0 || seq_text_length < 0 || seq_text_length > L) {
    g_where = "bad sizes (L must be a multiple of the mini-batch size)";
    return cudaErrorInvalidValue;
  }
  g_where = "process_input launch";
  dim3 grid((L + 7) / 8, H, B);
  ttt_process_input_kernel<<<grid, 256, 0, stream>>>(
      reinterpret_cast<const uint32_t*>(xq), reinterpret_cast<const uint32_t*>(xk), reinterpret_cast<const uint32_t*>(xv),
      lr_logit, cosT, sinT, ln_w, ln_b, index, reinterpret_cast<uint32_t*>(XQ), reinterpret_cast<uint32_t*>(XK),
      reinterpret_cast<uint32_t*>(XV), reinterpret_cast<__nv_bfloat16*>(last_eta), L, H, seq_text_length, mini_batch,
      base_lr / 64.f / (float)mini_batch);
  return cudaGetLastError();
}

}  // namespace tb
