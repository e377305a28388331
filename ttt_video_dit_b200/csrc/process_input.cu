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
// Half a warp per (token, head) row: lane = two interleaved RoPE pairs (4 of the 64 features, one 8-byte load per tensor),
// row reductions by 4 shuffles; every warp handles 4 rows and issues all of its loads before the first reduction (enough
// bytes in flight to cover the HBM latency); a CTA = 32 consecutive destination tokens of one head, so the [B,H,L,F]
// stores are 4 KB contiguous.  HBM-bound: 3 x 128 B in + 3 x 128 B out per row.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"
#include "ttt_internal.h"

namespace tb {

__device__ __forceinline__ float half_warp_sum(float v) {  // over the 16 lanes that share a row
#pragma unroll
  for (int m = 8; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

__global__ void __launch_bounds__(256)
ttt_process_input_kernel(const uint2* __restrict__ xq, const uint2* __restrict__ xk, const uint2* __restrict__ xv,
                         const float* __restrict__ lr_logit, const float2* __restrict__ cosT, const float2* __restrict__ sinT,
                         const float* __restrict__ ln_w, const float* __restrict__ ln_b, const int* __restrict__ index,
                         uint2* __restrict__ oq, uint2* __restrict__ ok, uint2* __restrict__ ov,
                         __nv_bfloat16* __restrict__ oeta, int L, int H, int seq_text, int CS, float eta_scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane & 15, sub = lane >> 4;
  const int h = blockIdx.y, b = blockIdx.z;
  const int l0 = blockIdx.x * 32 + warp * 4;
  int src[2];
  uint2 qp[2], kp[2], vp[2];
  float2 cs[2][2];  // [row][cos | sin] of this lane's two pairs
#pragma unroll
  for (int it = 0; it < 2; ++it) {  // all loads first
    const int l = min(l0 + 2 * it + sub, L - 1);
    src[it] = index ? index[l] : l;
    const size_t in = (((size_t)b * L + src[it]) * H + h) * 16 + c;
    qp[it] = xq[in]; kp[it] = xk[in]; vp[it] = xv[in];
    if (src[it] >= seq_text) {
      cs[it][0] = cosT[(size_t)(src[it] - seq_text) * 16 + c];
      cs[it][1] = sinT[(size_t)(src[it] - seq_text) * 16 + c];
    } else {
      cs[it][0] = make_float2(1.f, 1.f);  // text token: identity rotation
      cs[it][1] = make_float2(0.f, 0.f);
    }
  }
  const float4 gw = *reinterpret_cast<const float4*>(ln_w + h * 64 + 4 * c);
  const float4 gb = *reinterpret_cast<const float4*>(ln_b + h * 64 + 4 * c);
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int l = l0 + 2 * it + sub;
    float q0 = bf16_lo(qp[it].x), q1 = bf16_hi(qp[it].x), q2 = bf16_lo(qp[it].y), q3 = bf16_hi(qp[it].y);
    float k0 = bf16_lo(kp[it].x), k1 = bf16_hi(kp[it].x), k2 = bf16_lo(kp[it].y), k3 = bf16_hi(kp[it].y);
    float v0 = bf16_lo(vp[it].x), v1 = bf16_hi(vp[it].x), v2 = bf16_lo(vp[it].y), v3 = bf16_hi(vp[it].y);
    const float qn = 1.f / fmaxf(sqrtf(half_warp_sum(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3)), 1e-12f);  // F.normalize eps
    const float kn = 1.f / fmaxf(sqrtf(half_warp_sum(k0 * k0 + k1 * k1 + k2 * k2 + k3 * k3)), 1e-12f);
    q0 *= qn; q1 *= qn; q2 *= qn; q3 *= qn;
    k0 *= kn; k1 *= kn; k2 *= kn; k3 *= kn;
    {  // rotate each interleaved pair by the angle of the token's global video position (identity for text tokens)
      const float2 co = cs[it][0], si = cs[it][1];
      float a = q0 * co.x - q1 * si.x, bb = q0 * si.x + q1 * co.x;
      q0 = a; q1 = bb;
      a = q2 * co.y - q3 * si.y; bb = q2 * si.y + q3 * co.y;
      q2 = a; q3 = bb;
      a = k0 * co.x - k1 * si.x; bb = k0 * si.x + k1 * co.x;
      k0 = a; k1 = bb;
      a = k2 * co.y - k3 * si.y; bb = k2 * si.y + k3 * co.y;
      k2 = a; k3 = bb;
    }
    // reconstruction target: LayerNorm with the unbiased std of (XV - XK), eps added to the std
    float d0 = v0 - k0, d1 = v1 - k1, d2 = v2 - k2, d3 = v3 - k3;
    const float mean = half_warp_sum((d0 + d1) + (d2 + d3)) * (1.f / 64.f);
    d0 -= mean; d1 -= mean; d2 -= mean; d3 -= mean;
    const float inv = 1.f / (sqrtf(half_warp_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) * (1.f / 63.f)) + 1e-8f);
    v0 = fmaf(gw.x, d0 * inv, gb.x) + k0;
    v1 = fmaf(gw.y, d1 * inv, gb.y) + k1;
    v2 = fmaf(gw.z, d2 * inv, gb.z) + k2;
    v3 = fmaf(gw.w, d3 * inv, gb.w) + k3;
    if (l < L) {
      const size_t out = (((size_t)b * H + h) * L + l) * 16 + c;
      oq[out] = make_uint2(pack_bf16(q0, q1), pack_bf16(q2, q3));
      ok[out] = make_uint2(pack_bf16(k0, k1), pack_bf16(k2, k3));
      ov[out] = make_uint2(pack_bf16(v0, v1), pack_bf16(v2, v3));
      if (c == 0) {  // eta of (mini-batch n, column j): lr of token j of the source mini-batch of this mini-batch's last row
        const int n = l / CS, j = l - n * CS;
        const int last = n * CS + CS - 1;
        const int src_mb = (index ? index[last] : last) / CS;
        const float z = lr_logit[((size_t)b * L + (size_t)src_mb * CS + j) * H + h];
        oeta[((size_t)b * H + h) * L + l] = __float2bfloat16(eta_scale / (1.f + __expf(-z)));
      }
    }
  }
}

// Backward of the kernel above (autograd through ttt_layer.py:252-306 in the reference).  Same thread mapping; the
// forward quantities are recomputed from the saved Linear outputs.  Inputs: gQ/gK/gV bf16 [B,H,L,64] (gradients of the
// op inputs), g_eta f32 [B,H,L].  Outputs: gxq/gxk/gxv bf16 [B,L,H*64] at the SOURCE token of every destination token,
// g_logit f32 [B,L,H] (atomic: several interleaved mini-batches may read the lr row of one source mini-batch; pre-zeroed),
// g_ln_w / g_ln_b f32 [H,64] (atomic, pre-zeroed; reduced over the CTA's 32 tokens in shared memory first).
__global__ void __launch_bounds__(256)
ttt_process_input_bwd_kernel(const uint2* __restrict__ xq, const uint2* __restrict__ xk, const uint2* __restrict__ xv,
                             const float* __restrict__ lr_logit, const float2* __restrict__ cosT,
                             const float2* __restrict__ sinT, const float* __restrict__ ln_w, const int* __restrict__ index,
                             const uint2* __restrict__ gQ, const uint2* __restrict__ gK, const uint2* __restrict__ gV,
                             const float* __restrict__ g_eta, uint2* __restrict__ gxq, uint2* __restrict__ gxk,
                             uint2* __restrict__ gxv, float* __restrict__ g_logit, float* __restrict__ g_ln_w,
                             float* __restrict__ g_ln_b, int L, int H, int seq_text, int CS, float eta_scale) {
  __shared__ float red[2][64];  // d gamma, d beta partial sums of this CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane & 15, sub = lane >> 4;
  const int h = blockIdx.y, b = blockIdx.z;
  const int l0 = blockIdx.x * 32 + warp * 4;
  if (threadIdx.x < 128) red[threadIdx.x >> 6][threadIdx.x & 63] = 0.f;
  __syncthreads();
  const float4 gw = *reinterpret_cast<const float4*>(ln_w + h * 64 + 4 * c);
  float dgam[4] = {0.f, 0.f, 0.f, 0.f}, dbet[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int l = l0 + 2 * it + sub;
    const bool ok_row = l < L;
    const int lc = ok_row ? l : L - 1;
    const int src = index ? index[lc] : lc;
    const size_t in = (((size_t)b * L + src) * H + h) * 16 + c;
    const size_t go = (((size_t)b * H + h) * L + lc) * 16 + c;
    const uint2 qp = xq[in], kp = xk[in], vp = xv[in], gqp = gQ[go], gkp = gK[go], gvp = gV[go];
    float2 co = make_float2(1.f, 1.f), si = make_float2(0.f, 0.f);
    if (src >= seq_text) {
      co = cosT[(size_t)(src - seq_text) * 16 + c];
      si = sinT[(size_t)(src - seq_text) * 16 + c];
    }
    float q[4] = {bf16_lo(qp.x), bf16_hi(qp.x), bf16_lo(qp.y), bf16_hi(qp.y)};
    float k[4] = {bf16_lo(kp.x), bf16_hi(kp.x), bf16_lo(kp.y), bf16_hi(kp.y)};
    const float v[4] = {bf16_lo(vp.x), bf16_hi(vp.x), bf16_lo(vp.y), bf16_hi(vp.y)};
    float gq[4] = {bf16_lo(gqp.x), bf16_hi(gqp.x), bf16_lo(gqp.y), bf16_hi(gqp.y)};
    float gk[4] = {bf16_lo(gkp.x), bf16_hi(gkp.x), bf16_lo(gkp.y), bf16_hi(gkp.y)};
    const float gv[4] = {bf16_lo(gvp.x), bf16_hi(gvp.x), bf16_lo(gvp.y), bf16_hi(gvp.y)};
    // recompute: normalised q, k (before the rotation) and the rotated k
    const float qn = 1.f / fmaxf(sqrtf(half_warp_sum(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])), 1e-12f);
    const float kn = 1.f / fmaxf(sqrtf(half_warp_sum(k[0] * k[0] + k[1] * k[1] + k[2] * k[2] + k[3] * k[3])), 1e-12f);
#pragma unroll
    for (int e = 0; e < 4; ++e) { q[e] *= qn; k[e] *= kn; }
    const float kr[4] = {k[0] * co.x - k[1] * si.x, k[0] * si.x + k[1] * co.x, k[2] * co.y - k[3] * si.y, k[2] * si.y + k[3] * co.y};
    // reconstruction target: dn = cen / (s + eps), s = unbiased std of d = v - kr
    float cen[4] = {v[0] - kr[0], v[1] - kr[1], v[2] - kr[2], v[3] - kr[3]};
    const float mean = half_warp_sum((cen[0] + cen[1]) + (cen[2] + cen[3])) * (1.f / 64.f);
#pragma unroll
    for (int e = 0; e < 4; ++e) cen[e] -= mean;
    const float sdev = sqrtf(half_warp_sum(cen[0] * cen[0] + cen[1] * cen[1] + cen[2] * cen[2] + cen[3] * cen[3]) * (1.f / 63.f));
    const float inv = 1.f / (sdev + 1e-8f);
    const float gwv[4] = {gw.x, gw.y, gw.z, gw.w};
    float gdn[4], dot = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      gdn[e] = gv[e] * gwv[e];
      dot = fmaf(gdn[e], cen[e], dot);
      if (ok_row) { dgam[e] = fmaf(gv[e], cen[e] * inv, dgam[e]); dbet[e] += gv[e]; }
    }
    dot = half_warp_sum(dot);
    const float coef = dot * inv * inv / (63.f * fmaxf(sdev, 1e-30f));
    float gc[4], gcs = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) { gc[e] = fmaf(gdn[e], inv, -coef * cen[e]); gcs += gc[e]; }
    const float gcm = half_warp_sum(gcs) * (1.f / 64.f);
    float gd[4], gkr[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      gd[e] = gc[e] - gcm;                 // = d XV (pre-LN value input)
      gkr[e] = gk[e] + gv[e] - gd[e];      // rotated-k gradient: direct + residual path of the target - through d
    }
    // inverse rotation, then backward of x / ||x||
    const float gqh[4] = {gq[0] * co.x + gq[1] * si.x, -gq[0] * si.x + gq[1] * co.x, gq[2] * co.y + gq[3] * si.y, -gq[2] * si.y + gq[3] * co.y};
    const float gkh[4] = {gkr[0] * co.x + gkr[1] * si.x, -gkr[0] * si.x + gkr[1] * co.x, gkr[2] * co.y + gkr[3] * si.y, -gkr[2] * si.y + gkr[3] * co.y};
    const float qd = half_warp_sum(q[0] * gqh[0] + q[1] * gqh[1] + q[2] * gqh[2] + q[3] * gqh[3]);
    const float kd = half_warp_sum(k[0] * gkh[0] + k[1] * gkh[1] + k[2] * gkh[2] + k[3] * gkh[3]);
    if (ok_row) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = (gqh[e] - q[e] * qd) * qn;
      gxq[in] = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = (gkh[e] - k[e] * kd) * kn;
      gxk[in] = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
      gxv[in] = make_uint2(pack_bf16(gd[0], gd[1]), pack_bf16(gd[2], gd[3]));
      if (c == 0) {
        const int n = l / CS, j = l - n * CS;
        const int last = n * CS + CS - 1;
        const int src_mb = (index ? index[last] : last) / CS;
        const size_t zi = ((size_t)b * L + (size_t)src_mb * CS + j) * H + h;
        const float sg = 1.f / (1.f + __expf(-lr_logit[zi]));
        atomicAdd(g_logit + zi, g_eta[((size_t)b * H + h) * L + l] * eta_scale * sg * (1.f - sg));
      }
    }
  }
  // d gamma / d beta: fold the two rows of a warp, then the CTA, then one atomic per feature
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    dgam[e] += __shfl_xor_sync(0xffffffffu, dgam[e], 16);
    dbet[e] += __shfl_xor_sync(0xffffffffu, dbet[e], 16);
  }
  if (sub == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      atomicAdd(&red[0][4 * c + e], dgam[e]);
      atomicAdd(&red[1][4 * c + e], dbet[e]);
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) atomicAdd(g_ln_w + h * 64 + threadIdx.x, red[0][threadIdx.x]);
  else if (threadIdx.x < 128) atomicAdd(g_ln_b + h * 64 + threadIdx.x - 64, red[1][threadIdx.x - 64]);
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
  dim3 grid((L + 31) / 32, H, B);
  ttt_process_input_kernel<<<grid, 256, 0, stream>>>(
      reinterpret_cast<const uint2*>(xq), reinterpret_cast<const uint2*>(xk), reinterpret_cast<const uint2*>(xv), lr_logit,
      reinterpret_cast<const float2*>(cosT), reinterpret_cast<const float2*>(sinT), ln_w, ln_b, index,
      reinterpret_cast<uint2*>(XQ), reinterpret_cast<uint2*>(XK), reinterpret_cast<uint2*>(XV),
      reinterpret_cast<__nv_bfloat16*>(last_eta), L, H, seq_text_length, mini_batch,
      base_lr / 64.f / (float)mini_batch);
  return cudaGetLastError();
}

cudaError_t launch_process_input_backward(const void* xq, const void* xk, const void* xv, const float* lr_logit,
                                          const float* cosT, const float* sinT, const float* ln_w, const int* index,
                                          const void* gQ, const void* gK, const void* gV, const float* g_eta, void* gxq,
                                          void* gxk, void* gxv, float* g_logit, float* g_ln_w, float* g_ln_b, int B, int L,
                                          int H, int seq_text_length, int mini_batch, float base_lr, cudaStream_t stream) {
  if (B <= 0 || L <= 0 || H <= 0 || mini_batch <= 0 || L % mini_batch != 0 || seq_text_length < 0 || seq_text_length > L) {
    g_where = "bad sizes (L must be a multiple of the mini-batch size)";
    return cudaErrorInvalidValue;
  }
  TB_TRY(cudaMemsetAsync(g_logit, 0, (size_t)B * L * H * sizeof(float), stream), "memset g_logit");
  TB_TRY(cudaMemsetAsync(g_ln_w, 0, (size_t)H * 64 * sizeof(float), stream), "memset g_ln_w");
  TB_TRY(cudaMemsetAsync(g_ln_b, 0, (size_t)H * 64 * sizeof(float), stream), "memset g_ln_b");
  g_where = "process_input backward launch";
  dim3 grid((L + 31) / 32, H, B);
  ttt_process_input_bwd_kernel<<<grid, 256, 0, stream>>>(
      reinterpret_cast<const uint2*>(xq), reinterpret_cast<const uint2*>(xk), reinterpret_cast<const uint2*>(xv), lr_logit,
      reinterpret_cast<const float2*>(cosT), reinterpret_cast<const float2*>(sinT), ln_w, index,
      reinterpret_cast<const uint2*>(gQ), reinterpret_cast<const uint2*>(gK), reinterpret_cast<const uint2*>(gV), g_eta,
      reinterpret_cast<uint2*>(gxq), reinterpret_cast<uint2*>(gxk), reinterpret_cast<uint2*>(gxv), g_logit, g_ln_w, g_ln_b, L, H,
      seq_text_length, mini_batch, base_lr / 64.f / (float)mini_batch);
  return cudaGetLastError();
}

}  // namespace tb
