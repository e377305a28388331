/* libttt_b200.so -- C ABI of the B200-native TTT hot path (sm_100a).
 *
 * Drop-in boundary: these entry points are what the reference's native extension `test_time_training`
 * (ttt-tk/test_time_training.cpp:25-105, pybind11) binds for this path, restated with plain device pointers and
 * sizes instead of torch::Tensor.  All pointers are DEVICE pointers unless a name ends in `_host`.  Every function
 * is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream), performs no device
 * synchronisation (the reference blocks with cudaDeviceSynchronize, ttt-tk/kernels/ttt/ttt.cu:714-715), keeps no
 * state between calls, and returns 0 on success or a cudaError_t / negative argument-error code; the message is
 * available from ttt_b200_last_error().  Caller allocates everything (reference: mlp_tk.py:92-98,192-225).
 *
 * Tensor layouts (contiguous, row-major):
 *   XQ, XK, XV, Out        bf16 [B, H, NC, CS, 64]        (ttt.cu:621-630 TORCH_CHECKs)
 *   last_eta               bf16 [B, H, NC, CS]            (= eta[:, :, :, -1, :], mlp_tk.py:105)
 *   ln_weight, ln_bias     f32  [H, 64]                   (ttt_norm_weight/bias as [1,H,1,64], mlp_tk.py:112-113)
 *   TTT-MLP   W1 f32 [B,H,64,256]  b1 f32 [B,H,256]  W2 f32 [B,H,256,64]  b2 f32 [B,H,64]   (CS = 64)
 *   checkpoints            f32  [B, H, K, ...] with K = ceil(NC / checkpoint_group_size): state ENTERING group k
 */
#ifndef TTT_B200_H
#define TTT_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int ttt_b200_version(void);
const char* ttt_b200_last_error(void);

/* TTT-MLP forward scan.  Replaces test_time_training.ttt_forward (ttt-tk/test_time_training.cpp:25-42,
 * ttt-tk/kernels/ttt/ttt.cu:594-721).  Argument order follows the reference (Q, K, V).
 * W*_ckpt may be NULL (no checkpoints written); W*_last may be NULL (final state not exported; the reference never
 * exports it -- it is the hand-off message of the sequence-sharded mode). */
int ttt_b200_mlp_forward(const void* XQ, const void* XK, const void* XV, const void* last_eta,
                         const float* ln_weight, const float* ln_bias,
                         const float* W1, const float* b1, const float* W2, const float* b2,
                         float* W1_ckpt, float* b1_ckpt, float* W2_ckpt, float* b2_ckpt,
                         float* W1_last, float* b1_last, float* W2_last, float* b2_last,
                         void* Out, int B, int H, int NC, int checkpoint_group_size, void* stream);

/* TTT-MLP backward.  Replaces test_time_training.ttt_backward (ttt-tk/test_time_training.cpp:47-91,
 * ttt-tk/kernels/ttt_backward/ttt.cu:1451-1917).  Inputs: the forward's inputs, its fp32 checkpoints and the upstream
 * gradient dOut (bf16, layout of Out).  The upstream gradient of the final state is zero at this boundary
 * (mlp_tk.py:179-182).  Outputs (caller-allocated; need not be pre-zeroed):
 *   d_ln_weight, d_ln_bias  f32 [B,H,64]   (the caller sums over B, mlp_tk.py:277-278)
 *   dW1 f32 [B,H,64,256]  db1 f32 [B,H,256]  dW2 f32 [B,H,256,64]  db2 f32 [B,H,64]   (w.r.t. the initial state)
 *   d_last_eta bf16 [B,H,NC,CS]   dXQ, dXK, dXV bf16 [B,H,NC,CS,64]
 * workspace: device scratch of at least ttt_b200_mlp_backward_workspace_bytes(B,H,G) bytes; it replaces the 16
 * re-materialisation buffers the reference's caller allocates (mlp_tk.py:192-210). */
size_t ttt_b200_mlp_backward_workspace_bytes(int B, int H, int checkpoint_group_size);
int ttt_b200_mlp_backward(const void* XQ, const void* XK, const void* XV, const void* last_eta,
                          const float* ln_weight, const float* ln_bias,
                          const float* W1_ckpt, const float* b1_ckpt, const float* W2_ckpt, const float* b2_ckpt,
                          const void* dOut,
                          float* d_ln_weight, float* d_ln_bias, float* dW1, float* db1, float* dW2, float* db2,
                          void* d_last_eta, void* dXQ, void* dXK, void* dXV,
                          void* workspace, size_t workspace_bytes,
                          int B, int H, int NC, int checkpoint_group_size, void* stream);

/* Same backward with a non-zero upstream gradient of the FINAL state: dW1_last f32 [B,H,64,256], db1_last f32 [B,H,256],
 * dW2_last f32 [B,H,256,64], db2_last f32 [B,H,64] (all four set, or all four NULL = ttt_b200_mlp_backward).  The
 * reference has no such input (mlp_tk.py:179-182 passes zeros): it is the hand-off message of the sequence-sharded mode,
 * where the shard that owns the NEXT mini-batch range sends its dW1/db1/dW2/db2 (gradient w.r.t. its initial state) to
 * the shard that owns this range -- the mirror image of W*_last in ttt_b200_mlp_forward. */
int ttt_b200_mlp_backward_seeded(const void* XQ, const void* XK, const void* XV, const void* last_eta,
                                 const float* ln_weight, const float* ln_bias,
                                 const float* W1_ckpt, const float* b1_ckpt, const float* W2_ckpt, const float* b2_ckpt,
                                 const void* dOut,
                                 const float* dW1_last, const float* db1_last, const float* dW2_last, const float* db2_last,
                                 float* d_ln_weight, float* d_ln_bias, float* dW1, float* db1, float* dW2, float* db2,
                                 void* d_last_eta, void* dXQ, void* dXK, void* dXV,
                                 void* workspace, size_t workspace_bytes,
                                 int B, int H, int NC, int checkpoint_group_size, void* stream);

/* TTT-Linear forward scan (CS = 16, head_dim 64).  Replaces the Triton launch in ttt/models/ssm/linear_triton.py:96-131
 * (kernel ttt/models/ssm/kernels/linear_forward.py:5-148).  XQ/XK/XV/Out bf16 [B,H,NC,16,64]; last_eta bf16 [B,H,NC,16];
 * W1 f32 [B,H,64,64], b1 f32 [B,H,64]; checkpoints [B,H,K,64,64] / [B,H,K,64] (may be NULL); W1_last/b1_last as the
 * reference's W1_last/b1_last outputs (may be NULL). */
int ttt_b200_linear_forward(const void* XQ, const void* XK, const void* XV, const void* last_eta,
                            const float* ln_weight, const float* ln_bias, const float* W1, const float* b1,
                            float* W1_ckpt, float* b1_ckpt, float* W1_last, float* b1_last, void* Out,
                            int B, int H, int NC, int checkpoint_group_size, void* stream);

/* TTT-Linear backward scan.  Replaces the Triton launch in ttt/models/ssm/linear_triton.py:203-246 (kernel
 * ttt/models/ssm/kernels/linear_backward.py:207-520) and its eight per-group spill buffers (linear_triton.py:168-182).
 * W1_ckpt/b1_ckpt are the checkpoints the forward wrote ([B,H,K,64,64] / [B,H,K,64], K = ceil(NC/group)).  Outputs:
 * d_ln_weight/d_ln_bias f32 [B,H,64] per-batch partials (overwritten; the reference sums them over B after the kernel,
 * linear_triton.py:251-252), dW1 f32 [B,H,64,64], db1 f32 [B,H,64] (w.r.t. the initial state), d_last_eta f32
 * [B,H,NC,16] (gradient of the last eta row, the only row the scan reads), dXQ/dXK/dXV bf16 [B,H,NC,16,64].
 * workspace: ttt_b200_linear_backward_workspace_bytes(B,H,NC,group) bytes of device memory (state trajectory images). */
size_t ttt_b200_linear_backward_workspace_bytes(int B, int H, int NC, int checkpoint_group_size);
int ttt_b200_linear_backward(const void* XQ, const void* XK, const void* XV, const void* last_eta,
                             const float* ln_weight, const float* ln_bias, const float* W1_ckpt, const float* b1_ckpt,
                             const void* dOut, float* d_ln_weight, float* d_ln_bias, float* dW1, float* db1,
                             float* d_last_eta, void* dXQ, void* dXK, void* dXV, void* workspace,
                             size_t workspace_bytes, int B, int H, int NC, int checkpoint_group_size, void* stream);

/* Non-causal self-attention forward over one segment, head_dim 64.  Replaces F.scaled_dot_product_attention(q, k, v,
 * attn_mask=None, dropout_p=0, is_causal=False) at ttt/models/cogvideo/dit.py:196-198.  q/k/v/out: bf16 [B, T, H, 64]
 * contiguous = the "b t (h d)" output of the q/k/v Linears, so the reference's rearranges around SDPA disappear.
 * scale = 1/sqrt(64) for the reference's call. */
int ttt_b200_attention_forward(const void* q, const void* k, const void* v, void* out, int B, int T, int H,
                               float scale, void* stream);

/* Training pair of the same attention.  _forward_lse additionally writes lse2 f32 [B,H,T] = log2-domain log-sum-exp of
 * every query row (what the library's flash-attention forward saves for its backward).  _backward computes the gradient
 * that autograd takes through F.scaled_dot_product_attention at dit.py:196-198: dq/dk/dv bf16 [B,T,H,64] from
 * q, k, v, out, dout (same layout) and lse2; delta_scratch: f32 [B,H,T] device scratch (rowsum(dout * out)). */
int ttt_b200_attention_forward_lse(const void* q, const void* k, const void* v, void* out, float* lse2, int B, int T, int H,
                                   float scale, void* stream);
int ttt_b200_attention_backward(const void* q, const void* k, const void* v, const void* out, const void* dout,
                                const float* lse2, float* delta_scratch, void* dq, void* dk, void* dv, int B, int T, int H,
                                float scale, void* stream);

/* Input preparation of the TTT op after the q/k/v Linears.  Replaces the torch-op chain of
 * TTTBase.process_input, ttt/models/ssm/ttt_layer.py:252-306: L2-normalise q,k (:265-266), RoPE with global positions on
 * the video tokens (:271-276), reconstruction target XV <- LN(XV - XK)*gamma + beta + XK (:219-235), transpose to
 * mini-batches (:237-250), multi-scene interleave (:157-189) and eta (:143-155,287-288).
 * xq/xk/xv: bf16 [B,L,H*64] (outputs of wq/wk/wv); lr_logit: f32 [B,L,H] = X.w_h + b_h; rope_cos/rope_sin: f32 [Lv,32]
 * (angle per video position and feature pair, ssm/utils.py:9-53); ln_weight/ln_bias: f32 [H,64]; interleave_index: int32 [L]
 * source token of every destination token, or NULL for a single scene.  Outputs: XQ/XK/XV bf16 [B,H,L/CS,CS,64],
 * last_eta bf16 [B,H,L/CS,CS] = the one eta row the scan reads (the reference materialises [B,H,NC,CS,CS]). */
int ttt_b200_process_input(const void* xq, const void* xk, const void* xv, const float* lr_logit, const float* rope_cos,
                           const float* rope_sin, const float* ln_weight, const float* ln_bias, const int* interleave_index,
                           void* XQ, void* XK, void* XV, void* last_eta, int B, int L, int H, int seq_text_length,
                           int mini_batch_size, float ttt_base_lr, void* stream);

/* Backward of ttt_b200_process_input (what autograd computes through ttt_layer.py:252-306).  dXQ/dXK/dXV bf16
 * [B,H,L/CS,CS,64], d_last_eta f32 [B,H,L/CS,CS]; outputs dxq/dxk/dxv bf16 [B,L,H*64], d_lr_logit f32 [B,L,H],
 * d_ln_weight / d_ln_bias f32 [H,64] (all overwritten). */
int ttt_b200_process_input_backward(const void* xq, const void* xk, const void* xv, const float* lr_logit,
                                    const float* rope_cos, const float* rope_sin, const float* ln_weight,
                                    const int* interleave_index, const void* dXQ, const void* dXK, const void* dXV,
                                    const float* d_last_eta, void* dxq, void* dxk, void* dxv, float* d_lr_logit,
                                    float* d_ln_weight, float* d_ln_bias, int B, int L, int H, int seq_text_length,
                                    int mini_batch_size, float ttt_base_lr, void* stream);

/* Output side of the TTT layer before wo: op_out bf16 [B,H,L/CS,CS,64] -> transpose to [B,L,H*64]
 * (ttt_layer.py:456,472), post_norm LayerNorm(H*64, eps) (ttt_layer.py:71,324) and undo_interleave (ttt_layer.py:191-217,
 * as a gather index int32 [L], NULL for a single scene; applied before the per-token wo Linear, with which it commutes).
 * out: bf16 [B,L,H*64].  _backward: d_out bf16 [B,L,H*64] -> d_op_out bf16 [B,H,L/CS,CS,64], d_post_norm_weight /
 * d_post_norm_bias f32 [H*64] (overwritten). */
int ttt_b200_output_norm(const void* op_out, const float* post_norm_weight, const float* post_norm_bias,
                         const int* undo_interleave_index, void* out, int B, int L, int H, float eps, void* stream);
int ttt_b200_output_norm_backward(const void* op_out, const float* post_norm_weight, const int* undo_interleave_index,
                                  const void* d_out, void* d_op_out, float* d_post_norm_weight, float* d_post_norm_bias, int B,
                                  int L, int H, float eps, void* stream);

/* Learned residual gate (+ optional sequence reversal) of the bidirectional TTT pass.
 * Replaces SeqModelingBlock._gate / SSMGating / _reverse_text_chunks / torch.flip in
 * ttt/models/cogvideo/dit.py:90-103,213-222,241-266.  Tensors are bf16 [B, L, E], text tokens first
 * (text_len = seq_text_length, split in num_chunks equal chunks), alpha f32 [E].
 *   out[b,l,:] = res[b,l,:] + tanh(alpha(l)) * s[b, perm_s ? perm(l) : l, :]      alpha(l) = l < text_len ? text : video
 *   if rev != NULL:  rev[b, perm(l), :] = out[b,l,:]                             (input of the reversed TTT pass)
 * perm = text chunks in reverse order + video tokens flipped (an involution).  Backward: dres, ds (same layout as s),
 * d_alpha_text / d_alpha_video f32 [E] (overwritten). */
int ttt_b200_gate_forward(const void* res, const void* s, const float* alpha_text, const float* alpha_video,
                          void* out, void* rev, int B, int L, int E, int text_len, int num_chunks, int perm_s,
                          void* stream);
int ttt_b200_gate_backward(const void* dout, const void* drev, const void* s, const float* alpha_text,
                           const float* alpha_video, void* dres, void* ds, float* d_alpha_text, float* d_alpha_video,
                           int B, int L, int E, int text_len, int num_chunks, int perm_s, void* stream);

/* Prologue of the local attention: per-head LayerNorm of q and k + 3-D rotary embedding of the video tokens with segment-local
 * positions, one pass.  Replaces self.q_norm / self.k_norm / self.rotary at ttt/models/cogvideo/dit.py:188-194 (nn.LayerNorm(64,
 * eps), cogvideo/utils.py:93-99,432-437).  q/k/q_out/k_out bf16 [B,T,H,64] (layout of the q / k Linear outputs), tokens
 * t < text_len are text (not rotated), token t >= text_len uses table row t - text_len; norm_weight / norm_bias f32 [2,64]
 * (q_norm then k_norm); rope_cos / rope_sin f32 [>= T - text_len, 64] with every angle repeated over its feature pair.
 * _backward: dq_out/dk_out -> dq/dk (bf16), d_norm_weight / d_norm_bias f32 [2,64] (overwritten). */
int ttt_b200_qk_norm_rope(const void* q, const void* k, const float* norm_weight, const float* norm_bias, const float* rope_cos,
                          const float* rope_sin, void* q_out, void* k_out, int B, int T, int H, int text_len, float eps,
                          void* stream);
int ttt_b200_qk_norm_rope_backward(const void* q, const void* k, const float* norm_weight, const float* rope_cos,
                                   const float* rope_sin, const void* dq_out, const void* dk_out, void* dq, void* dk,
                                   float* d_norm_weight, float* d_norm_bias, int B, int T, int H, int text_len, float eps,
                                   void* stream);

/* adaLN shell of the DiT TransformerLayer around the hot path (ttt/models/cogvideo/dit.py:321-382).
 * ln_affine: out[b,l,:] = LayerNorm_noaffine(x[b,l,:]; eps) * A[b,s,:] + C[b,s,:], s = (l < text_len ? 0 : 1) -- replaces
 * pre_seq_layernorm / pre_mlp_layernorm + modulate(x, shift, scale) (dit.py:344-345,367-368) with the caller folding
 * A = gamma * (1 + scale), C = beta * (1 + scale) + shift (f32 [B,2,E], text row first).  x/out bf16 [B,L,E], text tokens
 * first, E % 64 == 0, E <= 4096.  _backward: d_out bf16 -> d_x bf16, d_A / d_C f32 [B,2,E] (overwritten).
 * gate_add: out = x + G[b,s,:] * y -- the gated residuals emb + gate * block_out (dit.py:349-350,381-382), G f32 [B,2,E];
 * _backward: d_y = G * d_out (bf16), d_G f32 [B,2,E] = sum over the rows of g * y (overwritten); d_x = d_out. */
int ttt_b200_ln_affine(const void* x, const float* A, const float* C, void* out, int B, int L, int E, int text_len, float eps,
                       void* stream);
int ttt_b200_ln_affine_backward(const void* x, const float* A, const void* d_out, void* d_x, float* d_A, float* d_C, int B, int L,
                                int E, int text_len, float eps, void* stream);
int ttt_b200_gate_add(const void* x, const void* y, const float* G, void* out, int B, int L, int E, int text_len, void* stream);
int ttt_b200_gate_add_backward(const void* d_out, const void* y, const float* G, void* d_y, float* d_G, int B, int L, int E,
                               int text_len, void* stream);

#ifdef __cplusplus
}
#endif
#endif
