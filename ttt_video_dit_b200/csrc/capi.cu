// extern "C" surface of libttt_b200.so (declared in include/ttt_b200.h).
#include "../../include/ttt_b200.h"
#include "capi_util.h"

namespace tb { thread_local const char* g_where = ""; unsigned* g_timing_buf = nullptr; }

extern "C" {

int ttt_b200_version(void) { return 100; }
const char* ttt_b200_last_error(void) { return g_err; }

int ttt_b200_mlp_forward(const void* XQ, const void* XK, const void* XV, const void* last_eta, const float* ln_weight,
                         const float* ln_bias, const float* W1, const float* b1, const float* W2, const float* b2,
                         float* W1_ckpt, float* b1_ckpt, float* W2_ckpt, float* b2_ckpt, float* W1_last,
                         float* b1_last, float* W2_last, float* b2_last, void* Out, int B, int H, int NC,
                         int checkpoint_group_size, void* stream) {
  if (!XQ || !XK || !XV || !last_eta || !ln_weight || !ln_bias || !W1 || !b1 || !W2 || !b2 || !Out)
    return fail(-1, "ttt_b200_mlp_forward: null pointer argument");
  if (B <= 0 || H <= 0 || NC <= 0 || checkpoint_group_size <= 0)
    return fail(-2, "ttt_b200_mlp_forward: B, H, NC and checkpoint_group_size must be positive");
  const bool any_ck = W1_ckpt || b1_ckpt || W2_ckpt || b2_ckpt, all_ck = W1_ckpt && b1_ckpt && W2_ckpt && b2_ckpt;
  if (any_ck && !all_ck) return fail(-3, "ttt_b200_mlp_forward: checkpoint buffers must be all set or all NULL");
  const bool any_l = W1_last || b1_last || W2_last || b2_last, all_l = W1_last && b1_last && W2_last && b2_last;
  if (any_l && !all_l) return fail(-3, "ttt_b200_mlp_forward: final-state buffers must be all set or all NULL");
  TB_BIND_DEVICE(XQ);
  return cuda_ret(tb::launch_mlp_forward(XQ, XK, XV, last_eta, ln_weight, ln_bias, W1, b1, W2, b2, W1_ckpt, b1_ckpt,
                                         W2_ckpt, b2_ckpt, W1_last, b1_last, W2_last, b2_last, Out, B, H, NC,
                                         checkpoint_group_size, (cudaStream_t)stream),
                  "ttt_b200_mlp_forward");
}

size_t ttt_b200_mlp_backward_workspace_bytes(int B, int H, int G) { return tb::mlp_backward_workspace_bytes(B, H, G); }

int ttt_b200_mlp_backward_seeded(const void* XQ, const void* XK, const void* XV, const void* last_eta,
                                 const float* ln_weight, const float* ln_bias, const float* W1_ckpt, const float* b1_ckpt,
                                 const float* W2_ckpt, const float* b2_ckpt, const void* dOut, const float* dW1_last,
                                 const float* db1_last, const float* dW2_last, const float* db2_last, float* d_ln_weight,
                                 float* d_ln_bias, float* dW1, float* db1, float* dW2, float* db2, void* d_last_eta,
                                 void* dXQ, void* dXK, void* dXV, void* workspace, size_t workspace_bytes, int B, int H,
                                 int NC, int checkpoint_group_size, void* stream) {
  if (!XQ || !XK || !XV || !last_eta || !ln_weight || !ln_bias || !W1_ckpt || !b1_ckpt || !W2_ckpt || !b2_ckpt || !dOut ||
      !d_ln_weight || !d_ln_bias || !dW1 || !db1 || !dW2 || !db2 || !d_last_eta || !dXQ || !dXK || !dXV || !workspace)
    return fail(-1, "ttt_b200_mlp_backward: null pointer argument");
  const bool any_u = dW1_last || db1_last || dW2_last || db2_last, all_u = dW1_last && db1_last && dW2_last && db2_last;
  if (any_u && !all_u) return fail(-3, "ttt_b200_mlp_backward_seeded: upstream state gradients must be all set or all NULL");
  if (B <= 0 || H <= 0 || NC <= 0 || checkpoint_group_size <= 0)
    return fail(-2, "ttt_b200_mlp_backward: B, H, NC and checkpoint_group_size must be positive");
  if (workspace_bytes < tb::mlp_backward_workspace_bytes(B, H, checkpoint_group_size))
    return fail(-4, "ttt_b200_mlp_backward: workspace too small (see ttt_b200_mlp_backward_workspace_bytes)");
  TB_BIND_DEVICE(XQ);
  return cuda_ret(tb::launch_mlp_backward(XQ, XK, XV, last_eta, ln_weight, ln_bias, W1_ckpt, b1_ckpt, W2_ckpt, b2_ckpt,
                                          dOut, d_ln_weight, d_ln_bias, dW1, db1, dW2, db2, d_last_eta, dXQ, dXK, dXV,
                                          workspace, workspace_bytes, B, H, NC, checkpoint_group_size,
                                          (cudaStream_t)stream, dW1_last, db1_last, dW2_last, db2_last),
                  "ttt_b200_mlp_backward");
}

int ttt_b200_mlp_backward(const void* XQ, const void* XK, const void* XV, const void* last_eta, const float* ln_weight,
                          const float* ln_bias, const float* W1_ckpt, const float* b1_ckpt, const float* W2_ckpt,
                          const float* b2_ckpt, const void* dOut, float* d_ln_weight, float* d_ln_bias, float* dW1,
                          float* db1, float* dW2, float* db2, void* d_last_eta, void* dXQ, void* dXK, void* dXV,
                          void* workspace, size_t workspace_bytes, int B, int H, int NC, int checkpoint_group_size,
                          void* stream) {
  return ttt_b200_mlp_backward_seeded(XQ, XK, XV, last_eta, ln_weight, ln_bias, W1_ckpt, b1_ckpt, W2_ckpt, b2_ckpt, dOut,
                                      nullptr, nullptr, nullptr, nullptr, d_ln_weight, d_ln_bias, dW1, db1, dW2, db2,
                                      d_last_eta, dXQ, dXK, dXV, workspace, workspace_bytes, B, H, NC,
                                      checkpoint_group_size, stream);
}

int ttt_b200_linear_forward(const void* XQ, const void* XK, const void* XV, const void* last_eta, const float* ln_weight,
                            const float* ln_bias, const float* W1, const float* b1, float* W1_ckpt, float* b1_ckpt,
                            float* W1_last, float* b1_last, void* Out, int B, int H, int NC, int checkpoint_group_size,
                            void* stream) {
  if (!XQ || !XK || !XV || !last_eta || !ln_weight || !ln_bias || !W1 || !b1 || !Out)
    return fail(-1, "ttt_b200_linear_forward: null pointer argument");
  if ((W1_ckpt == nullptr) != (b1_ckpt == nullptr) || (W1_last == nullptr) != (b1_last == nullptr))
    return fail(-3, "ttt_b200_linear_forward: W1/b1 buffer pairs must both be set or both be NULL");
  TB_BIND_DEVICE(XQ);
  return cuda_ret(tb::launch_linear_forward(XQ, XK, XV, last_eta, ln_weight, ln_bias, W1, b1, W1_ckpt, b1_ckpt, W1_last,
                                            b1_last, Out, B, H, NC, checkpoint_group_size, (cudaStream_t)stream),
                  "ttt_b200_linear_forward");
}

size_t ttt_b200_linear_backward_workspace_bytes(int B, int H, int NC, int G) {
  if (B <= 0 || H <= 0 || NC <= 0 || G <= 0) return 0;
  return tb::linear_backward_workspace_bytes(B, H, NC, G > NC ? NC : G);
}

int ttt_b200_linear_backward(const void* XQ, const void* XK, const void* XV, const void* last_eta, const float* ln_weight,
                             const float* ln_bias, const float* W1_ckpt, const float* b1_ckpt, const void* dOut,
                             float* d_ln_weight, float* d_ln_bias, float* dW1, float* db1, float* d_last_eta, void* dXQ,
                             void* dXK, void* dXV, void* workspace, size_t workspace_bytes, int B, int H, int NC,
                             int checkpoint_group_size, void* stream) {
  if (!XQ || !XK || !XV || !last_eta || !ln_weight || !ln_bias || !W1_ckpt || !b1_ckpt || !dOut || !d_ln_weight ||
      !d_ln_bias || !dW1 || !db1 || !d_last_eta || !dXQ || !dXK || !dXV || !workspace)
    return fail(-1, "ttt_b200_linear_backward: null pointer argument");
  if (B <= 0 || H <= 0 || NC <= 0 || checkpoint_group_size <= 0)
    return fail(-2, "ttt_b200_linear_backward: B, H, NC and checkpoint_group_size must be positive");
  if (workspace_bytes < ttt_b200_linear_backward_workspace_bytes(B, H, NC, checkpoint_group_size))
    return fail(-4, "ttt_b200_linear_backward: workspace too small (see ttt_b200_linear_backward_workspace_bytes)");
  TB_BIND_DEVICE(XQ);
  return cuda_ret(tb::launch_linear_backward(XQ, XK, XV, last_eta, ln_weight, ln_bias, W1_ckpt, b1_ckpt, dOut,
                                             d_ln_weight, d_ln_bias, dW1, db1, d_last_eta, dXQ, dXK, dXV, workspace,
                                             workspace_bytes, B, H, NC, checkpoint_group_size, (cudaStream_t)stream),
                  "ttt_b200_linear_backward");
}

int ttt_b200_attention_forward(const void* q, const void* k, const void* v, void* out, int B, int T, int H, float scale,
                               void* stream) {
  if (!q || !k || !v || !out) return fail(-1, "ttt_b200_attention_forward: null pointer argument");
  TB_BIND_DEVICE(q);
  return cuda_ret(tb::launch_attention_forward(q, k, v, out, nullptr, B, T, H, scale, (cudaStream_t)stream),
                  "ttt_b200_attention_forward");
}

int ttt_b200_attention_forward_lse(const void* q, const void* k, const void* v, void* out, float* lse2, int B, int T, int H,
                                   float scale, void* stream) {
  if (!q || !k || !v || !out || !lse2) return fail(-1, "ttt_b200_attention_forward_lse: null pointer argument");
  TB_BIND_DEVICE(q);
  return cuda_ret(tb::launch_attention_forward(q, k, v, out, lse2, B, T, H, scale, (cudaStream_t)stream),
                  "ttt_b200_attention_forward_lse");
}

int ttt_b200_attention_backward(const void* q, const void* k, const void* v, const void* out, const void* dout,
                                const float* lse2, float* delta_scratch, void* dq, void* dk, void* dv, int B, int T, int H,
                                float scale, void* stream) {
  if (!q || !k || !v || !out || !dout || !lse2 || !delta_scratch || !dq || !dk || !dv)
    return fail(-1, "ttt_b200_attention_backward: null pointer argument");
  TB_BIND_DEVICE(q);
  return cuda_ret(tb::launch_attention_backward(q, k, v, out, dout, lse2, delta_scratch, dq, dk, dv, B, T, H, scale,
                                                (cudaStream_t)stream),
                  "ttt_b200_attention_backward");
}

int ttt_b200_process_input(const void* xq, const void* xk, const void* xv, const float* lr_logit, const float* rope_cos,
                           const float* rope_sin, const float* ln_weight, const float* ln_bias, const int* interleave_index,
                           void* XQ, void* XK, void* XV, void* last_eta, int B, int L, int H, int seq_text_length,
                           int mini_batch_size, float ttt_base_lr, void* stream) {
  if (!xq || !xk || !xv || !lr_logit || !rope_cos || !rope_sin || !ln_weight || !ln_bias || !XQ || !XK || !XV || !last_eta)
    return fail(-1, "ttt_b200_process_input: null pointer argument");
  TB_BIND_DEVICE(xq);
  return cuda_ret(tb::launch_process_input(xq, xk, xv, lr_logit, rope_cos, rope_sin, ln_weight, ln_bias, interleave_index, XQ,
                                           XK, XV, last_eta, B, L, H, seq_text_length, mini_batch_size, ttt_base_lr,
                                           (cudaStream_t)stream),
                  "ttt_b200_process_input");
}

int ttt_b200_process_input_backward(const void* xq, const void* xk, const void* xv, const float* lr_logit,
                                    const float* rope_cos, const float* rope_sin, const float* ln_weight,
                                    const int* interleave_index, const void* dXQ, const void* dXK, const void* dXV,
                                    const float* d_last_eta, void* dxq, void* dxk, void* dxv, float* d_lr_logit,
                                    float* d_ln_weight, float* d_ln_bias, int B, int L, int H, int seq_text_length,
                                    int mini_batch_size, float ttt_base_lr, void* stream) {
  if (!xq || !xk || !xv || !lr_logit || !rope_cos || !rope_sin || !ln_weight || !dXQ || !dXK || !dXV || !d_last_eta || !dxq ||
      !dxk || !dxv || !d_lr_logit || !d_ln_weight || !d_ln_bias)
    return fail(-1, "ttt_b200_process_input_backward: null pointer argument");
  TB_BIND_DEVICE(xq);
  return cuda_ret(tb::launch_process_input_backward(xq, xk, xv, lr_logit, rope_cos, rope_sin, ln_weight, interleave_index, dXQ,
                                                    dXK, dXV, d_last_eta, dxq, dxk, dxv, d_lr_logit, d_ln_weight, d_ln_bias, B,
                                                    L, H, seq_text_length, mini_batch_size, ttt_base_lr, (cudaStream_t)stream),
                  "ttt_b200_process_input_backward");
}

int ttt_b200_output_norm(const void* op_out, const float* post_norm_weight, const float* post_norm_bias,
                         const int* undo_interleave_index, void* out, int B, int L, int H, float eps, void* stream) {
  if (!op_out || !post_norm_weight || !post_norm_bias || !out) return fail(-1, "ttt_b200_output_norm: null pointer argument");
  TB_BIND_DEVICE(op_out);
  return cuda_ret(tb::launch_output_norm(op_out, post_norm_weight, post_norm_bias, undo_interleave_index, out, B, L, H, eps,
                                         (cudaStream_t)stream),
                  "ttt_b200_output_norm");
}

int ttt_b200_output_norm_backward(const void* op_out, const float* post_norm_weight, const int* undo_interleave_index,
                                  const void* d_out, void* d_op_out, float* d_post_norm_weight, float* d_post_norm_bias, int B,
                                  int L, int H, float eps, void* stream) {
  if (!op_out || !post_norm_weight || !d_out || !d_op_out || !d_post_norm_weight || !d_post_norm_bias)
    return fail(-1, "ttt_b200_output_norm_backward: null pointer argument");
  TB_BIND_DEVICE(op_out);
  return cuda_ret(tb::launch_output_norm_backward(op_out, post_norm_weight, undo_interleave_index, d_out, d_op_out,
                                                  d_post_norm_weight, d_post_norm_bias, B, L, H, eps, (cudaStream_t)stream),
                  "ttt_b200_output_norm_backward");
}

int ttt_b200_gate_forward(const void* res, const void* s, const float* alpha_text, const float* alpha_video, void* out,
                          void* rev, int B, int L, int E, int text_len, int num_chunks, int perm_s, void* stream) {
  if (!res || !s || !alpha_text || !alpha_video || !out) return fail(-1, "ttt_b200_gate_forward: null pointer argument");
  TB_BIND_DEVICE(res);
  return cuda_ret(tb::launch_gate_forward(res, s, alpha_text, alpha_video, out, rev, B, L, E, text_len, num_chunks,
                                          perm_s, (cudaStream_t)stream),
                  "ttt_b200_gate_forward");
}

int ttt_b200_gate_backward(const void* dout, const void* drev, const void* s, const float* alpha_text,
                           const float* alpha_video, void* dres, void* ds, float* d_alpha_text, float* d_alpha_video,
                           int B, int L, int E, int text_len, int num_chunks, int perm_s, void* stream) {
  if (!dout || !s || !alpha_text || !alpha_video || !dres || !ds || !d_alpha_text || !d_alpha_video)
    return fail(-1, "ttt_b200_gate_backward: null pointer argument");
  TB_BIND_DEVICE(dout);
  return cuda_ret(tb::launch_gate_backward(dout, drev, s, alpha_text, alpha_video, dres, ds, d_alpha_text,
                                           d_alpha_video, B, L, E, text_len, num_chunks, perm_s, (cudaStream_t)stream),
                  "ttt_b200_gate_backward");
}

int ttt_b200_qk_norm_rope(const void* q, const void* k, const float* norm_weight, const float* norm_bias, const float* rope_cos,
                          const float* rope_sin, void* q_out, void* k_out, int B, int T, int H, int text_len, float eps,
                          void* stream) {
  if (!q || !k || !norm_weight || !norm_bias || !rope_cos || !rope_sin || !q_out || !k_out)
    return fail(-1, "ttt_b200_qk_norm_rope: null pointer argument");
  TB_BIND_DEVICE(q);
  return cuda_ret(tb::launch_qk_norm_rope(q, k, norm_weight, norm_bias, rope_cos, rope_sin, q_out, k_out, B, T, H, text_len, eps,
                                          (cudaStream_t)stream),
                  "ttt_b200_qk_norm_rope");
}

int ttt_b200_qk_norm_rope_backward(const void* q, const void* k, const float* norm_weight, const float* rope_cos,
                                   const float* rope_sin, const void* dq_out, const void* dk_out, void* dq, void* dk,
                                   float* d_norm_weight, float* d_norm_bias, int B, int T, int H, int text_len, float eps,
                                   void* stream) {
  if (!q || !k || !norm_weight || !rope_cos || !rope_sin || !dq_out || !dk_out || !dq || !dk || !d_norm_weight || !d_norm_bias)
    return fail(-1, "ttt_b200_qk_norm_rope_backward: null pointer argument");
  TB_BIND_DEVICE(q);
  return cuda_ret(tb::launch_qk_norm_rope_backward(q, k, norm_weight, rope_cos, rope_sin, dq_out, dk_out, dq, dk, d_norm_weight,
                                                   d_norm_bias, B, T, H, text_len, eps, (cudaStream_t)stream),
                  "ttt_b200_qk_norm_rope_backward");
}

int ttt_b200_ln_affine(const void* x, const float* A, const float* C, void* out, int B, int L, int E, int text_len, float eps,
                       void* stream) {
  if (!x || !A || !C || !out) return fail(-1, "ttt_b200_ln_affine: null pointer argument");
  TB_BIND_DEVICE(x);
  return cuda_ret(tb::launch_ln_affine(x, A, C, out, B, L, E, text_len, eps, (cudaStream_t)stream), "ttt_b200_ln_affine");
}

int ttt_b200_ln_affine_backward(const void* x, const float* A, const void* d_out, void* d_x, float* d_A, float* d_C, int B, int L,
                                int E, int text_len, float eps, void* stream) {
  if (!x || !A || !d_out || !d_x || !d_A || !d_C) return fail(-1, "ttt_b200_ln_affine_backward: null pointer argument");
  TB_BIND_DEVICE(x);
  return cuda_ret(tb::launch_ln_affine_backward(x, A, d_out, d_x, d_A, d_C, B, L, E, text_len, eps, (cudaStream_t)stream),
                  "ttt_b200_ln_affine_backward");
}

int ttt_b200_gate_add(const void* x, const void* y, const float* G, void* out, int B, int L, int E, int text_len, void* stream) {
  if (!x || !y || !G || !out) return fail(-1, "ttt_b200_gate_add: null pointer argument");
  TB_BIND_DEVICE(x);
  return cuda_ret(tb::launch_gate_add(x, y, G, out, B, L, E, text_len, (cudaStream_t)stream), "ttt_b200_gate_add");
}

int ttt_b200_gate_add_backward(const void* d_out, const void* y, const float* G, void* d_y, float* d_G, int B, int L, int E,
                               int text_len, void* stream) {
  if (!d_out || !y || !G || !d_y || !d_G) return fail(-1, "ttt_b200_gate_add_backward: null pointer argument");
  TB_BIND_DEVICE(d_out);
  return cuda_ret(tb::launch_gate_add_backward(d_out, y, G, d_y, d_G, B, L, E, text_len, (cudaStream_t)stream),
                  "ttt_b200_gate_add_backward");
}

#ifdef TTT_PHASE_TIMING
/* phase-timing build only (lib/libttt_b200_dbg.so): device buffer (>= 512 bytes) receiving per-phase cycle counts */
int ttt_b200_debug_set_timing_buffer(void* dev_buf_512_bytes) {
  tb::g_timing_buf = reinterpret_cast<unsigned*>(dev_buf_512_bytes);
  return 1;
}
#endif

}  // extern "C"
