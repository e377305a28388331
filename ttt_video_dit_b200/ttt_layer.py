"""Mirror of the reference TTT layer forward on libttt_b200.so: ``TTTMLP.forward`` / ``TTTLinear.forward``
(ttt/models/ssm/ttt_layer.py:314-334) = process_input (:252-306) -> ``ttt`` (:429-473 / :360-398) -> post_norm -> wo ->
undo_interleave.  Everything except the q/k/v/lr/wo Linears (library GEMMs, as in the reference) runs in this repo's
kernels: csrc/process_input.cu, the TTT-MLP / TTT-Linear scans, csrc/output_norm.cu.  Differentiable end to end for
kind="ttt_mlp" and "ttt_linear" (every custom op has a native backward).  No eager fallback.

``P`` uses the reference module's state_dict names: wq/wk/wv/wo ``.weight``/``.bias``, ``learnable_ttt_lr_weight`` [H,1,E],
``learnable_ttt_lr_bias`` [H,1], ``ttt_norm_weight``/``ttt_norm_bias`` [H,64], ``post_norm.weight``/``.bias`` [E],
``W1`` [H,64,256] / ``b1`` [H,1,256] / ``W2`` [H,256,64] / ``b2`` [H,1,64] (TTT-MLP) or ``W1`` [H,64,64] / ``b1`` [H,1,64]
(TTT-Linear).
"""
import torch
import torch.nn.functional as F

from . import linear_triton, mlp_tk, process_input


def _tile(p, B):  # ttt_layer.py:434-437: the per-head initial state is shared by the batch
    return torch.tile(p.unsqueeze(0), dims=(B,) + (1,) * p.dim()).contiguous()


def ttt_layer_forward(hidden_states, P, rope_cos, rope_sin, seq_text_length, mini_batch_size, ttt_base_lr,
                      scan_checkpoint_group_size, kind="ttt_mlp", interleave_index=None, undo_interleave_index=None,
                      post_norm_eps=1e-6):
    """hidden_states bf16 [B, L, E] -> [B, L, E] (``TTTBase.forward``, ttt_layer.py:314-334)."""
    B, L, E = hidden_states.shape
    if L % mini_batch_size:
        raise RuntimeError("Sequence len must be multiple of mini batch size.")  # ttt_layer.py:320-322
    inp = process_input.process_input(hidden_states, P, rope_cos, rope_sin, seq_text_length, mini_batch_size, ttt_base_lr,
                                      interleave_index)
    NC = L // mini_batch_size
    G = min(max(int(scan_checkpoint_group_size), 1), NC)  # ttt_layer.py:439
    if kind == "ttt_mlp":
        out = mlp_tk.ttt_mlp_op(P["ttt_norm_weight"], P["ttt_norm_bias"], _tile(P["W1"], B), _tile(P["b1"], B), _tile(P["W2"], B),
                                _tile(P["b2"], B), inp["XQ"], inp["XV"], inp["XK"], inp["last_eta"], G)
    elif kind == "ttt_linear":
        out = linear_triton.ttt_linear_op(P["ttt_norm_weight"], P["ttt_norm_bias"], _tile(P["W1"], B), _tile(P["b1"], B),
                                          inp["XQ"], inp["XV"], inp["XK"], inp["last_eta"], G)
    else:
        raise ValueError(kind)
    x = process_input.output_norm(out.contiguous(), P["post_norm.weight"], P["post_norm.bias"], post_norm_eps, undo_interleave_index)
    return F.linear(x, P["wo.weight"], P["wo.bias"])
