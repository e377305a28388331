"""Drop-in for ``ttt.models.ssm.linear_triton.TritonLinear`` (reference: ttt/models/ssm/linear_triton.py:12-362) on
libttt_b200.so -- hand-written sm_100a CUDA, no Triton.

``TritonLinear.apply(ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, XQ, XV, XK, eta, checkpoint_group_size)`` keeps
the reference's signature (Q, V, K order, linear_triton.py:16-27) and the saved checkpoints (W1/b1 every
``checkpoint_group_size`` mini-batches, linear_triton.py:87-88).  Round-1 state: the forward scan is native; the
backward scan kernel (reference: kernels/linear_backward.py) is not built yet, so calling ``.backward`` raises -- there
is deliberately no eager fallback.
"""
import math

import torch

from . import _lib


def linear_forward(XQ, XK, XV, last_eta, ln_w, ln_b, W1, b1, checkpoint_group_size, want_last=False):
    """Native TTT-Linear forward.  XQ/XK/XV bf16 [B,H,NC,16,64]; last_eta bf16 [B,H,NC,16(,1)]; W1 [B,H,64,64]; b1 [B,H,1,64].
    Returns (out bf16, (W1_ckpt, b1_ckpt), (W1_last, b1_last) or None)."""
    B, H, NC, CS, F = XQ.shape
    if CS != 16 or F != 64:
        raise RuntimeError("TTT-Linear kernel is specialised for mini_batch_size 16, head_dim 64")
    for t, n in ((XQ, "XQ"), (XK, "XK"), (XV, "XV")):
        if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous()):
            raise RuntimeError(f"{n} must be a contiguous CUDA bf16 tensor")
    dev = XQ.device
    G = int(checkpoint_group_size)
    K = math.ceil(NC / G)
    le = last_eta.to(torch.bfloat16).reshape(B, H, NC, CS).contiguous()
    lw = ln_w.detach().reshape(H, F).float().contiguous()
    lb = ln_b.detach().reshape(H, F).float().contiguous()
    W1f = W1.detach().float().contiguous()
    b1f = b1.detach().float().reshape(B, H, F).contiguous()
    out = torch.empty_like(XQ)
    W1c = torch.empty(B, H, K, F, F, device=dev, dtype=torch.float32)
    b1c = torch.empty(B, H, K, 1, F, device=dev, dtype=torch.float32)
    W1l = torch.empty(B, H, F, F, device=dev, dtype=torch.float32) if want_last else None
    b1l = torch.empty(B, H, 1, F, device=dev, dtype=torch.float32) if want_last else None
    p = _lib.ptr
    code = _lib.lib().ttt_b200_linear_forward(p(XQ), p(XK), p(XV), p(le), p(lw), p(lb), p(W1f), p(b1f), p(W1c), p(b1c),
                                              p(W1l), p(b1l), p(out), B, H, NC, G, _lib.current_stream())
    _lib.check(code, "ttt_b200_linear_forward")
    return out, (W1c, b1c), ((W1l, b1l) if want_last else None)


class TritonLinear(torch.autograd.Function):
    """Same name / call signature as the reference's TritonLinear (linear_triton.py:12)."""

    sharded_mode = False

    @staticmethod
    def forward(ctx, ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, XQ_batch, XV_batch, XK_batch, eta_batch,
                checkpoint_group_size):
        mp = XQ_batch.dtype
        bf = torch.bfloat16
        last_eta = eta_batch[:, :, :, -1, :]  # only the last row enters the scan (kernels/linear_forward.py:90-101)
        out, ck, _ = linear_forward(XQ_batch.to(bf).contiguous(), XK_batch.to(bf).contiguous(), XV_batch.to(bf).contiguous(),
                                    last_eta, ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, checkpoint_group_size)
        ctx.save_for_backward(XQ_batch, XV_batch, XK_batch, last_eta, ttt_norm_weight, ttt_norm_bias, *ck)
        return out.to(mp)

    @staticmethod
    def backward(ctx, grad_out):
        raise RuntimeError("ttt_b200: the TTT-Linear backward kernel is not built in this round (no eager fallback by design)")
