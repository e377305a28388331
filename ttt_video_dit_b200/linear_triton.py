"""Drop-in for ``ttt.models.ssm.linear_triton.TritonLinear`` (reference: ttt/models/ssm/linear_triton.py:12-362) on
libttt_b200.so -- hand-written sm_100a CUDA, no Triton.

``TritonLinear.apply(ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, XQ, XV, XK, eta, checkpoint_group_size)`` keeps
the reference's signature (Q, V, K order, linear_triton.py:16-27) and the saved checkpoints (W1/b1 every
``checkpoint_group_size`` mini-batches, linear_triton.py:87-88).  Forward and backward scans are both native
(csrc/ttt_linear_fwd.cu, csrc/ttt_linear_bwd.cu); there is deliberately no eager fallback.
"""
import math

import torch

from . import _lib

try:
    from functools import partial
    from torch.distributed._tensor import Shard
    from torch.distributed._tensor.experimental import local_map
except Exception:  # pragma: no cover
    Shard = None
    local_map = None


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
                                              p(W1l), p(b1l), p(out), B, H, NC, G, _lib.current_stream(XQ))
    _lib.check(code, "ttt_b200_linear_forward")
    return out, (W1c, b1c), ((W1l, b1l) if want_last else None)


def linear_backward(XQ, XK, XV, last_eta, ln_w, ln_b, W1c, b1c, dOut, checkpoint_group_size):
    """Native TTT-Linear backward (reference: linear_triton.py:148-265).  Tensors as in linear_forward; dOut bf16
    [B,H,NC,16,64].  Returns (d ln_w [H,64], d ln_b [H,64], dW1 [B,H,64,64], db1 [B,H,1,64], dXQ, dXV, dXK (bf16),
    d last_eta f32 [B,H,NC,16])."""
    B, H, NC, CS, F = XQ.shape
    dev = XQ.device
    G = min(int(checkpoint_group_size), NC)
    le = last_eta.to(torch.bfloat16).reshape(B, H, NC, CS).contiguous()
    lw = ln_w.detach().reshape(H, F).float().contiguous()
    lb = ln_b.detach().reshape(H, F).float().contiguous()
    go = dOut.to(torch.bfloat16).contiguous()
    f32 = dict(device=dev, dtype=torch.float32)
    dlw = torch.empty(B, H, F, **f32)
    dlb = torch.empty(B, H, F, **f32)
    dW1 = torch.empty(B, H, F, F, **f32)
    db1 = torch.empty(B, H, 1, F, **f32)
    de = torch.empty(B, H, NC, CS, **f32)
    dq, dk, dv = torch.empty_like(XQ), torch.empty_like(XK), torch.empty_like(XV)
    L = _lib.lib()
    nbytes = L.ttt_b200_linear_backward_workspace_bytes(B, H, NC, G)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    p = _lib.ptr
    code = L.ttt_b200_linear_backward(p(XQ), p(XK), p(XV), p(le), p(lw), p(lb), p(W1c), p(b1c), p(go), p(dlw), p(dlb),
                                      p(dW1), p(db1), p(de), p(dq), p(dk), p(dv), p(ws), nbytes, B, H, NC, G,
                                      _lib.current_stream(XQ))
    _lib.check(code, "ttt_b200_linear_backward")
    return dlw.sum(0), dlb.sum(0), dW1, db1, dq, dv, dk, de


class _TTTLinearLastEta(torch.autograd.Function):
    """TTT-Linear op on the [B,H,NC,CS] eta row (what ``process_input.prepare`` produces): no [B,H,NC,CS,CS] tensor."""

    @staticmethod
    def forward(ctx, ln_w, ln_b, W1, b1, XQ, XV, XK, last_eta, checkpoint_group_size):
        bf = torch.bfloat16
        out, ck, _ = linear_forward(XQ.to(bf).contiguous(), XK.to(bf).contiguous(), XV.to(bf).contiguous(), last_eta, ln_w, ln_b,
                                    W1, b1, checkpoint_group_size)
        ctx.save_for_backward(XQ, XV, XK, last_eta, ln_w, ln_b, *ck)
        ctx.group = int(checkpoint_group_size)
        return out.to(XQ.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        XQ, XV, XK, last_eta, ln_w, ln_b, W1c, b1c = ctx.saved_tensors
        bf = torch.bfloat16
        dlw, dlb, dW1, db1, dq, dv, dk, de = linear_backward(XQ.to(bf).contiguous(), XK.to(bf).contiguous(), XV.to(bf).contiguous(),
                                                             last_eta, ln_w, ln_b, W1c, b1c, grad_out, ctx.group)
        mp = XQ.dtype
        return (dlw.reshape(ln_w.shape).to(ln_w.dtype), dlb.reshape(ln_b.shape).to(ln_b.dtype), dW1.to(W1c.dtype),
                db1.to(b1c.dtype), dq.to(mp), dv.to(mp), dk.to(mp), de.reshape(last_eta.shape).to(last_eta.dtype), None)


def ttt_linear_op(ln_w, ln_b, W1, b1, XQ, XV, XK, last_eta, checkpoint_group_size):
    return _TTTLinearLastEta.apply(ln_w, ln_b, W1, b1, XQ, XV, XK, last_eta, checkpoint_group_size)


class TritonLinear(torch.autograd.Function):
    """Same name / call signature as the reference's TritonLinear (linear_triton.py:12)."""

    sharded_mode = False

    @staticmethod
    def forward(ctx, ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, XQ_batch, XV_batch, XK_batch, eta_batch,
                checkpoint_group_size):
        fn = TritonLinear.forward_sharded if TritonLinear.sharded_mode else TritonLinear._forward_core
        return fn(ctx, ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, XQ_batch, XV_batch, XK_batch, eta_batch,
                  checkpoint_group_size)

    @staticmethod
    def backward(ctx, grad_out):
        fn = TritonLinear.backward_sharded if TritonLinear.sharded_mode else TritonLinear._backward_core
        return fn(ctx, grad_out)

    @staticmethod
    def _forward_core(ctx, ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, XQ_batch, XV_batch, XK_batch, eta_batch,
                      checkpoint_group_size):
        mp = XQ_batch.dtype
        bf = torch.bfloat16
        last_eta = eta_batch[:, :, :, -1, :]  # only the last row enters the scan (kernels/linear_forward.py:90-101)
        out, ck, _ = linear_forward(XQ_batch.to(bf).contiguous(), XK_batch.to(bf).contiguous(), XV_batch.to(bf).contiguous(),
                                    last_eta, ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, checkpoint_group_size)
        ctx.save_for_backward(XQ_batch, XV_batch, XK_batch, last_eta, ttt_norm_weight, ttt_norm_bias, *ck)
        ctx.group = int(checkpoint_group_size)
        ctx.eta_shape = eta_batch.shape
        return out.to(mp)

    @staticmethod
    def _backward_core(ctx, grad_out):
        XQ, XV, XK, last_eta, ln_w, ln_b, W1c, b1c = ctx.saved_tensors
        mp = XQ.dtype
        bf = torch.bfloat16
        dlw, dlb, dW1, db1, dq, dv, dk, de = linear_backward(XQ.to(bf).contiguous(), XK.to(bf).contiguous(),
                                                             XV.to(bf).contiguous(), last_eta, ln_w, ln_b, W1c, b1c,
                                                             grad_out, ctx.group)
        # the scan reads only the last eta row (kernels/linear_forward.py:90-101): the other rows get zero gradient
        d_eta = torch.zeros(ctx.eta_shape, device=XQ.device, dtype=mp)
        d_eta[:, :, :, -1, :] = de.to(mp)
        return (dlw.reshape(ln_w.shape).to(ln_w.dtype), dlb.reshape(ln_b.shape).to(ln_b.dtype), dW1.to(mp), db1.to(mp),
                dq.to(mp), dv.to(mp), dk.to(mp), d_eta, None)

    # --- head-sharded mode (reference linear_triton.py:262-362): entered through local_map with heads Shard(1) of the op
    #     inputs and Shard(0) of the [H,F] norm parameters; the kernels run on the local head shard, no collective inside.
    if local_map is not None:
        @staticmethod
        @partial(local_map, in_placements=(None, [Shard(0)], [Shard(0)], [Shard(1)], [Shard(1)], [Shard(1)], [Shard(1)],
                                           [Shard(1)], [Shard(1)], None),
                 out_placements=([Shard(1)],))
        def forward_sharded(ctx, *a):
            return TritonLinear._forward_core(ctx, *a)

        @staticmethod
        @partial(local_map, in_placements=(None, [Shard(1)]),
                 out_placements=([Shard(0)], [Shard(0)], [Shard(1)], [Shard(1)], [Shard(1)], [Shard(1)], [Shard(1)],
                                 [Shard(1)], None))
        def backward_sharded(ctx, g):
            return TritonLinear._backward_core(ctx, g)
    else:  # pragma: no cover
        forward_sharded, backward_sharded = _forward_core, _backward_core
