"""Drop-in for ``ttt.models.ssm.mlp_tk.TkMLP`` (reference: ttt/models/ssm/mlp_tk.py:9-404) on libttt_b200.so.

``TkMLP.apply(ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, W2_init, b2_init, XQ, XV, XK, eta, ckpt_group)``
keeps the reference's call signature (note the Q, V, K order, mlp_tk.py:13-26), dtype contract (bf16 in/out,
mlp_tk.py:89), the 10-gradients-plus-None backward return (mlp_tk.py:282-294) and the DTensor ``local_map``
sharded/unsharded split (mlp_tk.py:297-404), so ``ttt/models/ssm/ttt_layer.py:442-454`` can call it unchanged.

``ttt_mlp_op`` is the same op taking only the last eta row ``[B,H,NC,CS]`` -- it removes the need to materialise the
[B,H,NC,CS,CS] eta tensor (2.2 GB at 63 s).
"""
import math
from functools import partial

import torch

from . import test_time_training as ttt_native

try:  # DTensor plumbing exactly as the reference (mlp_tk.py:5-6)
    from torch.distributed._tensor import Shard
    from torch.distributed._tensor.experimental import local_map
except Exception:  # pragma: no cover
    Shard = None
    local_map = None

HAVE_BACKWARD = hasattr(ttt_native, "ttt_backward")


def launches_per_call(mode):
    """Number of OUR kernels launched by one op call (bench.py's gpu_launches claim)."""
    return ttt_native.LAUNCHES_FWD + (ttt_native.launches_bwd() if mode == "fwdbwd" else 0)


def _forward_impl(ctx, ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, W2_init, b2_init, XQ, XV, XK, last_eta, G):
    """mlp_tk.py:65-152 (_forward_core) minus the materialised eta; last_eta is [B,H,NC,CS,1] bf16 contiguous."""
    B, NH, NC, CS, F = XQ.shape
    K = math.ceil(NC / G)
    dev, mp = XQ.device, XQ.dtype
    assert mp == torch.bfloat16, "B200 TTT-MLP kernel runs in mixed-precision bfloat16 (reference: mlp_tk.py:89)."
    out = torch.empty(B, NH, NC, CS, F, device=dev, dtype=mp)
    W1c = torch.empty(B, NH, K, F, 4 * F, device=dev, dtype=torch.float32)
    b1c = torch.empty(B, NH, K, 1, 4 * F, device=dev, dtype=torch.float32)
    W2c = torch.empty(B, NH, K, 4 * F, F, device=dev, dtype=torch.float32)
    b2c = torch.empty(B, NH, K, 1, F, device=dev, dtype=torch.float32)
    XQ, XV, XK = XQ.contiguous(), XV.contiguous(), XK.contiguous()
    W1 = W1_init.to(torch.float32).contiguous(); b1 = b1_init.to(torch.float32).contiguous()
    W2 = W2_init.to(torch.float32).contiguous(); b2 = b2_init.to(torch.float32).contiguous()
    lw = ttt_norm_weight.detach().reshape(1, NH, 1, F).to(torch.float32).contiguous()
    lb = ttt_norm_bias.detach().reshape(1, NH, 1, F).to(torch.float32).contiguous()
    ttt_native.ttt_forward(XQ, XK, XV, last_eta, lw, lb, W1, b1, W2, b2, W1c, b1c, W2c, b2c, out, G)
    if ctx is not None:
        # the forward output is NOT saved: the native backward recomputes everything from the checkpoints (the
        # reference pins it, mlp_tk.py:146-150 -- 2 GB per layer-direction at 63 s for nothing)
        ctx.save_for_backward(XQ, XV, XK, last_eta, lw, lb, W1c, b1c, W2c, b2c)
        ctx.group = G
    return out


def _backward_impl(ctx, grad_out):
    """mlp_tk.py:154-294 (_backward_core).  Returns grads for (ln_w, ln_b, W1, b1, W2, b2, XQ, XV, XK, last_eta)."""
    if not HAVE_BACKWARD:
        raise RuntimeError("ttt_b200 backward kernel is not available in this build")
    XQ, XV, XK, last_eta, lw, lb, W1c, b1c, W2c, b2c = ctx.saved_tensors
    return ttt_native.ttt_backward_simple(XQ, XK, XV, last_eta, lw, lb, W1c, b1c, W2c, b2c,
                                          grad_out.to(torch.bfloat16).contiguous(), ctx.group)


class _TTTMLPLastEta(torch.autograd.Function):
    """Op on the last eta row only (extension; removes the [.., CS, CS] eta)."""

    @staticmethod
    def forward(ctx, ln_w, ln_b, W1, b1, W2, b2, XQ, XV, XK, last_eta, G):
        if last_eta.numel() != XQ.shape[0] * XQ.shape[1] * XQ.shape[2] * XQ.shape[3]:
            raise RuntimeError(f"last_eta must hold one value per token: expected [B,H,NC,CS] = {tuple(XQ.shape[:4])} "
                               f"(any shape with that many elements), got {tuple(last_eta.shape)}")
        if not last_eta.is_floating_point():
            raise RuntimeError(f"last_eta must be a floating-point tensor, got {last_eta.dtype}")
        le = last_eta.to(torch.bfloat16).reshape(*XQ.shape[:4], 1).contiguous()
        ctx.eta_shape = last_eta.shape
        return _forward_impl(ctx, ln_w, ln_b, W1, b1, W2, b2, XQ, XV, XK, le, G)

    @staticmethod
    def backward(ctx, grad_out):
        dlw, dlb, dW1, db1, dW2, db2, dq, dv, dk, de = _backward_impl(ctx, grad_out)
        mp = torch.bfloat16
        return (dlw.to(mp), dlb.to(mp), dW1.to(mp), db1.to(mp), dW2.to(mp), db2.to(mp), dq, dv, dk,
                de.reshape(ctx.eta_shape), None)


def ttt_mlp_op(ln_w, ln_b, W1, b1, W2, b2, XQ, XV, XK, last_eta, checkpoint_group_size):
    return _TTTMLPLastEta.apply(ln_w, ln_b, W1, b1, W2, b2, XQ, XV, XK, last_eta, checkpoint_group_size)


class TkMLP(torch.autograd.Function):
    """Same name / signature / return contract as the reference's TkMLP (mlp_tk.py:9)."""

    sharded_mode = False

    @staticmethod
    def forward(ctx, ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, W2_init, b2_init, XQ_batch, XV_batch, XK_batch,
                eta_batch, checkpoint_group_size):
        fn = TkMLP.forward_sharded if TkMLP.sharded_mode else TkMLP.forward_unsharded
        return fn(ctx, ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, W2_init, b2_init, XQ_batch, XV_batch, XK_batch,
                  eta_batch, checkpoint_group_size)

    @staticmethod
    def backward(ctx, grad_L_XQW_batch):
        fn = TkMLP.backward_sharded if TkMLP.sharded_mode else TkMLP.backward_unsharded
        return fn(ctx, grad_L_XQW_batch)

    @staticmethod
    def _forward_core(ctx, ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, W2_init, b2_init, XQ_batch, XV_batch,
                      XK_batch, eta_batch, checkpoint_group_size):
        mp = XQ_batch.dtype
        # only the last row of eta enters the kernel (reference: mlp_tk.py:104-105)
        last_eta = eta_batch.to(mp)[:, :, :, -1, :, None].contiguous()
        ctx.cs = eta_batch.shape[-2]
        return _forward_impl(ctx, ttt_norm_weight, ttt_norm_bias, W1_init, b1_init, W2_init, b2_init, XQ_batch, XV_batch,
                             XK_batch, last_eta, checkpoint_group_size)

    @staticmethod
    def _backward_core(ctx, grad_L_XQW_batch):
        dlw, dlb, dW1, db1, dW2, db2, dq, dv, dk, de = _backward_impl(ctx, grad_L_XQW_batch)
        mp = torch.bfloat16
        # d eta is non-zero only in the last row (reference: mlp_tk.py:280, pad of CS-1 rows)
        grad_eta = torch.nn.functional.pad(de.transpose(-2, -1), (0, 0, ctx.cs - 1, 0))
        return (dlw.to(mp), dlb.to(mp), dW1.to(mp), db1.to(mp), dW2.to(mp), db2.to(mp), dq.to(mp), dv.to(mp), dk.to(mp),
                grad_eta.to(mp), None)

    # --- local_map wrappers, placements as in the reference (mlp_tk.py:297-404): heads are Shard(1) of the op inputs,
    #     Shard(0) of the [H,F] norm parameters; no collective inside the op.
    if local_map is not None:
        @staticmethod
        @partial(local_map, in_placements=(None, [Shard(0)], [Shard(0)], [Shard(1)], [Shard(1)], [Shard(1)], [Shard(1)],
                                           [Shard(1)], [Shard(1)], [Shard(1)], [Shard(1)], None),
                 out_placements=([Shard(1)],))
        def forward_sharded(ctx, *a):
            return TkMLP._forward_core(ctx, *a)

        @staticmethod
        @partial(local_map, in_placements=None, out_placements=None)
        def forward_unsharded(ctx, *a):
            return TkMLP._forward_core(ctx, *a)

        @staticmethod
        @partial(local_map, in_placements=(None, [Shard(1)]),
                 out_placements=([Shard(0)], [Shard(0)], [Shard(1)], [Shard(1)], [Shard(1)], [Shard(1)], [Shard(1)],
                                 [Shard(1)], [Shard(1)], [Shard(1)], None))
        def backward_sharded(ctx, g):
            return TkMLP._backward_core(ctx, g)

        @staticmethod
        @partial(local_map, in_placements=None, out_placements=None)
        def backward_unsharded(ctx, g):
            return TkMLP._backward_core(ctx, g)
    else:  # pragma: no cover
        forward_sharded = forward_unsharded = _forward_core
        backward_sharded = backward_unsharded = _backward_core
