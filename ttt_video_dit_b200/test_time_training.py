"""Mirror of the reference's native module ``test_time_training`` (ttt-tk/test_time_training.cpp:95-105).

Same function names, argument order and in-place/ownership semantics as the pybind11 module the reference imports in
ttt/models/ssm/mlp_tk.py:77,156 -- so ``import test_time_training as ttt_mlp`` can be pointed here unchanged -- but
implemented on libttt_b200.so (sm_100a, tcgen05) through the C-ABI.  Differences, all deliberate:
  * no cudaDeviceSynchronize after the launch (ttt-tk/kernels/ttt/ttt.cu:714-715 blocks; we stay async on the
    current stream);
  * CUDA errors raise (the reference only printf's them, ttt.cu:708-718);
  * LayerNorm eps is 1e-8 in forward AND backward (eager value; SURVEY parity trap #2).
"""
import torch

from . import _lib


def _chk(t, name, dtype, shape=None):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")


def ttt_forward(XQ, XK, XV, last_eta, ttt_norm_weight, ttt_norm_bias, W1, b1, W2, b2,
                W1_checkpoints, b1_checkpoints, W2_checkpoints, b2_checkpoints, Out, checkpoint_group_size,
                W_last=None):
    """test_time_training.ttt_forward (test_time_training.cpp:25-42).  Writes Out and the checkpoints in place and
    returns Out.  ``W_last`` (extension): optional tuple (W1,b1,W2,b2) of fp32 buffers receiving the final state."""
    B, H, NC, CS, F = XQ.shape
    if CS != 64 or F != 64:
        raise RuntimeError("TTT-MLP kernel is specialised for mini_batch_size 64, head_dim 64 (ttt.cu:20,612-613)")
    K = (NC + checkpoint_group_size - 1) // checkpoint_group_size
    bf, f32 = torch.bfloat16, torch.float32
    for t, n in ((XQ, "XQ"), (XK, "XK"), (XV, "XV"), (Out, "Out")):
        _chk(t, n, bf, (B, H, NC, CS, F))
    _chk(last_eta, "last_eta", bf, (B, H, NC, CS, 1))
    _chk(ttt_norm_weight, "ttt_norm_weight", f32, (1, H, 1, F))
    _chk(ttt_norm_bias, "ttt_norm_bias", f32, (1, H, 1, F))
    _chk(W1, "W1", f32, (B, H, F, 4 * F)); _chk(b1, "b1", f32, (B, H, 1, 4 * F))
    _chk(W2, "W2", f32, (B, H, 4 * F, F)); _chk(b2, "b2", f32, (B, H, 1, F))
    _chk(W1_checkpoints, "W1_checkpoints", f32, (B, H, K, F, 4 * F)); _chk(b1_checkpoints, "b1_checkpoints", f32, (B, H, K, 1, 4 * F))
    _chk(W2_checkpoints, "W2_checkpoints", f32, (B, H, K, 4 * F, F)); _chk(b2_checkpoints, "b2_checkpoints", f32, (B, H, K, 1, F))
    wl = [None] * 4
    if W_last is not None:
        wl = list(W_last)
        for t, n, s in zip(wl, ("W1_last", "b1_last", "W2_last", "b2_last"),
                           ((B, H, F, 4 * F), (B, H, 1, 4 * F), (B, H, 4 * F, F), (B, H, 1, F))):
            _chk(t, n, f32, s)
    p = _lib.ptr
    code = _lib.lib().ttt_b200_mlp_forward(
        p(XQ), p(XK), p(XV), p(last_eta), p(ttt_norm_weight), p(ttt_norm_bias), p(W1), p(b1), p(W2), p(b2),
        p(W1_checkpoints), p(b1_checkpoints), p(W2_checkpoints), p(b2_checkpoints),
        p(wl[0]), p(wl[1]), p(wl[2]), p(wl[3]), p(Out), B, H, NC, int(checkpoint_group_size), _lib.current_stream(XQ))
    _lib.check(code, "ttt_b200_mlp_forward")
    return Out


LAUNCHES_FWD = 1  # one persistent scan kernel per forward call
_last_call = {"groups": 1, "bh": 1, "seeded": False, "sms": 148}


def backward_is_persistent(bh, sms):
    """Mirror of the launcher's choice (csrc/ttt_mlp_bwd.cu): one K-side launch for the whole backward, hand-shaking with the
    recompute kernels through device-side flags, while the SMs can hold B*H K-side CTAs + B*H waiting trajectory CTAs + a
    few for the Q-side kernel; per-group launches ordered by events otherwise (or with TTT_B200_PERSISTENT=0)."""
    import os
    return os.environ.get("TTT_B200_PERSISTENT", "1") != "0" and 2 * bh + 16 <= sms


def launches_bwd():
    """Our kernels of the last backward call: per checkpoint group one trajectory + one Q-side kernel, plus the sequential
    K-side kernel -- once (persistent mode) or once per group -- plus the seed kernel of a chained call (the memsets are
    library calls)."""
    c = _last_call
    k_side = 1 if backward_is_persistent(c["bh"], c["sms"]) else c["groups"]
    return 2 * c["groups"] + k_side + (1 if c["seeded"] else 0)


def _workspace(B, H, G, device):
    """Recompute workspace of one backward call.  Allocated per call from torch's caching allocator (which makes the
    reuse cheap and stream-ordered): nothing is pinned for the life of the process, and the block returns to the
    allocator as soon as the call's tensors die.  Safe with the orchestrator's side streams: every side-stream kernel
    of a call is consumed by a later main-stream kernel of the same call, so main-stream order covers them."""
    n = _lib.lib().ttt_b200_mlp_backward_workspace_bytes(B, H, G)
    return torch.empty(n, dtype=torch.uint8, device=device), n


def ttt_backward_simple(XQ, XK, XV, last_eta, ttt_norm_weight, ttt_norm_bias, W1_checkpoints, b1_checkpoints,
                        W2_checkpoints, b2_checkpoints, grad_out, checkpoint_group_size, dW_last=None):
    """Backward with our own buffer management.  Returns, in TkMLP.backward order (mlp_tk.py:282-294):
    (d ln_w [H,F], d ln_b [H,F], dW1 [B,H,F,4F], db1 [B,H,1,4F], dW2, db2, dXQ, dXV, dXK, d last_eta [B,H,NC,CS,1]).
    ``dW_last`` (extension): optional (dW1, db1, dW2, db2) fp32 = upstream gradient of the FINAL state -- zero in the
    reference (mlp_tk.py:179-182); the backward hand-off message of the sequence-sharded chain (seq_shard.py)."""
    B, H, NC, CS, F = XQ.shape
    dev = XQ.device
    f32, bf = torch.float32, torch.bfloat16
    K = (NC + checkpoint_group_size - 1) // checkpoint_group_size
    _chk(grad_out, "grad_out", bf, (B, H, NC, CS, F))
    _chk(W1_checkpoints, "W1_checkpoints", f32, (B, H, K, F, 4 * F))
    dlw = torch.empty(B, H, F, device=dev, dtype=f32); dlb = torch.empty(B, H, F, device=dev, dtype=f32)
    dW1 = torch.empty(B, H, F, 4 * F, device=dev, dtype=f32); db1 = torch.empty(B, H, 1, 4 * F, device=dev, dtype=f32)
    dW2 = torch.empty(B, H, 4 * F, F, device=dev, dtype=f32); db2 = torch.empty(B, H, 1, F, device=dev, dtype=f32)
    dq = torch.empty_like(XQ); dk = torch.empty_like(XQ); dv = torch.empty_like(XQ)
    de = torch.empty(B, H, NC, CS, 1, device=dev, dtype=bf)
    ws, n = _workspace(B, H, checkpoint_group_size, dev)
    p = _lib.ptr
    up = [None] * 4
    if dW_last is not None:
        up = [t.to(f32).contiguous() for t in dW_last]
        for t, nm, shp in zip(up, ("dW1_last", "db1_last", "dW2_last", "db2_last"),
                              ((B, H, F, 4 * F), (B, H, 1, 4 * F), (B, H, 4 * F, F), (B, H, 1, F))):
            _chk(t, nm, f32, shp)
    code = _lib.lib().ttt_b200_mlp_backward_seeded(
        p(XQ), p(XK), p(XV), p(last_eta), p(ttt_norm_weight), p(ttt_norm_bias),
        p(W1_checkpoints), p(b1_checkpoints), p(W2_checkpoints), p(b2_checkpoints), p(grad_out),
        p(up[0]), p(up[1]), p(up[2]), p(up[3]),
        p(dlw), p(dlb), p(dW1), p(db1), p(dW2), p(db2), p(de), p(dq), p(dk), p(dv), p(ws), n,
        B, H, NC, int(checkpoint_group_size), _lib.current_stream(XQ))
    _lib.check(code, "ttt_b200_mlp_backward")
    _last_call.update(groups=K, bh=B * H, seeded=dW_last is not None, sms=torch.cuda.get_device_properties(dev).multi_processor_count)
    return dlw.sum(0), dlb.sum(0), dW1, db1, dW2, db2, dq, dv, dk, de


def ttt_backward(XQ, XK, XV, last_eta, ttt_norm_weight, ttt_norm_bias, W1_checkpoints, b1_checkpoints, W2_checkpoints,
                 b2_checkpoints, Out, W1_init_group, b1_init_group, W2_init_group, b2_init_group, x_hat_ln_group,
                 std_ln_group, X2_group, Z1_group, Z1_bar_group, X2_bar_group, grad_l_wrt_Z2_group, grad_l_wrt_Z1_group,
                 x_hat_fused_group, grad_x_hat_fused_group, grad_output_fused_group, std_fused_group, grad_L_W1_last,
                 grad_L_b1_last, grad_L_W2_last, grad_L_b2_last, grad_L_XQW_mini_batch, grad_L_ttt_norm_weight,
                 grad_L_ttt_norm_bias, grad_L_W1_init, grad_L_b1_init, grad_L_W2_init, grad_L_b2_init, grad_L_last_eta,
                 grad_L_XQ, grad_L_XK, grad_L_XV, checkpoint_group_size):
    """test_time_training.ttt_backward with the reference's 43-argument signature (test_time_training.cpp:47-91).
    The 16 re-materialisation buffers (args 12-27) are accepted and left untouched: this implementation keeps its
    recompute state in a private L2-resident workspace.  grad_L_W*_last (zeros in the reference, mlp_tk.py:179-182) seed the
    carried state gradient.  Gradients are written in place into the caller's buffers
    (the reference accumulates into pre-zeroed buffers, mlp_tk.py:213-225; writing gives the same result)."""
    r = ttt_backward_simple(XQ, XK, XV, last_eta, ttt_norm_weight, ttt_norm_bias, W1_checkpoints, b1_checkpoints,
                            W2_checkpoints, b2_checkpoints, grad_L_XQW_mini_batch.to(torch.bfloat16).contiguous(),
                            checkpoint_group_size,
                            dW_last=(grad_L_W1_last, grad_L_b1_last, grad_L_W2_last, grad_L_b2_last))
    B, H = XQ.shape[:2]
    # per-batch LN gradients: the reference buffer is [B,H,1,F]; we already reduced over B -> put the sum in row 0
    grad_L_ttt_norm_weight.zero_(); grad_L_ttt_norm_bias.zero_()
    grad_L_ttt_norm_weight[0, :, 0, :] = r[0]; grad_L_ttt_norm_bias[0, :, 0, :] = r[1]
    grad_L_W1_init.copy_(r[2]); grad_L_b1_init.copy_(r[3]); grad_L_W2_init.copy_(r[4]); grad_L_b2_init.copy_(r[5])
    grad_L_XQ.copy_(r[6]); grad_L_XV.copy_(r[7]); grad_L_XK.copy_(r[8]); grad_L_last_eta.copy_(r[9])
    return grad_L_XQ
