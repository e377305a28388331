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
        p(wl[0]), p(wl[1]), p(wl[2]), p(wl[3]), p(Out), B, H, NC, int(checkpoint_group_size), _lib.current_stream())
    _lib.check(code, "ttt_b200_mlp_forward")
    return Out


LAUNCHES_FWD = 1  # one persistent scan kernel per forward call
