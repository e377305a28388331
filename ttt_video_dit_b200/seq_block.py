"""Bidirectional gated TTT pass of the DiT block on libttt_b200.so.

Mirror of ``SeqModelingBlock._ssm_forward`` / ``_gate`` / ``SSMGating`` / ``_reverse_text_chunks``
(ttt/models/cogvideo/dit.py:90-103, 213-266):

    y1  = x  + tanh(a_fwd) * ssm(x)
    rev = perm(y1)                       (text chunks in reverse order, video tokens flipped)
    out = y1 + tanh(a_bwd) * perm(ssm(rev))

Each gate (+ the reversal) is ONE fused HBM pass instead of the reference's clones / flips / cats.
"""
import torch

from . import _lib


def _check(x, name):
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 3):
        raise RuntimeError(f"{name} must be a contiguous CUDA bf16 tensor [B, L, E]")


class GatedResidual(torch.autograd.Function):
    """out = res + tanh(alpha) * (perm_s ? perm(s) : s); optionally also returns perm(out)."""

    @staticmethod
    def forward(ctx, res, s, a_text, a_video, text_len, num_chunks, perm_s, want_rev):
        _check(res, "res"); _check(s, "s")
        B, L, E = res.shape
        at = a_text.detach().float().contiguous(); av = a_video.detach().float().contiguous()
        out = torch.empty_like(res)
        rev = torch.empty_like(res) if want_rev else None
        code = _lib.lib().ttt_b200_gate_forward(_lib.ptr(res), _lib.ptr(s), _lib.ptr(at), _lib.ptr(av), _lib.ptr(out),
                                                _lib.ptr(rev), B, L, E, int(text_len), int(num_chunks), int(perm_s),
                                                _lib.current_stream(res))
        _lib.check(code, "ttt_b200_gate_forward")
        ctx.save_for_backward(s, at, av)
        ctx.meta = (int(text_len), int(num_chunks), int(perm_s), a_text.dtype)
        ctx.mark_non_differentiable()
        return (out, rev) if want_rev else out

    @staticmethod
    def backward(ctx, dout, drev=None):
        s, at, av = ctx.saved_tensors
        text_len, num_chunks, perm_s, adt = ctx.meta
        B, L, E = s.shape
        dout = dout.to(torch.bfloat16).contiguous()
        drev = None if drev is None else drev.to(torch.bfloat16).contiguous()
        dres = torch.empty_like(s); ds = torch.empty_like(s)
        dat = torch.empty(E, device=s.device, dtype=torch.float32); dav = torch.empty_like(dat)
        code = _lib.lib().ttt_b200_gate_backward(_lib.ptr(dout), _lib.ptr(drev), _lib.ptr(s), _lib.ptr(at), _lib.ptr(av),
                                                 _lib.ptr(dres), _lib.ptr(ds), _lib.ptr(dat), _lib.ptr(dav), B, L, E,
                                                 text_len, num_chunks, perm_s, _lib.current_stream(dout))
        _lib.check(code, "ttt_b200_gate_backward")
        return dres, ds, dat.to(adt), dav.to(adt), None, None, None, None


def ssm_forward(emb, ssm, seq_text_length, num_chunks, is_multiscene, gate_fwd_text, gate_fwd_video, gate_bwd_text,
                gate_bwd_video):
    """``SeqModelingBlock._ssm_forward`` (dit.py:224-266).  ``ssm``: callable [B,L,E] -> [B,L,E] (the TTT layer, applied
    twice with the same parameters); the four gates are the ``gating_alpha`` vectors [E]."""
    nc = num_chunks if is_multiscene else 1
    y1, rev = GatedResidual.apply(emb, ssm(emb), gate_fwd_text, gate_fwd_video, seq_text_length, nc, False, True)
    return GatedResidual.apply(y1, ssm(rev), gate_bwd_text, gate_bwd_video, seq_text_length, nc, True, False)
