"""Per-segment local self-attention of the DiT block on libttt_b200.so (forward; sm_100a tcgen05 FlashAttention-style).

``sdpa_bthd`` replaces ``F.scaled_dot_product_attention(q, k, v, is_causal=False)`` (ttt/models/cogvideo/dit.py:196-198)
for tensors kept in the Linear-output layout [B, T, H, 64]; ``local_attention`` mirrors
``SeqModelingBlock._attn_forward`` (dit.py:163-211): per segment i, tokens = text chunk i + latent frames
[12 i, 12 i + 13); q/k/v Linear -> per-head LayerNorm(q), (k) -> RoPE on the video part (segment-local positions) ->
attention -> o Linear; text rows written, video rows accumulated and divided by the overlap count.
``sdpa_bthd`` is differentiable (csrc/attn_bwd.cu); the Linears / LayerNorm / RoPE around the kernel are the same
library / elementwise ops the reference uses.
"""
import math

import torch
import torch.nn.functional as F

from . import _lib


def _check_qkv(q, k, v):
    B, T, H, D = q.shape
    if D != 64:
        raise RuntimeError("attention kernel is specialised for head_dim 64")
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous() and t.shape == q.shape):
            raise RuntimeError(f"{n} must be a contiguous CUDA bf16 tensor [B, T, H, 64]")
    return B, T, H, D


class _SDPA(torch.autograd.Function):
    """Forward saves the per-row log-sum-exp; backward = ttt_b200_attention_backward (csrc/attn_bwd.cu)."""

    @staticmethod
    def forward(ctx, q, k, v, scale):
        B, T, H, D = _check_qkv(q, k, v)
        out = torch.empty_like(q)
        lse = torch.empty(B, H, T, device=q.device, dtype=torch.float32)
        code = _lib.lib().ttt_b200_attention_forward_lse(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(out), _lib.ptr(lse),
                                                         B, T, H, scale, _lib.current_stream(q))
        _lib.check(code, "ttt_b200_attention_forward_lse")
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        B, T, H, D = q.shape
        dout = dout.to(torch.bfloat16).contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        delta = torch.empty_like(lse)
        p = _lib.ptr
        code = _lib.lib().ttt_b200_attention_backward(p(q), p(k), p(v), p(out), p(dout), p(lse), p(delta), p(dq), p(dk), p(dv),
                                                      B, T, H, ctx.scale, _lib.current_stream(q))
        _lib.check(code, "ttt_b200_attention_backward")
        return dq, dk, dv, None


def sdpa_bthd(q, k, v, scale=None):
    """q, k, v: bf16 [B, T, H, 64] contiguous -> out [B, T, H, 64] (non-causal softmax(q k^T * scale) v); differentiable."""
    B, T, H, D = _check_qkv(q, k, v)
    sc = float(scale if scale is not None else 1.0 / math.sqrt(D))
    if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
        return _SDPA.apply(q, k, v, sc)
    out = torch.empty_like(q)
    code = _lib.lib().ttt_b200_attention_forward(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(out), B, T, H, sc,
                                                 _lib.current_stream(q))
    _lib.check(code, "ttt_b200_attention_forward")
    return out


def _rotate_half(x):
    x = x.reshape(*x.shape[:-1], -1, 2)
    x1, x2 = x.unbind(-1)
    return torch.stack((-x2, x1), dim=-1).flatten(-2)


def _sdpa_library(q, k, v):
    """The call the reference makes (dit.py:196-198): torch's library SDPA on [b, h, t, d] views (cuDNN / flash backend)."""
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=None, dropout_p=0.0,
                                       is_causal=False)
    return o.transpose(1, 2).contiguous()


def local_attention(vid, text, P, num_heads, text_length, tokens_per_frame, num_chunks, attn_length, prefix_len, sin, cos,
                    ln_eps=1e-6, impl="b200"):
    """``_attn_forward`` (dit.py:163-211).  vid [B, Lv, E], text [B, Lt, E] bf16; P: dict with q/k/v/o ``.weight``/``.bias``
    and q_norm/k_norm ``.weight``/``.bias``; sin/cos: RoPE tables [(t h w), 64] (cogvideo/utils.py:388-425).
    ``impl``: "b200" = this repo's tcgen05 attention kernels (csrc/attn_fwd.cu, attn_bwd.cu); "library" = the library SDPA the
    reference calls (a GPU library call, as the q/k/v/o GEMMs are -- measured 1.6x faster than our kernel on B200 today,
    which is why the integration patch leaves it in place)."""
    if impl not in ("b200", "library"):
        raise ValueError(impl)
    if not vid.is_cuda:
        raise RuntimeError("local_attention: tensors must be CUDA tensors (there is no CPU path)")
    B, _, E = vid.shape
    D = E // num_heads
    out_vid = torch.zeros_like(vid, dtype=torch.float32)
    out_txt = torch.zeros_like(text)
    cnt = torch.zeros(vid.shape[0], vid.shape[1], 1, device=vid.device, dtype=torch.float32)
    for i in range(num_chunks):
        s = i * attn_length * tokens_per_frame
        e = (prefix_len + (i + 1) * attn_length) * tokens_per_frame
        ts, te = i * text_length, (i + 1) * text_length
        cur = torch.cat([text[:, ts:te], vid[:, s:e]], dim=1)
        T = cur.shape[1]
        q = F.linear(cur, P["q.weight"], P["q.bias"]).reshape(B, T, num_heads, D)
        k = F.linear(cur, P["k.weight"], P["k.bias"]).reshape(B, T, num_heads, D)
        v = F.linear(cur, P["v.weight"], P["v.bias"]).reshape(B, T, num_heads, D)
        q = F.layer_norm(q, (D,), P["q_norm.weight"].to(q.dtype), P["q_norm.bias"].to(q.dtype), ln_eps)
        k = F.layer_norm(k, (D,), P["k_norm.weight"].to(k.dtype), P["k_norm.bias"].to(k.dtype), ln_eps)
        Lv = T - text_length
        c, sn = cos[:Lv].to(q.dtype)[None, :, None, :], sin[:Lv].to(q.dtype)[None, :, None, :]
        q = torch.cat([q[:, :text_length], q[:, text_length:] * c + _rotate_half(q[:, text_length:]) * sn], dim=1)
        k = torch.cat([k[:, :text_length], k[:, text_length:] * c + _rotate_half(k[:, text_length:]) * sn], dim=1)
        core = sdpa_bthd if impl == "b200" else _sdpa_library
        a = core(q.contiguous(), k.contiguous(), v.contiguous()).reshape(B, T, E)
        a = F.linear(a, P["o.weight"], P["o.bias"])
        out_txt[:, ts:te] = a[:, :text_length]
        out_vid[:, s:e] += a[:, text_length:].float()
        cnt[:, s:e] += 1
    return torch.cat((out_txt, (out_vid / cnt).to(vid.dtype)), dim=1)
