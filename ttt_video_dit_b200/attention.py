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


class QKNormRope(torch.autograd.Function):
    """Fused attention prologue (csrc/attn_prologue.cu): per-head LayerNorm of q and k + RoPE on the video rows, one pass
    forward and one backward.  q, k: bf16 [B, T, H, 64]; weights / biases of q_norm and k_norm ([64] each); sin / cos tables
    [>= T - text_len, 64]."""

    @staticmethod
    def forward(ctx, q, k, qw, qb, kw, kb, sin, cos, text_len, eps):
        B, T, H, D = _check_qkv(q, k, k)
        f32 = lambda t: t.detach().to(device=q.device, dtype=torch.float32)
        gamma = torch.stack((f32(qw), f32(kw))).contiguous()
        beta = torch.stack((f32(qb), f32(kb))).contiguous()
        c, s = f32(cos).contiguous(), f32(sin).contiguous()
        if c.shape[0] < T - text_len or c.shape[-1] != 64:
            raise RuntimeError("QKNormRope: RoPE tables must be [>= video tokens of the segment, 64]")
        qo, ko = torch.empty_like(q), torch.empty_like(k)
        p = _lib.ptr
        code = _lib.lib().ttt_b200_qk_norm_rope(p(q), p(k), p(gamma), p(beta), p(c), p(s), p(qo), p(ko), B, T, H, int(text_len),
                                                float(eps), _lib.current_stream(q))
        _lib.check(code, "ttt_b200_qk_norm_rope")
        ctx.save_for_backward(q, k, gamma, c, s)
        ctx.meta = (int(text_len), float(eps), qw.dtype, qb.dtype, kw.dtype, kb.dtype)
        return qo, ko

    @staticmethod
    def backward(ctx, dqo, dko):
        q, k, gamma, c, s = ctx.saved_tensors
        text_len, eps, dt_qw, dt_qb, dt_kw, dt_kb = ctx.meta
        B, T, H, D = q.shape
        dqo, dko = dqo.to(torch.bfloat16).contiguous(), dko.to(torch.bfloat16).contiguous()
        dq, dk = torch.empty_like(q), torch.empty_like(k)
        dg = torch.empty(2, 64, device=q.device, dtype=torch.float32)
        db = torch.empty_like(dg)
        p = _lib.ptr
        code = _lib.lib().ttt_b200_qk_norm_rope_backward(p(q), p(k), p(gamma), p(c), p(s), p(dqo), p(dko), p(dq), p(dk), p(dg), p(db),
                                                         B, T, H, text_len, eps, _lib.current_stream(q))
        _lib.check(code, "ttt_b200_qk_norm_rope_backward")
        return dq, dk, dg[0].to(dt_qw), db[0].to(dt_qb), dg[1].to(dt_kw), db[1].to(dt_kb), None, None, None, None


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
    reference calls (a GPU library call, as the q/k/v/o GEMMs are -- measured 1.4x faster than our kernel on B200 today,
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
        # q/k LayerNorm + segment-local RoPE (dit.py:188-194): one fused pass instead of ~10 elementwise ops per tensor
        q, k = QKNormRope.apply(q.contiguous(), k.contiguous(), P["q_norm.weight"], P["q_norm.bias"], P["k_norm.weight"],
                                P["k_norm.bias"], sin, cos, text_length, ln_eps)
        core = sdpa_bthd if impl == "b200" else _sdpa_library
        a = core(q.contiguous(), k.contiguous(), v.contiguous()).reshape(B, T, E)
        a = F.linear(a, P["o.weight"], P["o.bias"])
        out_txt[:, ts:te] = a[:, :text_length]
        out_vid[:, s:e] += a[:, text_length:].float()
        cnt[:, s:e] += 1
    return torch.cat((out_txt, (out_vid / cnt).to(vid.dtype)), dim=1)
