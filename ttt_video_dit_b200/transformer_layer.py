"""adaLN shell of the DiT ``TransformerLayer`` around the hot path, on libttt_b200.so.

Mirror of ttt/models/cogvideo/dit.py:321-382 (``TransformerLayer.forward``) and :268-278 (``SeqModelingBlock.forward``):

    (shift, scale, gate) x (video, text) = Linear(SiLU(t_emb))                        pre_seq_adaLN_modulation
    x_in  = modulate(LayerNorm(emb), shift, scale)                                    csrc/adaln.cu  ln_affine
    y     = local attention per 3-second segment, then bidirectional gated TTT        attention.py, seq_block.py, ttt_layer.py
    emb   = emb + gate * y                                                            csrc/adaln.cu  gate_add
    the same once more around the token MLP (Linear - GELU(tanh) - Linear)            cuBLAS GEMMs, as in the reference

The LayerNorm + modulate pair and the gated residual are each ONE HBM pass forward and one backward (the reference runs
LayerNorm, two broadcasts multiplies / adds and a cat per stream); the tiny per-batch vectors (A = gamma (1 + scale),
C = beta (1 + scale) + shift, the gates) are folded in torch so that autograd carries their chain rule to the adaLN Linear.
Tokens are kept as ONE tensor [B, L, E] with the text tokens first (what the reference builds by torch.cat at dit.py:369).
Parameters ``P`` use the reference module's state_dict names.  No eager fallback: every custom op raises on host tensors.
"""
from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib, attention, interleave, rope, seq_block, ttt_layer


def _check3(x, name):
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 3):
        raise RuntimeError(f"{name} must be a contiguous CUDA bf16 tensor [B, L, E]")


class LnAffine(torch.autograd.Function):
    """out = LayerNorm_noaffine(x) * A[b, seg] + C[b, seg]   (seg = text / video);  A, C: [B, 2, E]."""

    @staticmethod
    def forward(ctx, x, A, C, text_len, eps):
        _check3(x, "x")
        B, L, E = x.shape
        Af, Cf = A.detach().float().contiguous(), C.detach().float().contiguous()
        if Af.shape != (B, 2, E) or Cf.shape != (B, 2, E):
            raise RuntimeError("LnAffine: A and C must be [B, 2, E]")
        out = torch.empty_like(x)
        code = _lib.lib().ttt_b200_ln_affine(_lib.ptr(x), _lib.ptr(Af), _lib.ptr(Cf), _lib.ptr(out), B, L, E, int(text_len),
                                             float(eps), _lib.current_stream(x))
        _lib.check(code, "ttt_b200_ln_affine")
        ctx.save_for_backward(x, Af)
        ctx.meta = (int(text_len), float(eps), A.dtype, C.dtype)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, Af = ctx.saved_tensors
        text_len, eps, adt, cdt = ctx.meta
        B, L, E = x.shape
        gout = gout.to(torch.bfloat16).contiguous()
        gx = torch.empty_like(x)
        dA = torch.empty(B, 2, E, device=x.device, dtype=torch.float32)
        dC = torch.empty_like(dA)
        code = _lib.lib().ttt_b200_ln_affine_backward(_lib.ptr(x), _lib.ptr(Af), _lib.ptr(gout), _lib.ptr(gx), _lib.ptr(dA),
                                                      _lib.ptr(dC), B, L, E, text_len, eps, _lib.current_stream(x))
        _lib.check(code, "ttt_b200_ln_affine_backward")
        return gx, dA.to(adt), dC.to(cdt), None, None


class GateAdd(torch.autograd.Function):
    """out = x + G[b, seg] * y;  G: [B, 2, E]."""

    @staticmethod
    def forward(ctx, x, y, G, text_len):
        _check3(x, "x"); _check3(y, "y")
        B, L, E = x.shape
        Gf = G.detach().float().contiguous()
        if Gf.shape != (B, 2, E) or y.shape != x.shape:
            raise RuntimeError("GateAdd: G must be [B, 2, E] and y shaped like x")
        out = torch.empty_like(x)
        code = _lib.lib().ttt_b200_gate_add(_lib.ptr(x), _lib.ptr(y), _lib.ptr(Gf), _lib.ptr(out), B, L, E, int(text_len),
                                            _lib.current_stream(x))
        _lib.check(code, "ttt_b200_gate_add")
        ctx.save_for_backward(y, Gf)
        ctx.meta = (int(text_len), G.dtype)
        return out

    @staticmethod
    def backward(ctx, gout):
        y, Gf = ctx.saved_tensors
        text_len, gdt = ctx.meta
        B, L, E = y.shape
        gout = gout.to(torch.bfloat16).contiguous()
        dy = torch.empty_like(y)
        dG = torch.empty(B, 2, E, device=y.device, dtype=torch.float32)
        code = _lib.lib().ttt_b200_gate_add_backward(_lib.ptr(gout), _lib.ptr(y), _lib.ptr(Gf), _lib.ptr(dy), _lib.ptr(dG), B, L, E,
                                                     text_len, _lib.current_stream(y))
        _lib.check(code, "ttt_b200_gate_add_backward")
        return gout, dy, dG.to(gdt), None


@dataclass
class LayerMeta:
    """Shape metadata of one sequence (the fields of the reference's SequenceMetadata, cogvideo/utils.py:220-238, and the
    ModelConfig entries the layer reads) plus the tables derived from them once (``prepare``)."""
    num_heads: int
    text_length: int          # text tokens per scene
    num_chunks: int           # scenes (3-second segments)
    num_frames: int           # latent frames of the whole video
    latent_height: int
    latent_width: int
    mini_batch_size: int = 64
    ttt_base_lr: float = 0.1
    scan_checkpoint_group_size: int = 16
    ssm_layer: str = "ttt_mlp"
    attn_length: int = 12
    prefix_temporal_length: int = 1
    layer_norm_eps: float = 1e-6
    theta: float = 10000.0
    attention_impl: str = "b200"   # "library": the library SDPA the reference calls, see attention.local_attention
    tables: dict = field(default_factory=dict, repr=False)

    @property
    def tokens_per_frame(self):
        return self.latent_height * self.latent_width

    @property
    def seq_text_length(self):
        return self.text_length * self.num_chunks

    @property
    def is_multiscene(self):
        return self.num_chunks > 1

    def prepare(self, device, head_dim=64):
        """RoPE tables of both consumers and the interleave gather indices (cogvideo/utils.py:16-26 offsets)."""
        t = self.tables
        if t.get("device") == device:
            return t
        L = self.seq_text_length + self.num_frames * self.tokens_per_frame
        t["attn_sin"], t["attn_cos"] = rope.attention_tables(self.latent_height, self.latent_width, self.num_frames, head_dim,
                                                             self.theta, device)
        t["ttt_cos"], t["ttt_sin"] = rope.ttt_tables(self.latent_height, self.latent_width, self.num_frames, head_dim, self.theta,
                                                     device)
        t["il"], t["undo"] = None, None
        if self.is_multiscene:
            per = self.num_frames // self.num_chunks
            base = per * self.tokens_per_frame + self.text_length
            init = (per + self.num_frames % per) * self.tokens_per_frame + self.text_length
            t["il"] = interleave.interleave_index(L, self.text_length, self.num_chunks, init, device)
            t["undo"] = interleave.undo_interleave_index(L, self.text_length, self.num_chunks, init, base, device)
        t["device"] = device
        return t


def _sub(P, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in P.items() if k.startswith(prefix)}


def seq_modeling_block_forward(emb, P, meta: LayerMeta):
    """``SeqModelingBlock.forward`` (dit.py:268-278) on one token tensor [B, L, E] (text first): local attention per segment,
    then forward TTT -> gate -> reversed TTT (same parameters) -> gate.  P: the block's state_dict."""
    t = meta.prepare(emb.device)
    Lt = meta.seq_text_length
    text, vid = emb[:, :Lt], emb[:, Lt:]
    y = attention.local_attention(vid, text, P, meta.num_heads, meta.text_length, meta.tokens_per_frame, meta.num_chunks,
                                  meta.attn_length, meta.prefix_temporal_length, t["attn_sin"], t["attn_cos"], meta.layer_norm_eps,
                                  impl=meta.attention_impl)
    Pt = _sub(P, "ssm.ttt.")

    def ssm(x):
        return ttt_layer.ttt_layer_forward(x, Pt, t["ttt_cos"], t["ttt_sin"], Lt, meta.mini_batch_size, meta.ttt_base_lr,
                                           meta.scan_checkpoint_group_size, kind=meta.ssm_layer, interleave_index=t["il"],
                                           undo_interleave_index=t["undo"])
    return seq_block.ssm_forward(y.contiguous(), ssm, Lt, meta.num_chunks, meta.is_multiscene,
                                 P["forward_ssm_gating_text.gating_alpha"], P["forward_ssm_gating_video.gating_alpha"],
                                 P["backward_ssm_gating_text.gating_alpha"], P["backward_ssm_gating_video.gating_alpha"])


def _modulation(t_emb, w, b):
    """adaLN vectors (dit.py:331-338): Linear(SiLU(t_emb)) -> (shift, scale, gate, text_shift, text_scale, text_gate)."""
    return F.linear(F.silu(t_emb), w, b).float().chunk(6, dim=1)


def _affine(gamma, beta, shift_v, scale_v, shift_t, scale_t):
    """[B, 2, E] A / C of ln_affine: modulate(LayerNorm(x), shift, scale) = x_hat * gamma (1 + scale) + beta (1 + scale) + shift."""
    g, b = gamma.float()[None], beta.float()[None]
    A = torch.stack((g * (1 + scale_t), g * (1 + scale_v)), dim=1)
    C = torch.stack((b * (1 + scale_t) + shift_t, b * (1 + scale_v) + shift_v), dim=1)
    return A, C


def transformer_layer_forward(emb, t_emb, P, meta: LayerMeta):
    """``TransformerLayer.forward`` (dit.py:321-382).  emb bf16 [B, L, E], text tokens first; t_emb [B, time_embed_dim];
    P: the layer's state_dict (reference names).  Returns the updated [B, L, E]."""
    _check3(emb, "emb")
    Lt, eps = meta.seq_text_length, meta.layer_norm_eps
    sh, sc, g, tsh, tsc, tg = _modulation(t_emb.to(P["pre_seq_adaLN_modulation.1.weight"].dtype),
                                          P["pre_seq_adaLN_modulation.1.weight"], P["pre_seq_adaLN_modulation.1.bias"])
    A, C = _affine(P["pre_seq_layernorm.weight"], P["pre_seq_layernorm.bias"], sh, sc, tsh, tsc)
    x = LnAffine.apply(emb, A, C, Lt, eps)
    y = seq_modeling_block_forward(x, _sub(P, "seq_modeling_block."), meta)
    emb = GateAdd.apply(emb, y.contiguous(), torch.stack((tg, g), dim=1), Lt)

    sh, sc, g, tsh, tsc, tg = _modulation(t_emb.to(P["pre_mlp_adaLN_modulation.1.weight"].dtype),
                                          P["pre_mlp_adaLN_modulation.1.weight"], P["pre_mlp_adaLN_modulation.1.bias"])
    A, C = _affine(P["pre_mlp_layernorm.weight"], P["pre_mlp_layernorm.bias"], sh, sc, tsh, tsc)
    x = LnAffine.apply(emb, A, C, Lt, eps)
    h = F.gelu(F.linear(x, P["mlp.layer1.weight"], P["mlp.layer1.bias"]), approximate="tanh")  # dit.py:69-76
    y = F.linear(h, P["mlp.layer2.weight"], P["mlp.layer2.bias"])
    return GateAdd.apply(emb, y.contiguous(), torch.stack((tg, g), dim=1), Lt)
