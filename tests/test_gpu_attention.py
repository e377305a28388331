"""GPU parity: tcgen05 attention forward vs the math-form oracle (fp32 softmax attention on the same bf16 inputs) and
vs the reference's _attn_forward fixture (tests/golden/seq_block_ref.pt)."""
import os

import pytest
import torch

from oracle import ttt_oracle as O
from ttt_video_dit_b200 import attention

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T,H", [(1, 128, 1), (2, 300, 3), (1, 1000, 2), (1, 129, 1), (1, 2048, 4)])
def test_sdpa_matches_math_attention(B, T, H):
    g = torch.Generator().manual_seed(T)
    q, k, v = (torch.randn(B, T, H, 64, generator=g).to(torch.bfloat16) for _ in range(3))
    q = q * 2.0  # sharper softmax
    out = attention.sdpa_bthd(q.cuda(), k.cuda(), v.cuda())
    torch.cuda.synchronize()
    tr = lambda t: t.float().permute(0, 2, 1, 3)
    ref = O.sdpa_math(tr(q), tr(k), tr(v)).permute(0, 2, 1, 3)
    assert O.rel_err(out.float().cpu(), ref) < 1e-2
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("B,T,H", [(1, 128, 1), (2, 300, 3), (1, 1000, 2), (1, 129, 1), (1, 640, 2)])
def test_sdpa_backward_matches_autograd_of_math_attention(B, T, H):
    # gradient oracle = torch autograd through the fp32 math-form attention on the same bf16 inputs (what autograd through
    # F.scaled_dot_product_attention computes, dit.py:196-198); bf16 P / dS operands in the kernel -> 2e-2 relative
    g = torch.Generator().manual_seed(1000 + T)
    q, k, v, go = (torch.randn(B, T, H, 64, generator=g).to(torch.bfloat16) for _ in range(4))
    q = q * 2.0
    qc, kc, vc = (t.cuda().requires_grad_(True) for t in (q, k, v))
    out = attention.sdpa_bthd(qc, kc, vc)
    out.backward(go.cuda())
    torch.cuda.synchronize()
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    tr = lambda t: t.permute(0, 2, 1, 3)
    ref = O.sdpa_math(tr(qr), tr(kr), tr(vr)).permute(0, 2, 1, 3)
    ref.backward(go.float())
    assert O.rel_err(out.float().cpu(), ref.detach()) < 1e-2
    for name, a, b in (("dq", qc.grad, qr.grad), ("dk", kc.grad, kr.grad), ("dv", vc.grad, vr.grad)):
        assert torch.isfinite(a).all(), name
        assert O.rel_err(a.float().cpu(), b) < 2e-2, (name, O.rel_err(a.float().cpu(), b))


def test_local_attention_block_matches_reference_fixture():
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "seq_block_ref.pt"), weights_only=False)
    c = fx["cfg"]
    bf = lambda t: t.to(torch.bfloat16).cuda()
    P = {k: bf(v) for k, v in fx["P"].items()}
    sin, cos = O.rope3d_tables(c["Hh"], c["Ww"], c["frames"], c["E"] // c["NH"])
    # head_dim of the fixture block is 64 (E=128, 2 heads)
    out = attention.local_attention(bf(fx["vid"]), bf(fx["txt"]), P, c["NH"], c["TL"], c["Hh"] * c["Ww"], c["chunks"],
                                    c["attn_length"], c["prefix"], sin.cuda(), cos.cuda(), c["ln_eps"])
    torch.cuda.synchronize()
    assert O.rel_err(out.float().cpu(), fx["attn_ref"]) < 3e-2  # bf16 Linears + bf16 LN/RoPE on top of the kernel tolerance


@pytest.mark.parametrize("B,T,H,text_len", [(1, 160, 2, 32), (2, 333, 3, 0), (1, 128, 48, 128)])
def test_qk_norm_rope_prologue(B, T, H, text_len):
    """Fused q/k LayerNorm + segment-local RoPE (csrc/attn_prologue.cu) vs the reference formulas in torch fp32 + autograd
    (dit.py:188-194, cogvideo/utils.py:93-99): forward, input gradients and the four norm-parameter gradients."""
    from ttt_video_dit_b200 import attention
    g = torch.Generator().manual_seed(T + H)
    rn = lambda *s: torch.randn(*s, generator=g)
    q, k, dq, dk = (rn(B, T, H, 64).bfloat16() for _ in range(4))
    prm = [1 + 0.2 * rn(64), 0.2 * rn(64), 1 + 0.2 * rn(64), 0.2 * rn(64)]
    sin, cos = O.rope3d_tables(4, 4, 24, 64)
    eps = 1e-6

    def rot(x):
        x = x.reshape(*x.shape[:-1], 32, 2)
        a, b = x.unbind(-1)
        return torch.stack((-b, a), dim=-1).flatten(-2)

    def ref():
        leaves = [t.float().requires_grad_(True) for t in (q, k)] + [t.clone().requires_grad_(True) for t in prm]
        outs = []
        for x, w, b_ in ((leaves[0], leaves[2], leaves[3]), (leaves[1], leaves[4], leaves[5])):
            y = torch.nn.functional.layer_norm(x, (64,), w, b_, eps)
            Lv = T - text_len
            c, s_ = cos[:Lv][None, :, None, :], sin[:Lv][None, :, None, :]
            outs.append(torch.cat([y[:, :text_len], y[:, text_len:] * c + rot(y[:, text_len:]) * s_], dim=1))
        (outs[0] * dq.float()).sum().add((outs[1] * dk.float()).sum()).backward()
        return [o.detach() for o in outs], [t.grad for t in leaves]

    leaves = [t.cuda().requires_grad_(True) for t in (q, k)] + [t.cuda().requires_grad_(True) for t in prm]
    qo, ko = attention.QKNormRope.apply(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[5], sin.cuda(), cos.cuda(), text_len, eps)
    (qo.float() * dq.cuda().float()).sum().add((ko.float() * dk.cuda().float()).sum()).backward()
    torch.cuda.synchronize()
    r_out, r_g = ref()
    assert O.rel_err(qo.float().cpu(), r_out[0]) < 1e-2 and O.rel_err(ko.float().cpu(), r_out[1]) < 1e-2
    for n, a, b_ in zip(("dq", "dk", "dqw", "dqb", "dkw", "dkb"), [t.grad for t in leaves], r_g):
        assert O.rel_err(a.float().cpu(), b_) < 1.5e-2, (n, O.rel_err(a.float().cpu(), b_))
