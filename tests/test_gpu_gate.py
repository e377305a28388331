"""GPU parity: fused gate / reversal kernels vs the oracle restatement of dit.py:213-266 (pinned to the reference's
SeqModelingBlock._ssm_forward by tests/golden/seq_block_ref.pt), forward and backward."""
import os

import pytest
import torch

from oracle import ttt_oracle as O
from ttt_video_dit_b200 import seq_block

pytestmark = pytest.mark.gpu


def stub(x):
    return torch.cumsum(x.float(), dim=1).to(x.dtype) * 0.01 + torch.roll(x, 1, dims=-1) * 0.5


@pytest.mark.parametrize("B,TL,chunks,Lv,E", [(2, 16, 2, 400, 128), (1, 12, 3, 77, 64), (1, 8, 1, 100, 3072)])
def test_ssm_forward_and_backward(B, TL, chunks, Lv, E):
    torch.manual_seed(0)
    L = TL + Lv
    x = torch.randn(B, L, E).to(torch.bfloat16)
    al = [0.1 + 0.3 * torch.randn(E) for _ in range(4)]
    g = torch.randn(B, L, E).to(torch.bfloat16)
    # oracle (fp32 on the same bf16 inputs)
    xo = x.float().requires_grad_(True)
    alo = [a.clone().requires_grad_(True) for a in al]
    ro = O.ssm_bidirectional(xo, lambda t: stub(t), TL, chunks, chunks > 1, *alo)
    ro.backward(g.float())
    # ours
    xc = x.cuda().requires_grad_(True)
    alc = [a.cuda().requires_grad_(True) for a in al]
    rc = seq_block.ssm_forward(xc, stub, TL, chunks, chunks > 1, *alc)
    rc.backward(g.cuda())
    torch.cuda.synchronize()
    assert O.rel_err(rc.float().cpu(), ro.detach()) < 1e-2
    assert O.rel_err(xc.grad.float().cpu(), xo.grad) < 2e-2
    for a, b in zip(alc, alo):
        assert O.rel_err(a.grad.float().cpu(), b.grad) < 2e-2


def test_gate_against_reference_fixture():
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "seq_block_ref.pt"), weights_only=False)
    c = fx["cfg"]
    emb = torch.cat((fx["txt"], fx["vid"]), dim=1).to(torch.bfloat16).cuda()
    al = fx["alphas"]

    def ref_stub(x):  # same stub as oracle/make_golden.py, evaluated in fp32
        xf = x.float()
        return (torch.cumsum(xf, dim=1) * 0.01 + torch.roll(xf, 1, dims=-1) * 0.5).to(torch.bfloat16)
    out = seq_block.ssm_forward(emb, ref_stub, c["TL"] * c["chunks"], c["chunks"], True,
                                al["forward_ssm_gating_text.gating_alpha"].cuda(), al["forward_ssm_gating_video.gating_alpha"].cuda(),
                                al["backward_ssm_gating_text.gating_alpha"].cuda(), al["backward_ssm_gating_video.gating_alpha"].cuda())
    assert O.rel_err(out.float().cpu(), fx["ssm_ref"]) < 2e-2
