"""GPU parity: TTT-Linear forward kernel (C-ABI) vs the eager oracle (ttt/models/ssm/ops/ttt_linear.py) and the
reference fixture.  Tolerance 1e-2 relative (bf16 kernel vs fp32 eager)."""
import os

import pytest
import torch

from oracle import ttt_oracle as O
from ttt_video_dit_b200 import linear_triton

pytestmark = pytest.mark.gpu


def run(d, G):
    bf = lambda t: t.to(torch.bfloat16).cuda().contiguous()
    q, k, v = bf(d["XQ"]), bf(d["XK"]), bf(d["XV"])
    le = bf(d["eta"][:, :, :, -1, :])
    out, ck, last = linear_triton.linear_forward(q, k, v, le, d["ln_w"].cuda(), d["ln_b"].cuda(), d["W1"].cuda(), d["b1"].cuda(),
                                                 G, want_last=True)
    torch.cuda.synchronize()
    return (q, k, v, le), out, ck, last


@pytest.mark.parametrize("B,H,NC,G", [(1, 1, 1, 1), (1, 2, 6, 4), (1, 3, 9, 4), (2, 2, 40, 16)])
def test_linear_forward_matches_oracle(B, H, NC, G):
    d = O.make_inputs(B, H, NC, CS=16, seed=60 + NC, base_lr=1.0, linear=True)
    (q, k, v, le), out, ck, last = run(d, G)
    ref, (W1l, b1l) = O.ttt_linear_primal_forward(q.float().cpu(), k.float().cpu(), v.float().cpu(), le.float().cpu()[..., None],
                                                  d["ln_w"], d["ln_b"], d["W1"], d["b1"])
    assert O.rel_err(out.float().cpu(), ref) < 1e-2
    assert O.rel_err(last[0].cpu(), W1l) < 1e-2
    assert O.rel_err(last[1].cpu(), b1l) < 1e-2
    assert torch.equal(ck[0][:, :, 0].cpu(), d["W1"])


def test_linear_reference_signature_and_fixture():
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ttt_linear_ref.pt"), weights_only=False)[0]
    c = fx["cfg"]
    d = O.make_inputs(c["B"], c["H"], c["NC"], CS=16, seed=c["seed"], base_lr=1.0, linear=True)
    bf = lambda t: t.to(torch.bfloat16).cuda()
    out = linear_triton.TritonLinear.apply(d["ln_w"].cuda(), d["ln_b"].cuda(), d["W1"].cuda(), d["b1"].cuda(),
                                           bf(d["XQ"]), bf(d["XV"]), bf(d["XK"]), bf(d["eta"]), 4)
    assert out.dtype == torch.bfloat16
    assert O.rel_err(out.float().cpu(), fx["out"]) < 2e-2
