"""GPU parity of the whole TTT layer forward (ttt_layer.ttt_layer_forward: process_input kernel -> scan kernel ->
output_norm kernel, Linears in bf16) vs the reference modules TTTMLP / TTTLinear run in eager fp64 by oracle/make_golden.py
(tests/golden/ttt_layer_ref.pt; ttt/models/ssm/ttt_layer.py:314-334).  bf16 Linears on both sides of the kernels: 3e-2."""
import os

import pytest
import torch

from oracle import ttt_oracle as O
from ttt_video_dit_b200 import ttt_layer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("which", [0, 1])
def test_layer_forward_matches_reference_module(which):
    fx = torch.load(os.path.join(GOLD, "ttt_layer_ref.pt"), weights_only=False)[which]
    c = fx["cfg"]
    lin = ("wq.", "wk.", "wv.", "wo.", "learnable_ttt_lr")
    P = {k: (v.to(torch.bfloat16) if k.startswith(lin) else v).cuda() for k, v in fx["P"].items()}
    cos, sin = O.ttt_rope_tables(c["Hh"], c["Ww"], c["frames"], c["E"] // c["NH"])
    with torch.no_grad():
        out = ttt_layer.ttt_layer_forward(fx["X"].to(torch.bfloat16).cuda(), P, cos, sin, c["TL"], c["CS"], c["base_lr"], c["group"],
                                          kind=fx["kind"])
    torch.cuda.synchronize()
    assert out.shape == fx["ref"].shape
    err = O.rel_err(out.float().cpu(), fx["ref"])
    assert err < 3e-2, (fx["kind"], err)


@pytest.mark.parametrize("which", [0, 1])
def test_layer_backward_matches_autograd_of_reference_module(which):
    """Gradients of the whole layer (input + every parameter) vs autograd through the reference's eager module."""
    fx = torch.load(os.path.join(GOLD, "ttt_layer_ref.pt"), weights_only=False)[which]
    c = fx["cfg"]
    lin = ("wq.", "wk.", "wv.", "wo.", "learnable_ttt_lr")
    P = {k: (v.to(torch.bfloat16) if k.startswith(lin) else v).cuda().requires_grad_(True) for k, v in fx["P"].items()}
    X = fx["X"].to(torch.bfloat16).cuda().requires_grad_(True)
    cos, sin = O.ttt_rope_tables(c["Hh"], c["Ww"], c["frames"], c["E"] // c["NH"])
    out = ttt_layer.ttt_layer_forward(X, P, cos, sin, c["TL"], c["CS"], c["base_lr"], c["group"], kind=fx["kind"])
    out.backward(fx["gout"].to(torch.bfloat16).cuda())
    torch.cuda.synchronize()
    errs = {"X": O.rel_err(X.grad.float().cpu(), fx["grads"]["X"])}
    for n, g in fx["grads"].items():
        if n != "X":
            errs[n] = O.rel_err(P[n].grad.float().cpu().reshape(g.shape), g)
    bad = {k: v for k, v in errs.items() if not (v < 5e-2)}  # bf16 Linears + bf16 activations between five kernels
    assert not bad, (fx["kind"], bad, errs)

