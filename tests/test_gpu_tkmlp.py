"""GPU: the TkMLP-compatible autograd.Function (reference signature, ttt/models/ssm/mlp_tk.py:13-26)."""
import pytest
import torch

from oracle import ttt_oracle as O
from ttt_video_dit_b200 import mlp_tk

pytestmark = pytest.mark.gpu


def test_tkmlp_forward_reference_signature():
    d = O.make_inputs(2, 2, 5, seed=21)
    dev = "cuda"
    bf = lambda t: t.to(torch.bfloat16).to(dev)
    args = [d["ln_w"].to(dev), d["ln_b"].to(dev), d["W1"].to(dev), d["b1"].to(dev), d["W2"].to(dev), d["b2"].to(dev),
            bf(d["XQ"]), bf(d["XV"]), bf(d["XK"]), bf(d["eta"]), 2]
    out = mlp_tk.TkMLP.apply(*args)
    assert out.dtype == torch.bfloat16 and out.shape == d["XQ"].shape
    r = lambda t: t.to(torch.bfloat16).float()
    ref, _ = O.ttt_mlp_eager(r(d["XK"]), r(d["XQ"]), r(d["XV"]), r(d["eta"]), d["ln_w"], d["ln_b"], d["W1"], d["b1"], d["W2"], d["b2"])
    assert O.rel_err(out.float().cpu(), ref.permute(0, 3, 1, 2, 4)) < 1e-2
    # last-row entry point gives the same bits
    out2 = mlp_tk.ttt_mlp_op(*args[:9], bf(d["eta"])[:, :, :, -1, :], 2)
    assert torch.equal(out, out2)


def test_tkmlp_rejects_fp32_like_reference():
    d = O.make_inputs(1, 1, 1, seed=1)
    with pytest.raises(AssertionError):
        mlp_tk.TkMLP.apply(d["ln_w"].cuda(), d["ln_b"].cuda(), d["W1"].cuda(), d["b1"].cuda(), d["W2"].cuda(), d["b2"].cuda(),
                           d["XQ"].cuda(), d["XV"].cuda(), d["XK"].cuda(), d["eta"].cuda(), 1)
