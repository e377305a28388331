"""GPU parity: TTT-Linear backward (trajectory + reverse kernels through the C-ABI) vs autograd through the eager oracle
(= torch autograd through ttt/models/ssm/ops/ttt_linear.py, pinned by tests/golden/ttt_linear_ref.pt).  Tolerance 1e-2
relative per tensor (north_star); d eta is compared after summing the eager gradient over rows."""
import os

import pytest
import torch

from oracle import ttt_oracle as O
from ttt_video_dit_b200 import linear_triton

pytestmark = pytest.mark.gpu
NAMES = ["d_ln_w", "d_ln_b", "dW1", "db1", "dXQ", "dXV", "dXK", "d_eta"]


def run_fwd_bwd(d, G):
    dev = "cuda"
    bf = lambda t: t.to(torch.bfloat16).to(dev)
    leaf = lambda t: t.clone().to(dev).requires_grad_(True)
    prm = [leaf(d[k]) for k in ("ln_w", "ln_b", "W1", "b1")]
    q, v, k, e = [bf(d[n]).requires_grad_(True) for n in ("XQ", "XV", "XK", "eta")]
    out = linear_triton.TritonLinear.apply(*prm, q, v, k, e, G)
    out.backward(bf(d["dOut"]))
    torch.cuda.synchronize()
    return out, [t.grad for t in prm] + [q.grad, v.grad, k.grad, e.grad]


def oracle_grads(d):
    r = lambda t: t.to(torch.bfloat16).float()
    grads, out = O.ttt_linear_eager_grads(r(d["XQ"]), r(d["XK"]), r(d["XV"]), r(d["eta"]), d["ln_w"], d["ln_b"], d["W1"],
                                          d["b1"], r(d["dOut"]))
    grads[7] = grads[7].sum(dim=-2)  # eager spreads d eta over the CS rows; the scan reads (and reports) the last row
    return out, grads


def errors(d, G):
    out, g = run_fwd_bwd(d, G)
    ref_out, rg = oracle_grads(d)
    errs = {"out": O.rel_err(out.float().cpu(), ref_out)}
    assert torch.count_nonzero(g[7][:, :, :, :-1, :]) == 0  # only the last eta row carries gradient
    g[7] = g[7][:, :, :, -1, :]
    for n, a, b in zip(NAMES, g, rg):
        errs[n] = O.rel_err(a.float().cpu().reshape(b.shape), b)
    return errs


@pytest.mark.parametrize("B,H,NC,G", [(1, 1, 1, 1), (1, 2, 3, 2), (2, 2, 7, 3), (1, 3, 20, 16), (1, 1, 5, 1000)])
def test_backward_matches_autograd_of_eager(B, H, NC, G):
    d = O.make_inputs(B, H, NC, CS=16, seed=70 + NC, base_lr=1.0, linear=True)
    errs = errors(d, G)
    bad = {k: v for k, v in errs.items() if not (v < 1e-2)}
    assert not bad, f"rel errors above 1e-2: {bad} (all: {errs})"


def test_backward_several_recompute_windows():
    # NC > one 256-step window with group 16 -> two trajectory windows, carried state gradient through global memory
    d = O.make_inputs(1, 2, 300, CS=16, seed=77, base_lr=1.0, linear=True)
    errs = errors(d, 16)
    bad = {k: v for k, v in errs.items() if not (v < 1e-2)}
    assert not bad, f"rel errors above 1e-2: {bad} (all: {errs})"


def test_backward_golden_reference_fixture():
    for fx in torch.load(os.path.join(os.path.dirname(__file__), "golden", "ttt_linear_ref.pt"), weights_only=False):
        c = fx["cfg"]
        d = O.make_inputs(c["B"], c["H"], c["NC"], CS=16, seed=c["seed"], base_lr=1.0, linear=True)
        _, g = run_fwd_bwd(d, 4)
        g[7] = g[7][:, :, :, -1, :]
        for n, a, b in zip(NAMES, g, fx["grads"]):
            assert O.rel_err(a.float().cpu().reshape(b.shape), b) < 2e-2, n


def test_fp32_inputs_are_accepted():
    """The reference's TritonLinear takes fp32 or bf16 activations (linear_triton.py:79); ours computes on bf16 MMA operands
    either way (fp32 inputs are rounded once at the boundary) and returns outputs / gradients in the caller's dtype."""
    d = O.make_inputs(1, 2, 6, CS=16, seed=81, base_lr=1.0, linear=True)
    dev = "cuda"
    leaf = lambda t: t.clone().float().to(dev).requires_grad_(True)
    prm = [leaf(d[k]) for k in ("ln_w", "ln_b", "W1", "b1")]
    q, v, k, e = [leaf(d[n]) for n in ("XQ", "XV", "XK", "eta")]
    out = linear_triton.TritonLinear.apply(*prm, q, v, k, e, 4)
    assert out.dtype == torch.float32
    out.backward(d["dOut"].float().to(dev))
    torch.cuda.synchronize()
    assert all(t.grad is not None and t.grad.dtype == torch.float32 for t in (q, v, k, e))
    ref_out, rg = oracle_grads(d)
    assert O.rel_err(out.cpu(), ref_out) < 1e-2
    assert O.rel_err(q.grad.cpu(), rg[4]) < 1e-2 and O.rel_err(k.grad.cpu(), rg[6]) < 1e-2
