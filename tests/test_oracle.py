"""CPU: the oracle restatement is pinned to outputs of the unmodified reference (tests/golden/*.pt, produced by
oracle/make_golden.py in the build container).  No GPU, no /root/reference needed."""
import os

import pytest
import torch

from oracle import ttt_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def _check_inputs(d, fx):
    for k, v in fx["in_checksum"].items():
        assert abs(float(d[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, abs(v)), f"RNG drift in {k}"


@pytest.mark.parametrize("idx", [0, 1])
def test_mlp_eager_and_primal_match_reference(idx):
    fx = _load("ttt_mlp_ref.pt")[idx]
    c = fx["cfg"]
    d = O.make_inputs(c["B"], c["H"], c["NC"], CS=64, Fd=64, seed=c["seed"], dtype=torch.float64)
    _check_inputs(d, fx)
    grads, out = O.ttt_mlp_eager_grads(d["XQ"], d["XK"], d["XV"], d["eta"], d["ln_w"], d["ln_b"], d["W1"], d["b1"], d["W2"], d["b2"], d["dOut"])
    assert O.rel_err(out, fx["out"]) < 1e-6
    for g, r in zip(grads[:9], fx["grads"][:9]):
        assert O.rel_err(g, r) < 1e-5
    assert O.rel_err(grads[9].sum(-2), fx["grads"][9]) < 1e-5
    # primal form == the reference (row-uniform eta), incl. the hand-derived backward
    le = d["eta"][:, :, :, -1, :, None]
    po, ck, last = O.ttt_mlp_primal_forward(d["XQ"], d["XK"], d["XV"], le, d["ln_w"], d["ln_b"], d["W1"], d["b1"], d["W2"], d["b2"], 2)
    assert O.rel_err(po, fx["out"]) < 1e-6
    assert ck[0].shape[2] == (c["NC"] + 1) // 2
    assert torch.equal(ck[0][:, :, 0], d["W1"])
    pb = O.ttt_mlp_primal_backward(d["XQ"], d["XK"], d["XV"], le, d["ln_w"], d["ln_b"], d["W1"], d["b1"], d["W2"], d["b2"], d["dOut"])
    for n, r in zip(["dln_w", "dln_b", "dW1", "db1", "dW2", "db2", "dXQ", "dXV", "dXK"], fx["grads"][:9]):
        assert O.rel_err(pb[n], r) < 1e-5, n
    assert O.rel_err(pb["dlast_eta"].squeeze(-1), fx["grads"][9]) < 1e-5


@pytest.mark.parametrize("idx", [0, 1])
def test_linear_matches_reference(idx):
    fx = _load("ttt_linear_ref.pt")[idx]
    c = fx["cfg"]
    d = O.make_inputs(c["B"], c["H"], c["NC"], CS=16, Fd=64, seed=c["seed"], dtype=torch.float64, base_lr=1.0, linear=True)
    _check_inputs(d, fx)
    grads, out = O.ttt_linear_eager_grads(d["XQ"], d["XK"], d["XV"], d["eta"], d["ln_w"], d["ln_b"], d["W1"], d["b1"], d["dOut"])
    assert O.rel_err(out, fx["out"]) < 1e-6
    for g, r in zip(grads[:7], fx["grads"][:7]):
        assert O.rel_err(g, r) < 1e-5
    assert O.rel_err(grads[7].sum(-2), fx["grads"][7]) < 1e-5
    le = d["eta"][:, :, :, -1, :, None]
    po, _ = O.ttt_linear_primal_forward(d["XQ"], d["XK"], d["XV"], le, d["ln_w"], d["ln_b"], d["W1"], d["b1"])
    assert O.rel_err(po, fx["out"]) < 1e-6


def test_seq_block_gate_and_attention_match_reference():
    fx = _load("seq_block_ref.pt")
    c = fx["cfg"]
    vid, txt = fx["vid"].double(), fx["txt"].double()
    P = {k: v.double() for k, v in fx["P"].items()}
    sin, cos = O.rope3d_tables(c["Hh"], c["Ww"], c["frames"], c["E"] // c["NH"])
    a = O.local_attention_block(vid, txt, P, c["NH"], c["TL"], c["Hh"] * c["Ww"], c["chunks"], c["attn_length"], c["prefix"],
                                sin.double(), cos.double(), c["ln_eps"])
    assert O.rel_err(a, fx["attn_ref"]) < 1e-5

    def stub(x):
        return torch.cumsum(x, dim=1) * 0.01 + torch.roll(x, 1, dims=-1) * 0.5
    al = {k: v.double() for k, v in fx["alphas"].items()}
    emb = torch.cat((txt, vid), dim=1)
    s = O.ssm_bidirectional(emb, stub, c["TL"] * c["chunks"], c["chunks"], True,
                            al["forward_ssm_gating_text.gating_alpha"], al["forward_ssm_gating_video.gating_alpha"],
                            al["backward_ssm_gating_text.gating_alpha"], al["backward_ssm_gating_video.gating_alpha"])
    assert O.rel_err(s, fx["ssm_ref"]) < 1e-5


def test_dual_equals_primal_only_for_row_uniform_eta():
    """SURVEY parity trap #1: with row-non-uniform eta the eager dual form and the last-row primal form differ."""
    d = O.make_inputs(1, 2, 2, seed=5, dtype=torch.float64)
    le = d["eta"][:, :, :, -1, :, None]
    e, _ = O.ttt_mlp_eager(d["XK"], d["XQ"], d["XV"], d["eta"], d["ln_w"], d["ln_b"], d["W1"], d["b1"], d["W2"], d["b2"])
    p, _, _ = O.ttt_mlp_primal_forward(d["XQ"], d["XK"], d["XV"], le, d["ln_w"], d["ln_b"], d["W1"], d["b1"], d["W2"], d["b2"], 1)
    assert O.rel_err(p, e.permute(0, 3, 1, 2, 4)) < 1e-12
    eta2 = d["eta"].clone()
    eta2[:, :, :, :32] *= 3.0  # rows differ
    e2, _ = O.ttt_mlp_eager(d["XK"], d["XQ"], d["XV"], eta2, d["ln_w"], d["ln_b"], d["W1"], d["b1"], d["W2"], d["b2"])
    assert O.rel_err(p, e2.permute(0, 3, 1, 2, 4)) > 1e-6


def test_process_input_restatement_matches_reference_fixture():
    """oracle.ttt_process_input == the reference's TTTBase.process_input (ttt_layer.py:252-306), single and multi scene."""
    for fx in torch.load(os.path.join(GOLD, "process_input_ref.pt"), weights_only=False):
        c = fx["cfg"]
        cos, sin = O.ttt_rope_tables(c["Hh"], c["Ww"], c["frames"], c["E"] // c["NH"])
        idx = O.interleave_index(c["L"], c["TL"], c["chunks"], c["init_offset"]) if c["chunks"] > 1 else None
        q, k, v, eta = O.ttt_process_input(fx["q0"], fx["k0"], fx["v0"], fx["logit"], cos, sin, fx["ln_w"], fx["ln_b"],
                                           c["TL"] * c["chunks"], c["base_lr"], c["CS"], idx)
        for a, n in ((q, "XQ"), (k, "XK"), (v, "XV")):
            assert O.rel_err(a, fx["ref"][n]) < 1e-5, n
        assert O.rel_err(eta[:, :, :, -1, :], fx["ref_last_eta"]) < 1e-5
        if idx is not None:  # the interleave makes eta rows non-uniform (SURVEY trap #1): the scan reads the LAST row
            assert not torch.allclose(eta[:, :, :, 0, :], eta[:, :, :, -1, :])


def test_output_epilogue_restatement_matches_reference_fixture():
    """oracle.ttt_output_epilogue == post_norm + undo_interleave of the reference module (ttt_layer.py:324-331)."""
    for fx in torch.load(os.path.join(GOLD, "process_input_ref.pt"), weights_only=False):
        c = fx["cfg"]
        uidx = O.undo_interleave_index(c["L"], c["TL"], c["chunks"], c["init_offset"], c["base_offset"]) if c["chunks"] > 1 else None
        assert O.rel_err(O.ttt_output_epilogue(fx["ep_in"], fx["pn_w"], fx["pn_b"], 1e-6, uidx), fx["ep_ref"]) < 1e-5



def test_seeded_and_chunked_backward_match_autograd():
    """The analytic backward with an upstream gradient of the FINAL state (the sequence-sharded hand-off, and what lets the
    oracle run group by group) == autograd through the primal forward with that extra loss term."""
    import torch
    from oracle import ttt_oracle as O
    d = O.make_inputs(1, 2, 5, seed=9, dtype=torch.float64)
    le = d["eta"][:, :, :, -1, :, None]
    names = ("XQ", "XK", "XV")
    leaves = {n: d[n].clone().requires_grad_(True) for n in names + ("ln_w", "ln_b", "W1", "b1", "W2", "b2")}
    lel = le.clone().requires_grad_(True)
    out, _, last = O.ttt_mlp_primal_forward(leaves["XQ"], leaves["XK"], leaves["XV"], lel, leaves["ln_w"], leaves["ln_b"],
                                            leaves["W1"], leaves["b1"], leaves["W2"], leaves["b2"], 1 << 30)
    g = torch.Generator().manual_seed(1)
    up = [torch.randn(t.shape, generator=g, dtype=torch.float64) for t in last]
    loss = (out * d["dOut"]).sum() + sum((a * b).sum() for a, b in zip(last, up))
    loss.backward()
    for group in (None, 2):
        args = (d["XQ"], d["XK"], d["XV"], le, d["ln_w"], d["ln_b"], d["W1"], d["b1"], d["W2"], d["b2"], d["dOut"])
        r = (O.ttt_mlp_primal_backward(*args, dW_last=up) if group is None
             else O.ttt_mlp_primal_backward_chunked(*args, group, dW_last=up))
        for k, ref in (("dXQ", leaves["XQ"].grad), ("dXK", leaves["XK"].grad), ("dXV", leaves["XV"].grad), ("dlast_eta", lel.grad),
                       ("dW1", leaves["W1"].grad), ("db1", leaves["b1"].grad), ("dW2", leaves["W2"].grad), ("db2", leaves["b2"].grad),
                       ("dln_w", leaves["ln_w"].grad), ("dln_b", leaves["ln_b"].grad)):
            assert O.rel_err(r[k].reshape(ref.shape), ref) < 1e-9, (group, k)
