"""CPU oracle for the TTT hot path (TEST INFRASTRUCTURE -- never imported by the product).

This file restates, in plain torch-on-CPU (fp32 or fp64), the reference's own
PyTorch-eager TTT path.  It exists so that the CUDA kernels in
``ttt_video_dit_b200/csrc`` can be checked on a GPU box where ``/root/reference``
does not exist.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Parity pinning: ``oracle/make_golden.py`` imports the *unmodified* reference
(``ttt.models.ssm.ops.ttt_mlp`` / ``ttt_linear`` and ``SeqModelingBlock``) in the
build container, checks every function below against it and writes the small
fixtures in ``tests/golden/``.  ``tests/test_oracle.py`` re-checks the oracle
against those fixtures on every run, so the oracle is pinned to reference
outputs, not merely to itself.

Reference citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

LN_EPS = 1e-8  # ttt/models/ssm/ops/utils.py:4,21  (the eager value; TK fwd uses 1e-6, see SURVEY trap #2)


# --------------------------------------------------------------------------------------
# elementwise helpers  (ttt/models/ssm/ops/utils.py)
# --------------------------------------------------------------------------------------
def ln_fwd(x, gamma, beta, eps=LN_EPS):
    """ops/utils.py:4-18 -- LayerNorm over the last dim, biased variance, sqrt(var+eps)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = x.var(dim=-1, keepdim=True, unbiased=False)
    std = torch.sqrt(var + eps)
    return gamma * ((x - mu) / std) + beta


def ln_fused_l2_bwd(x, l2_target, gamma, beta, eps=LN_EPS):
    """ops/utils.py:21-48 -- d/dx of 0.5*||LN(x)-target||^2 in closed form."""
    D = x.shape[-1]
    mu = x.mean(dim=-1, keepdim=True)
    var = x.var(dim=-1, keepdim=True, unbiased=False)
    std = torch.sqrt(var + eps)
    x_hat = (x - mu) / std
    y = gamma * x_hat + beta
    grad_output = y - l2_target
    grad_x_hat = grad_output * gamma
    z = (
        (1.0 / D)
        * (D * grad_x_hat - grad_x_hat.sum(dim=-1, keepdim=True) - x_hat * (grad_x_hat * x_hat).sum(dim=-1, keepdim=True))
        / std
    )
    return z


def gelu_bwd(x):
    """ops/utils.py:51-54 -- derivative of tanh-GELU with the reference's truncated constants."""
    t = torch.tanh(0.79788456 * x * (1 + 0.044715 * x * x))
    return 0.5 * x * ((1 - t * t) * (0.79788456 + 0.1070322243 * x * x)) + 0.5 * (1 + t)


def gelu_bwd_derivative(x):
    """ttt-tk/kernels/ttt_backward/matching.py:47-55 -- second derivative of tanh-GELU."""
    t = torch.tanh(0.79788456 * x * (1 + 0.044715 * x * x))
    term1 = 0.79788456
    term2 = 6 * 0.79788456 * 0.044715 * x**2
    term3 = x * t * (0.79788456 + 3 * 0.79788456 * 0.044715 * x**2) ** 2
    return (1 - t**2) * (term1 + term2 - term3)


def gelu(x):
    return F.gelu(x, approximate="tanh")


# --------------------------------------------------------------------------------------
# TTT-MLP: eager dual form (the parity oracle)  (ttt/models/ssm/ops/ttt_mlp.py)
# --------------------------------------------------------------------------------------
def ttt_mlp_step_dual(p: Dict[str, torch.Tensor], XQ, XK, XV, eta):
    """One mini-batch, dual form.  ops/ttt_mlp.py:9-67.

    XQ/XK/XV: [B,H,CS,F]; eta: [B,H,CS,CS]; W1 [B,H,F,4F]; b1 [B,H,1,4F]; W2 [B,H,4F,F]; b2 [B,H,1,F].
    """
    W1, b1, W2, b2 = p["W1"], p["b1"], p["W2"], p["b2"]
    H, Fd = XQ.size(1), XQ.size(-1)
    ln_w = p["ln_w"].reshape(H, 1, Fd)
    ln_b = p["ln_b"].reshape(H, 1, Fd)

    Z1 = XK @ W1 + b1                                              # :29
    X2 = gelu(Z1)                                                  # :30
    Z2 = X2 @ W2 + b2                                              # :31
    target = XV - XK                                               # :32
    gZ2 = ln_fused_l2_bwd(Z2, target, ln_w, ln_b)                  # :36
    gZ1 = gZ2 @ W2.transpose(-2, -1) * gelu_bwd(Z1)                # :37

    Attn1 = XQ @ XK.transpose(-2, -1)                              # :39
    b1_bar = b1 - eta @ gZ1                                        # :40
    Z1_bar = XQ @ W1 - (eta * Attn1) @ gZ1 + b1_bar                # :41
    X2_bar = gelu(Z1_bar)                                          # :42
    Attn2 = X2_bar @ X2.transpose(-2, -1)                          # :44
    b2_bar = b2 - eta @ gZ2                                        # :45
    Z2_bar = X2_bar @ W2 - (eta * Attn2) @ gZ2 + b2_bar            # :46

    last_eta = eta[:, :, -1, :, None]                              # :48
    W1n = W1 - (last_eta * XK).transpose(-1, -2) @ gZ1             # :49
    b1n = b1 - torch.sum(last_eta * gZ1, dim=-2, keepdim=True)     # :50
    W2n = W2 - (last_eta * X2).transpose(-1, -2) @ gZ2             # :51
    b2n = b2 - torch.sum(last_eta * gZ2, dim=-2, keepdim=True)     # :52

    out = XQ + ln_fwd(Z2_bar, ln_w, ln_b)                          # :54-56
    return dict(W1=W1n, b1=b1n, W2=W2n, b2=b2n, ln_w=p["ln_w"], ln_b=p["ln_b"]), out


def ttt_mlp_eager(XK, XQ, XV, eta, ln_w, ln_b, W1, b1, W2, b2, checkpoint_group_size=0):
    """Driver with the reference's signature (note K,Q,V order).  ops/ttt_mlp.py:70-99 + ssm/utils.py:111-146.

    Inputs [B,H,NC,CS,*]; returns ([B,NC,CS,H,F], final params).  The reference's activation
    checkpointing (``scan`` groups) does not change values, so it is a plain loop here.
    """
    p = dict(W1=W1, b1=b1, W2=W2, b2=b2, ln_w=ln_w, ln_b=ln_b)
    NC = XK.shape[2]
    outs = []
    for n in range(NC):
        p, o = ttt_mlp_step_dual(p, XQ[:, :, n], XK[:, :, n], XV[:, :, n], eta[:, :, n])
        outs.append(o)
    out = torch.stack(outs, dim=0)            # [NC,B,H,CS,F]
    return out.permute(1, 0, 3, 2, 4), p       # ops/ttt_mlp.py:99


# --------------------------------------------------------------------------------------
# TTT-MLP: primal form with saved intermediates (the kernel's algebra)
#   ttt-tk/kernels/ttt/matching.py:30-125, ttt-tk/kernels/ttt_backward/matching.py:57-171
# --------------------------------------------------------------------------------------
def ttt_mlp_step_primal(W1, b1, W2, b2, XQ, XK, XV, last_eta, ln_w, ln_b, eps=LN_EPS):
    """last_eta: [B,H,CS,1] (the last row of eta as a column).  Returns new state, output and saved tensors."""
    H, Fd = XK.shape[1], XK.shape[-1]
    g = ln_w.reshape(H, 1, Fd)
    bt = ln_b.reshape(H, 1, Fd)
    Z1 = XK @ W1 + b1
    X2 = gelu(Z1)
    Z2 = X2 @ W2 + b2
    target = XV - XK
    mu = Z2.mean(-1, keepdim=True)
    var = Z2.var(-1, keepdim=True, unbiased=False)
    std_f = torch.sqrt(var + eps)
    xhat_f = (Z2 - mu) / std_f
    y = g * xhat_f + bt
    go = y - target
    gxh = go * g
    gZ2 = (1.0 / Fd) * (Fd * gxh - gxh.sum(-1, keepdim=True) - xhat_f * (gxh * xhat_f).sum(-1, keepdim=True)) / std_f
    gZ1 = gZ2 @ W2.transpose(-1, -2) * gelu_bwd(Z1)
    W1n = W1 - (last_eta * XK).transpose(-1, -2) @ gZ1
    b1n = b1 - (last_eta * gZ1).sum(-2, keepdim=True)
    W2n = W2 - (last_eta * X2).transpose(-1, -2) @ gZ2
    b2n = b2 - (last_eta * gZ2).sum(-2, keepdim=True)
    Z1b = XQ @ W1n + b1n
    X2b = gelu(Z1b)
    Z2b = X2b @ W2n + b2n
    mu_o = Z2b.mean(-1, keepdim=True)
    var_o = Z2b.var(-1, keepdim=True, unbiased=False)
    std_o = torch.sqrt(var_o + eps)
    xhat_o = (Z2b - mu_o) / std_o
    out = XQ + g * xhat_o + bt
    saved = dict(Z1=Z1, X2=X2, std_f=std_f, xhat_f=xhat_f, go=go, gxh=gxh, gZ2=gZ2, gZ1=gZ1,
                 Z1b=Z1b, X2b=X2b, std_o=std_o, xhat_o=xhat_o)
    return (W1n, b1n, W2n, b2n), out, saved


def ttt_mlp_primal_forward(XQ, XK, XV, last_eta, ln_w, ln_b, W1, b1, W2, b2, checkpoint_group_size):
    """Whole-sequence primal forward with the native ABI's semantics (Q,K,V order; checkpoints of the
    state *entering* each group).  ttt-tk/kernels/ttt/ttt.cu:276-584 (algorithm), mlp_tk.py:92-98 (buffers).

    last_eta: [B,H,NC,CS,1].  Returns Out [B,H,NC,CS,F], checkpoints (W1c [B,H,K,F,4F] ...), final state.
    """
    B, H, NC, CS, Fd = XQ.shape
    K = math.ceil(NC / checkpoint_group_size)
    W1c = torch.empty(B, H, K, *W1.shape[2:], dtype=W1.dtype)
    b1c = torch.empty(B, H, K, *b1.shape[2:], dtype=W1.dtype)
    W2c = torch.empty(B, H, K, *W2.shape[2:], dtype=W1.dtype)
    b2c = torch.empty(B, H, K, *b2.shape[2:], dtype=W1.dtype)
    out = torch.empty_like(XQ)
    st = (W1, b1, W2, b2)
    for n in range(NC):
        if n % checkpoint_group_size == 0:
            k = n // checkpoint_group_size
            W1c[:, :, k], b1c[:, :, k], W2c[:, :, k], b2c[:, :, k] = st
        st, o, _ = ttt_mlp_step_primal(*st, XQ[:, :, n], XK[:, :, n], XV[:, :, n], last_eta[:, :, n], ln_w, ln_b)
        out[:, :, n] = o
    return out, (W1c, b1c, W2c, b2c), st


def ttt_mlp_step_backward(XQ, XK, W1, W2, W1n, W2n, last_eta, ln_w, ln_b, s, dW1n, db1n, dW2n, db2n, dO):
    """Closed-form backward of one primal step.  ttt-tk/kernels/ttt_backward/matching.py:173-342.

    ``s`` = saved dict of ttt_mlp_step_primal; dW*n/db*n = gradient w.r.t. the state *after* this step.
    Returns (dln_w[H,F], dln_b[H,F], dW1, db1, dW2, db2, dXQ, dXV, dXK, dlast_eta[B,H,CS,1]).
    """
    H, Fd = XQ.shape[1], XQ.shape[-1]
    g = ln_w.reshape(H, 1, Fd)
    eta = last_eta
    # stage 4: output LN  (:213-225)
    dbeta_o = dO.sum(-2, keepdim=True).sum(0)
    dgamma_o = (dO * s["xhat_o"]).sum(-2, keepdim=True).sum(0)
    dxh = dO * g
    dZ2b = (1.0 / Fd) * (Fd * dxh - dxh.sum(-1, keepdim=True) - s["xhat_o"] * (dxh * s["xhat_o"]).sum(-1, keepdim=True)) / s["std_o"]
    # stage 3  (:228-262)
    dX2b = dZ2b @ W2n.transpose(-2, -1)
    dZ1b = dX2b * gelu_bwd(s["Z1b"])
    db2n = db2n + dZ2b.sum(-2, keepdim=True)
    dW2n = dW2n + s["X2b"].transpose(-2, -1) @ dZ2b
    db1n = db1n + dZ1b.sum(-2, keepdim=True)
    dW1n = dW1n + XQ.transpose(-2, -1) @ dZ1b
    dgZ1 = -(eta * XK) @ dW1n - eta * db1n
    dgZ2 = -(eta * s["X2"]) @ dW2n - eta * db2n + (dgZ1 * gelu_bwd(s["Z1"])) @ W2
    dXQ_u = dZ1b @ W1n.transpose(-2, -1)
    T1 = s["gZ1"] @ dW1n.transpose(-2, -1)
    T2 = s["gZ2"] @ dW2n.transpose(-2, -1)
    dXK_u = -T1 * eta
    deta = (-(T2 * s["X2"]).sum(-1, keepdim=True) - (db2n * s["gZ2"]).sum(-1, keepdim=True)
            - (T1 * XK).sum(-1, keepdim=True) - (db1n * s["gZ1"]).sum(-1, keepdim=True))
    # stage 2  (:269-305)
    dW2_a = (dgZ1 * gelu_bwd(s["Z1"])).transpose(-2, -1) @ s["gZ2"]
    dgxh = (1.0 / s["std_f"]) * (dgZ2 + (-1.0 / Fd) * (dgZ2.sum(-1, keepdim=True) + s["xhat_f"] * (dgZ2 * s["xhat_f"]).sum(-1, keepdim=True)))
    dy = g * dgxh
    dgamma_f = (s["go"] * dgxh + dy * s["xhat_f"]).sum(-2, keepdim=True).sum(0)
    dbeta_f = dy.sum(-2, keepdim=True).sum(0)
    dxh_f = dy * g + (-1.0 / (Fd * s["std_f"])) * (s["gxh"] * (dgZ2 * s["xhat_f"]).sum(-1, keepdim=True)
                                                    + dgZ2 * (s["gxh"] * s["xhat_f"]).sum(-1, keepdim=True))
    dstd = (-dxh_f * s["xhat_f"] - dgZ2 * s["gZ2"]) / s["std_f"]
    dZ2 = dxh_f / s["std_f"] + (1.0 / Fd) * (dstd.sum(-1, keepdim=True) * s["xhat_f"] - dxh_f.sum(-1, keepdim=True) / s["std_f"])
    dtarget = -g * dgxh
    # stage 1  (:311-327)
    dX2 = dZ2 @ W2.transpose(-2, -1) - T2 * eta
    dZ1 = dX2 * gelu_bwd(s["Z1"]) + (s["gZ2"] @ W2.transpose(-2, -1)) * dgZ1 * gelu_bwd_derivative(s["Z1"])
    dXQ = dO + dXQ_u
    dXK = -dtarget + dXK_u + dZ1 @ W1.transpose(-2, -1)
    dXV = dtarget
    dW2 = dW2n + dW2_a + s["X2"].transpose(-2, -1) @ dZ2
    db2 = db2n + dZ2.sum(-2, keepdim=True)
    dW1 = dW1n + XK.transpose(-2, -1) @ dZ1
    db1 = db1n + dZ1.sum(-2, keepdim=True)
    dln_w = (dgamma_o + dgamma_f).reshape(H, Fd)
    dln_b = (dbeta_o + dbeta_f).reshape(H, Fd)
    return dln_w, dln_b, dW1, db1, dW2, db2, dXQ, dXV, dXK, deta


def ttt_mlp_primal_backward(XQ, XK, XV, last_eta, ln_w, ln_b, W1, b1, W2, b2, dOut, dW_last=None):
    """Whole-sequence analytic backward (no checkpointing: keeps every state; small cases only).

    ``dW_last`` = optional upstream gradient (dW1, db1, dW2, db2) of the FINAL state (zero at the reference's op boundary,
    mlp_tk.py:179-182; non-zero when this range is followed by another one, see ttt_mlp_primal_backward_chunked and the
    sequence-sharded chain).
    Returns dict with dln_w, dln_b, dW1, db1, dW2, db2 (w.r.t. the *initial* state), dXQ, dXV, dXK, dlast_eta.
    """
    B, H, NC, CS, Fd = XQ.shape
    states = [(W1, b1, W2, b2)]
    saved = []
    for n in range(NC):
        st, _, s = ttt_mlp_step_primal(*states[-1], XQ[:, :, n], XK[:, :, n], XV[:, :, n], last_eta[:, :, n], ln_w, ln_b)
        states.append(st)
        saved.append(s)
    if dW_last is None:
        dW1 = torch.zeros_like(W1); db1 = torch.zeros_like(b1); dW2 = torch.zeros_like(W2); db2 = torch.zeros_like(b2)
    else:
        dW1, db1, dW2, db2 = [t.to(W1.dtype).reshape(r.shape) for t, r in zip(dW_last, (W1, b1, W2, b2))]
    dXQ = torch.empty_like(XQ); dXK = torch.empty_like(XK); dXV = torch.empty_like(XV)
    deta = torch.empty_like(last_eta)
    dlw = torch.zeros(H, Fd, dtype=XQ.dtype); dlb = torch.zeros(H, Fd, dtype=XQ.dtype)
    for n in reversed(range(NC)):
        W1_, _, W2_, _ = states[n]
        W1n, _, W2n, _ = states[n + 1]
        r = ttt_mlp_step_backward(XQ[:, :, n], XK[:, :, n], W1_, W2_, W1n, W2n, last_eta[:, :, n], ln_w, ln_b,
                                  saved[n], dW1, db1, dW2, db2, dOut[:, :, n])
        dlw += r[0]; dlb += r[1]
        dW1, db1, dW2, db2 = r[2:6]
        dXQ[:, :, n], dXV[:, :, n], dXK[:, :, n], deta[:, :, n] = r[6:10]
    return dict(dln_w=dlw, dln_b=dlb, dW1=dW1, db1=db1, dW2=dW2, db2=db2, dXQ=dXQ, dXV=dXV, dXK=dXK, dlast_eta=deta)


def ttt_mlp_primal_backward_chunked(XQ, XK, XV, last_eta, ln_w, ln_b, W1, b1, W2, b2, dOut, group, dW_last=None):
    """Same gradients as ttt_mlp_primal_backward with the memory of ONE checkpoint group: a forward pass keeps only the
    state entering every ``group`` mini-batches (what the native forward checkpoints, mlp_tk.py:95-98), then the groups are
    replayed last to first, the state gradient handed from group to group -- the reference backward kernel's structure
    (ttt-tk/kernels/ttt_backward/ttt.cu:453-470,824).  Lets the full 48-head, 282-step case run on a host in seconds."""
    NC = XQ.shape[2]
    _, ck, _ = ttt_mlp_primal_forward(XQ, XK, XV, last_eta, ln_w, ln_b, W1, b1, W2, b2, group)
    K = ck[0].shape[2]
    outs = {k: torch.empty_like(t) for k, t in (("dXQ", XQ), ("dXK", XK), ("dXV", XV), ("dlast_eta", last_eta))}
    dlw = torch.zeros_like(ln_w.reshape(XQ.shape[1], -1)); dlb = torch.zeros_like(dlw)
    carry = dW_last
    for k in reversed(range(K)):
        s, e = k * group, min(NC, (k + 1) * group)
        st = tuple(c[:, :, k] for c in ck)
        r = ttt_mlp_primal_backward(XQ[:, :, s:e], XK[:, :, s:e], XV[:, :, s:e], last_eta[:, :, s:e], ln_w, ln_b, *st,
                                    dOut[:, :, s:e], dW_last=carry)
        for name in outs:
            outs[name][:, :, s:e] = r[name]
        dlw += r["dln_w"]; dlb += r["dln_b"]
        carry = (r["dW1"], r["db1"], r["dW2"], r["db2"])
    return dict(dln_w=dlw, dln_b=dlb, dW1=carry[0], db1=carry[1], dW2=carry[2], db2=carry[3], **outs)


# --------------------------------------------------------------------------------------
# TTT-Linear  (ttt/models/ssm/ops/ttt_linear.py)
# --------------------------------------------------------------------------------------
def ttt_linear_step_dual(p, XQ, XK, XV, eta):
    """ops/ttt_linear.py:8-56."""
    W1, b1 = p["W1"], p["b1"]
    H, Fd = XQ.size(1), XQ.size(-1)
    ln_w = p["ln_w"].reshape(H, 1, Fd)
    ln_b = p["ln_b"].reshape(H, 1, Fd)
    Z1 = XK @ W1 + b1                                              # :28
    target = XV - XK
    gZ1 = ln_fused_l2_bwd(Z1, target, ln_w, ln_b)                  # :33
    Attn1 = XQ @ XK.transpose(-2, -1)                              # :35
    b1_bar = b1 - eta @ gZ1                                        # :36
    Z1_bar = XQ @ W1 - (eta * Attn1) @ gZ1 + b1_bar                # :37
    last_eta = eta[:, :, -1, :, None]
    W1n = W1 - (last_eta * XK).transpose(-1, -2) @ gZ1             # :40
    b1n = b1 - torch.sum(last_eta * gZ1, dim=-2, keepdim=True)     # :41
    out = XQ + ln_fwd(Z1_bar, ln_w, ln_b)                          # :43-45
    return dict(W1=W1n, b1=b1n, ln_w=p["ln_w"], ln_b=p["ln_b"]), out


def ttt_linear_eager(XK, XQ, XV, eta, ln_w, ln_b, W1, b1, checkpoint_group_size=0):
    """ops/ttt_linear.py:57-84 (K,Q,V order).  Returns ([B,NC,CS,H,F], final params)."""
    p = dict(W1=W1, b1=b1, ln_w=ln_w, ln_b=ln_b)
    outs = []
    for n in range(XK.shape[2]):
        p, o = ttt_linear_step_dual(p, XQ[:, :, n], XK[:, :, n], XV[:, :, n], eta[:, :, n])
        outs.append(o)
    return torch.stack(outs, 0).permute(1, 0, 3, 2, 4), p


def ttt_linear_primal_forward(XQ, XK, XV, last_eta, ln_w, ln_b, W1, b1):
    """Primal form of TTT-Linear: Z1_bar = XQ @ W1_next + b1_next (kernels/linear_forward.py:128-134)."""
    B, H, NC, CS, Fd = XQ.shape
    g = ln_w.reshape(H, 1, Fd); bt = ln_b.reshape(H, 1, Fd)
    out = torch.empty_like(XQ)
    for n in range(NC):
        q, k, v, e = XQ[:, :, n], XK[:, :, n], XV[:, :, n], last_eta[:, :, n]
        Z1 = k @ W1 + b1
        gZ1 = ln_fused_l2_bwd(Z1, v - k, g, bt)
        W1 = W1 - (e * k).transpose(-1, -2) @ gZ1
        b1 = b1 - (e * gZ1).sum(-2, keepdim=True)
        out[:, :, n] = q + ln_fwd(q @ W1 + b1, g, bt)
    return out, (W1, b1)


def ttt_linear_step_backward(XQ, XK, XV, W1, b1, W1n, b1n, last_eta, ln_w, ln_b, dW1n, db1n, dO):
    """Closed-form backward of one primal TTT-Linear step (the algebra of kernels/linear_backward.py:73-197; the MLP
    step above with the second layer dropped).  W1/b1 = state before the step, W1n/b1n = after; dW1n/db1n = gradient
    w.r.t. the state after the step.  Returns (dln_w, dln_b, dW1, db1, dXQ, dXV, dXK, dlast_eta[B,H,CS,1])."""
    H, Fd = XQ.shape[1], XQ.shape[-1]
    g = ln_w.reshape(H, 1, Fd); bt = ln_b.reshape(H, 1, Fd)
    eta = last_eta
    # recompute the forward intermediates
    Z1 = XK @ W1 + b1
    mu = Z1.mean(-1, keepdim=True); std_f = torch.sqrt(Z1.var(-1, keepdim=True, unbiased=False) + LN_EPS)
    xhat_f = (Z1 - mu) / std_f
    go = g * xhat_f + bt - (XV - XK)
    gxh = go * g
    gZ1 = (1.0 / Fd) * (Fd * gxh - gxh.sum(-1, keepdim=True) - xhat_f * (gxh * xhat_f).sum(-1, keepdim=True)) / std_f
    Z1b = XQ @ W1n + b1n
    mu_o = Z1b.mean(-1, keepdim=True); std_o = torch.sqrt(Z1b.var(-1, keepdim=True, unbiased=False) + LN_EPS)
    xhat_o = (Z1b - mu_o) / std_o
    # output LayerNorm + Q side
    dbeta_o = dO.sum(-2, keepdim=True).sum(0)
    dgamma_o = (dO * xhat_o).sum(-2, keepdim=True).sum(0)
    dxh = dO * g
    dZ1b = (1.0 / Fd) * (Fd * dxh - dxh.sum(-1, keepdim=True) - xhat_o * (dxh * xhat_o).sum(-1, keepdim=True)) / std_o
    db1n = db1n + dZ1b.sum(-2, keepdim=True)
    dW1n = dW1n + XQ.transpose(-2, -1) @ dZ1b
    dXQ = dO + dZ1b @ W1n.transpose(-2, -1)
    # update rule  W1n = W1 - (eta XK)^T gZ1,  b1n = b1 - sum(eta gZ1)
    dgZ1 = -(eta * XK) @ dW1n - eta * db1n
    T1 = gZ1 @ dW1n.transpose(-2, -1)
    dXK_u = -T1 * eta
    deta = -(T1 * XK).sum(-1, keepdim=True) - (db1n * gZ1).sum(-1, keepdim=True)
    # backward through gZ1 = ln_fused_l2_bwd(Z1, XV - XK)
    dgxh = (1.0 / std_f) * (dgZ1 + (-1.0 / Fd) * (dgZ1.sum(-1, keepdim=True) + xhat_f * (dgZ1 * xhat_f).sum(-1, keepdim=True)))
    dy = g * dgxh
    dgamma_f = (go * dgxh + dy * xhat_f).sum(-2, keepdim=True).sum(0)
    dbeta_f = dy.sum(-2, keepdim=True).sum(0)
    dxh_f = dy * g + (-1.0 / (Fd * std_f)) * (gxh * (dgZ1 * xhat_f).sum(-1, keepdim=True)
                                             + dgZ1 * (gxh * xhat_f).sum(-1, keepdim=True))
    dstd = (-dxh_f * xhat_f - dgZ1 * gZ1) / std_f
    dZ1 = dxh_f / std_f + (1.0 / Fd) * (dstd.sum(-1, keepdim=True) * xhat_f - dxh_f.sum(-1, keepdim=True) / std_f)
    dtarget = -dy
    dXK = -dtarget + dXK_u + dZ1 @ W1.transpose(-2, -1)
    dXV = dtarget
    dW1 = dW1n + XK.transpose(-2, -1) @ dZ1
    db1 = db1n + dZ1.sum(-2, keepdim=True)
    return ((dgamma_o + dgamma_f).reshape(H, Fd), (dbeta_o + dbeta_f).reshape(H, Fd), dW1, db1, dXQ, dXV, dXK, deta)


def ttt_linear_primal_backward(XQ, XK, XV, last_eta, ln_w, ln_b, W1, b1, dOut):
    """Whole-sequence analytic backward of TTT-Linear (keeps every state; small cases only).  last_eta [B,H,NC,CS,1]."""
    B, H, NC, CS, Fd = XQ.shape
    g = ln_w.reshape(H, 1, Fd); bt = ln_b.reshape(H, 1, Fd)
    states = [(W1, b1)]
    for n in range(NC):
        W, b = states[-1]
        k, v, e = XK[:, :, n], XV[:, :, n], last_eta[:, :, n]
        gZ1 = ln_fused_l2_bwd(k @ W + b, v - k, g, bt)
        states.append((W - (e * k).transpose(-1, -2) @ gZ1, b - (e * gZ1).sum(-2, keepdim=True)))
    dW1 = torch.zeros_like(W1); db1 = torch.zeros_like(b1)
    dXQ = torch.empty_like(XQ); dXK = torch.empty_like(XK); dXV = torch.empty_like(XV)
    deta = torch.empty_like(last_eta)
    dlw = torch.zeros(H, Fd, dtype=XQ.dtype); dlb = torch.zeros(H, Fd, dtype=XQ.dtype)
    for n in reversed(range(NC)):
        r = ttt_linear_step_backward(XQ[:, :, n], XK[:, :, n], XV[:, :, n], *states[n], *states[n + 1], last_eta[:, :, n],
                                     ln_w, ln_b, dW1, db1, dOut[:, :, n])
        dlw += r[0]; dlb += r[1]
        dW1, db1 = r[2], r[3]
        dXQ[:, :, n], dXV[:, :, n], dXK[:, :, n], deta[:, :, n] = r[4:8]
    return dict(dln_w=dlw, dln_b=dlb, dW1=dW1, db1=db1, dXQ=dXQ, dXV=dXV, dXK=dXK, dlast_eta=deta)


# --------------------------------------------------------------------------------------
# Autograd-of-eager gradient oracle (what TkMLP.backward / TritonLinear.backward must match)
# --------------------------------------------------------------------------------------
def ttt_mlp_eager_grads(XQ, XK, XV, eta, ln_w, ln_b, W1, b1, W2, b2, dOut):
    """Autograd through the eager dual form.  dOut is in the op layout [B,H,NC,CS,F].

    Returns grads in TkMLP.backward order (mlp_tk.py:282-294) with deta as the full [B,H,NC,CS,CS] tensor.
    """
    ins = [t.detach().clone().requires_grad_(True) for t in (ln_w, ln_b, W1, b1, W2, b2, XQ, XV, XK, eta)]
    lw, lb, w1, bb1, w2, bb2, q, v, k, e = ins
    out, _ = ttt_mlp_eager(k, q, v, e, lw, lb, w1, bb1, w2, bb2)
    out = out.permute(0, 3, 1, 2, 4)  # [B,NC,CS,H,F] -> [B,H,NC,CS,F]   (ttt_layer.py:456 inverse)
    out.backward(dOut)
    return [t.grad for t in ins], out.detach()


def ttt_linear_eager_grads(XQ, XK, XV, eta, ln_w, ln_b, W1, b1, dOut):
    ins = [t.detach().clone().requires_grad_(True) for t in (ln_w, ln_b, W1, b1, XQ, XV, XK, eta)]
    lw, lb, w1, bb1, q, v, k, e = ins
    out, _ = ttt_linear_eager(k, q, v, e, lw, lb, w1, bb1)
    out = out.permute(0, 3, 1, 2, 4)
    out.backward(dOut)
    return [t.grad for t in ins], out.detach()


# --------------------------------------------------------------------------------------
# Bidirectional gated TTT around an arbitrary sequence op  (ttt/models/cogvideo/dit.py:213-266)
# --------------------------------------------------------------------------------------
def reverse_text_chunks(text_emb, num_chunks):
    """dit.py:213-217."""
    B, L, E = text_emb.shape
    return torch.flip(text_emb.reshape(B, num_chunks, L // num_chunks, E), dims=[1]).reshape(B, L, E)


def gate(alpha_text, alpha_video, residual, ssm_out, text_length):
    """dit.py:219-222 with SSMGating (dit.py:90-103): residual + tanh(alpha) * ssm_out, separate text/video alpha."""
    g = torch.cat([torch.tanh(alpha_text) * ssm_out[:, :text_length], torch.tanh(alpha_video) * ssm_out[:, text_length:]], dim=1)
    return residual + g


def ssm_bidirectional(emb, ssm, text_length, num_chunks, is_multiscene, a_ft, a_fv, a_bt, a_bv):
    """dit.py:224-266.  ``ssm`` maps [B,L,E] -> [B,L,E]; text tokens come first (seq_text_length = text_length)."""
    res = emb.clone()
    emb = ssm(emb)
    emb = gate(a_ft, a_fv, res, emb, text_length)
    res = emb.clone()
    emb = emb.clone()
    if is_multiscene:
        emb[:, :text_length] = reverse_text_chunks(emb[:, :text_length], num_chunks)
    emb[:, text_length:] = torch.flip(res[:, text_length:], dims=[1])
    emb = ssm(emb)
    emb = emb.clone()
    if is_multiscene:
        emb[:, :text_length] = reverse_text_chunks(emb[:, :text_length], num_chunks)
    emb[:, text_length:] = torch.flip(emb[:, text_length:], dims=[1])
    return gate(a_bt, a_bv, res, emb, text_length)


# --------------------------------------------------------------------------------------
# Local per-segment attention core  (ttt/models/cogvideo/dit.py:163-211)
# --------------------------------------------------------------------------------------
def sdpa_math(q, k, v):
    """Non-causal softmax attention, math form, in the dtype of q (fp32/fp64 on CPU).  dit.py:196-198."""
    s = (q @ k.transpose(-2, -1)) / math.sqrt(q.shape[-1])
    return torch.softmax(s, dim=-1) @ v


# --------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------
def make_inputs(B, H, NC, CS=64, Fd=64, expansion=4, seed=0, dtype=torch.float32, base_lr=0.1, linear=False):
    """XQ,XK L2-normalised randn; XV randn; row-uniform eta = (base_lr/F)*sigmoid(randn)/CS;
    gamma = 1+0.1 randn, beta = 0.1 randn; W = 0.02 randn, b = 0 (ttt_layer.py:404-416)."""
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float32)
    XQ = F.normalize(r(B, H, NC, CS, Fd), dim=-1)
    XK = F.normalize(r(B, H, NC, CS, Fd), dim=-1)
    XV = r(B, H, NC, CS, Fd)
    lr = (base_lr / Fd) * torch.sigmoid(r(B, H, NC, 1, CS)) / CS
    eta = lr.repeat(1, 1, 1, CS, 1)
    ln_w = 1.0 + 0.1 * r(H, Fd)
    ln_b = 0.1 * r(H, Fd)
    hid = Fd if linear else expansion * Fd
    W1 = (0.02 * r(H, Fd, hid)).unsqueeze(0).repeat(B, 1, 1, 1)
    b1 = torch.zeros(B, H, 1, hid)
    d = dict(XQ=XQ, XK=XK, XV=XV, eta=eta, ln_w=ln_w, ln_b=ln_b, W1=W1, b1=b1)
    if not linear:
        d["W2"] = (0.02 * r(H, hid, Fd)).unsqueeze(0).repeat(B, 1, 1, 1)
        d["b2"] = torch.zeros(B, H, 1, Fd)
    d["dOut"] = r(B, H, NC, CS, Fd)
    return {k: v.to(dtype) for k, v in d.items()}


def rel_err(a, b):
    """||a-b|| / ||b||  -- the tolerance metric of SURVEY 8c (<= 1e-2 vs the fp32 eager path)."""
    a = a.double(); b = b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


# --------------------------------------------------------------------------------------
# Local attention block restatement (q/k/v/o Linear + per-head LN + 3-D RoPE + SDPA + overlap average)
#   ttt/models/cogvideo/dit.py:163-211, cogvideo/utils.py:93-99 (rotate_half), :363-437 (Rotary3DPositionEmbedding)
# --------------------------------------------------------------------------------------
def rope3d_tables(height, width, num_frames, head_dim, theta=10000.0):
    """cogvideo/utils.py:388-425: per-position sin/cos tables [(t h w), head_dim], pairs repeated (n r)."""
    dim_t = head_dim // 4
    dim_h = head_dim // 8 * 3
    dim_w = head_dim // 8 * 3
    f_t = 1.0 / (theta ** (torch.arange(0, dim_t, 2)[: dim_t // 2].float() / dim_t))
    f_h = 1.0 / (theta ** (torch.arange(0, dim_h, 2)[: dim_h // 2].float() / dim_h))
    f_w = 1.0 / (theta ** (torch.arange(0, dim_w, 2)[: dim_w // 2].float() / dim_w))
    g_t = torch.arange(num_frames, dtype=torch.float32)[:, None] * f_t[None]
    g_h = torch.arange(height, dtype=torch.float32)[:, None] * f_h[None]
    g_w = torch.arange(width, dtype=torch.float32)[:, None] * f_w[None]
    g_t, g_h, g_w = (x.repeat_interleave(2, dim=-1) for x in (g_t, g_h, g_w))
    T, Hh, Ww = num_frames, height, width
    fr = torch.cat([
        g_t[:, None, None, :].expand(T, Hh, Ww, -1),
        g_h[None, :, None, :].expand(T, Hh, Ww, -1),
        g_w[None, None, :, :].expand(T, Hh, Ww, -1)], dim=-1).reshape(T * Hh * Ww, -1)
    return fr.sin(), fr.cos()


def rotate_half(x):
    """cogvideo/utils.py:93-99: interleaved pairs (x1,x2) -> (-x2,x1)."""
    x = x.reshape(*x.shape[:-1], -1, 2)
    x1, x2 = x.unbind(-1)
    return torch.stack((-x2, x1), dim=-1).flatten(-2)


def apply_rope(t, sin, cos):
    """cogvideo/utils.py:432-437: segment-local positions freqs[:seq_len]."""
    L = t.shape[2]
    return t * cos[:L].to(t.dtype) + rotate_half(t) * sin[:L].to(t.dtype)


def local_attention_block(vid, text, P, num_heads, text_length, tokens_per_frame, num_chunks,
                          attn_length, prefix_len, sin, cos, ln_eps=1e-6):
    """dit.py:163-211.  P: dict of q/k/v/o weight+bias and q_norm/k_norm weight+bias."""
    B, _, E = vid.shape
    D = E // num_heads
    out_vid = torch.zeros_like(vid)
    out_txt = torch.zeros_like(text)
    cnt = torch.zeros_like(vid[..., :1])
    for i in range(num_chunks):
        s = i * attn_length * tokens_per_frame
        e = (prefix_len + (i + 1) * attn_length) * tokens_per_frame
        ts, te = i * text_length, (i + 1) * text_length
        cur = torch.cat([text[:, ts:te], vid[:, s:e]], dim=1)
        heads = lambda x: x.reshape(B, -1, num_heads, D).transpose(1, 2)
        q = heads(F.linear(cur, P["q.weight"], P["q.bias"]))
        k = heads(F.linear(cur, P["k.weight"], P["k.bias"]))
        v = heads(F.linear(cur, P["v.weight"], P["v.bias"]))
        q = F.layer_norm(q, (D,), P["q_norm.weight"], P["q_norm.bias"], ln_eps)
        k = F.layer_norm(k, (D,), P["k_norm.weight"], P["k_norm.bias"], ln_eps)
        q = torch.cat([q[:, :, :text_length], apply_rope(q[:, :, text_length:], sin, cos)], dim=2)
        k = torch.cat([k[:, :, :text_length], apply_rope(k[:, :, text_length:], sin, cos)], dim=2)
        a = sdpa_math(q, k, v).transpose(1, 2).reshape(B, -1, E)
        a = F.linear(a, P["o.weight"], P["o.bias"])
        out_txt[:, ts:te] = a[:, :text_length]
        out_vid[:, s:e] += a[:, text_length:]
        cnt[:, s:e] += 1
    return torch.cat((out_txt, out_vid / cnt), dim=1)


# --------------------------------------------------------------------------------------
# TTTBase.process_input (ttt/models/ssm/ttt_layer.py:252-306) without the q/k/v Linears: the op's input preparation
# --------------------------------------------------------------------------------------
def ttt_rope_tables(height, width, num_frames, head_dim, theta=10000.0):
    """ssm/utils.py:9-53 (precompute_freqs_cis_3d): angle per (video position, pair) -> (cos, sin) [(t h w), head_dim/2]."""
    dim_t = head_dim // 4
    dim_h = head_dim // 8 * 3
    dim_w = head_dim // 8 * 3
    f_t = 1.0 / (theta ** (torch.arange(0, dim_t, 2)[: dim_t // 2].float() / dim_t))
    f_h = 1.0 / (theta ** (torch.arange(0, dim_h, 2)[: dim_h // 2].float() / dim_h))
    f_w = 1.0 / (theta ** (torch.arange(0, dim_w, 2)[: dim_w // 2].float() / dim_w))
    g_t = torch.arange(num_frames, dtype=torch.float32)[:, None] * f_t[None]
    g_h = torch.arange(height, dtype=torch.float32)[:, None] * f_h[None]
    g_w = torch.arange(width, dtype=torch.float32)[:, None] * f_w[None]
    T, Hh, Ww = num_frames, height, width
    ang = torch.cat([
        g_t[:, None, None, :].expand(T, Hh, Ww, -1),
        g_h[None, :, None, :].expand(T, Hh, Ww, -1),
        g_w[None, None, :, :].expand(T, Hh, Ww, -1)], dim=-1).reshape(T * Hh * Ww, -1)
    return ang.cos(), ang.sin()


def interleave_index(L, text_length, num_chunks, init_offset):
    """ttt_layer.py:157-189 (interleave) as a gather index: out[:, :, l] = x[:, :, idx[l]] over the flattened token axis."""
    seq_text = text_length * num_chunks
    text = torch.arange(seq_text).chunk(num_chunks)
    video = torch.arange(seq_text, L)
    v0 = init_offset - text_length
    vids = (video[:v0],) + tuple(video[v0:].chunk(num_chunks - 1))
    return torch.cat([torch.cat((text[i], vids[i])) for i in range(num_chunks)])


def ttt_process_input(XQ, XK, XV, lr_logit, cos, sin, ln_w, ln_b, seq_text_length, base_lr, CS, index=None):
    """ttt_layer.py:252-306 after the Linears.  XQ/XK/XV [B,L,H,F] (Linear outputs), lr_logit [B,L,H] (= X.w_h + b_h,
    ttt_layer.py:148-151), cos/sin [Lv, F/2].  Returns XQ, XK, XV [B,H,NC,CS,F] and eta [B,H,NC,CS,CS] exactly as the
    reference builds them (eta rows are repeated BEFORE the interleave, so interleaved rows mix mini-batches)."""
    B, L, H, Fd = XQ.shape
    q = F.normalize(XQ, p=2, dim=-1)  # :265-266
    k = F.normalize(XK, p=2, dim=-1)

    def rope(x):  # ssm/utils.py:82-108: complex multiply on interleaved pairs, video tokens only
        xv = x[:, seq_text_length:].reshape(B, L - seq_text_length, H, Fd // 2, 2)
        c, s_ = cos[: L - seq_text_length, None, :].to(x.dtype), sin[: L - seq_text_length, None, :].to(x.dtype)
        re = xv[..., 0] * c - xv[..., 1] * s_
        im = xv[..., 0] * s_ + xv[..., 1] * c
        return torch.cat((x[:, :seq_text_length], torch.stack((re, im), dim=-1).flatten(-2)), dim=1)

    q, k = rope(q), rope(k)
    d = XV - k  # ln_reconstruction_target :219-235 (unbiased std, eps added to std)
    d = (d - d.mean(-1, keepdim=True)) / (d.std(-1, keepdim=True) + 1e-8)
    v = ln_w[None, None] * d + ln_b[None, None] + k
    NC = L // CS
    to_mb = lambda t: t.transpose(1, 2).reshape(B, H, NC, CS, Fd)  # :237-250
    q, k, v = to_mb(q), to_mb(k), to_mb(v)
    lr = base_lr * torch.sigmoid(lr_logit) / Fd  # :143-155
    lr = lr.transpose(1, 2).reshape(B, H, NC, 1, CS)
    eta = (1.0 / CS) * lr.repeat(1, 1, 1, CS, 1)  # :287-288
    if index is not None:  # :290-294
        g = lambda t: t.reshape(B, H, NC * CS, -1)[:, :, index].reshape(t.shape)
        q, k, v, eta = g(q), g(k), g(v), g(eta)
    return q, k, v, eta


def undo_interleave_index(L, text_length, num_chunks, init_offset, base_offset):
    """ttt_layer.py:191-217 (undo_interleave) as a gather index: out[:, m] = x[:, idx[m]]."""
    text, vid = [], []
    for i in range(num_chunks):
        s0 = 0 if i == 0 else init_offset + (i - 1) * base_offset
        e0 = init_offset if i == 0 else init_offset + i * base_offset
        text.append(torch.arange(s0, s0 + text_length))
        vid.append(torch.arange(s0 + text_length, e0))
    return torch.cat(text + vid)


def ttt_output_epilogue(O_bhncf, ln_w, ln_b, eps=1e-6, index=None):
    """TTTMLP.ttt tail + TTTBase.forward up to (not including) wo: permute(0,2,3,1,4).reshape(B, L, E)
    (ttt_layer.py:456,472), post_norm = LayerNorm(E, eps 1e-6) (:71,324); undo_interleave (:329-331) is a token permutation
    and commutes with the per-token wo Linear, so it is applied here, before wo."""
    B, H, NC, CS, Fd = O_bhncf.shape
    x = O_bhncf.permute(0, 2, 3, 1, 4).reshape(B, NC * CS, H * Fd)
    x = F.layer_norm(x, (H * Fd,), ln_w, ln_b, eps)
    return x if index is None else x[:, index]

