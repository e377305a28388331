"""Secondary measurements (not the driver's bench contract): TTT-Linear forward, local attention forward vs the library
SDPA the reference calls, gate kernels vs the HBM roofline.  One JSON line per kernel.  CUDA-event timing, 3 warm-ups."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

import __graft_entry__

__graft_entry__.build()
from ttt_video_dit_b200 import attention, linear_triton, process_input, seq_block, ttt_layer

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
dev = "cuda"


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_attention(T=18048, H=48, B=1):
    q, k, v = (torch.randn(B, T, H, 64, device=dev, dtype=torch.bfloat16) for _ in range(3))
    ms = timeit(lambda: attention.sdpa_bthd(q, k, v))
    qh, kh, vh = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    ms_lib = timeit(lambda: F.scaled_dot_product_attention(qh, kh, vh, is_causal=False))
    flop = 4.0 * T * T * 64 * H * B
    print(json.dumps({"kernel": "attn_fwd_kernel", "shape": [B, T, H, 64], "ms": ms, "tflops": flop / ms / 1e9,
                      "frac_of_bf16_peak": flop / ms / 1e9 / peaks["bf16_tflops"],
                      "library_sdpa_ms": ms_lib, "library_sdpa_tflops": flop / ms_lib / 1e9, "speed_vs_library": ms_lib / ms}))


def bench_linear(NC=1128, H=48, B=1):
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g)
    q = F.normalize(r(B, H, NC, 16, 64), dim=-1).to(torch.bfloat16).to(dev)
    k = F.normalize(r(B, H, NC, 16, 64), dim=-1).to(torch.bfloat16).to(dev)
    v = r(B, H, NC, 16, 64).to(torch.bfloat16).to(dev)
    le = ((1.0 / 64) * torch.sigmoid(r(B, H, NC, 16)) / 16).to(torch.bfloat16).to(dev)
    lw, lb = (1 + 0.1 * r(H, 64)).to(dev), (0.1 * r(H, 64)).to(dev)
    W1 = (0.02 * r(B, H, 64, 64)).to(dev); b1 = torch.zeros(B, H, 1, 64, device=dev)
    ms = timeit(lambda: linear_triton.linear_forward(q, k, v, le, lw, lb, W1, b1, 16))
    flop = 3 * 2 * 16 * 64 * 64 * NC * H * B
    print(json.dumps({"kernel": "ttt_linear_fwd_kernel", "shape": [B, H, NC, 16, 64], "ms": ms, "tokens_per_s": B * NC * 16 / ms * 1e3,
                      "tflops": flop / ms / 1e9, "us_per_minibatch": ms * 1e3 / NC}))
    # forward + backward through the TritonLinear mirror (trajectory recompute + reverse scan), checkpoint group 16
    eta = le.float()[:, :, :, None, :].expand(B, H, NC, 16, 16).to(torch.bfloat16).contiguous()
    prm = [t.clone().requires_grad_(True) for t in (lw, lb, W1, b1)]
    qq, vv, kk, ee = [t.clone().requires_grad_(True) for t in (q, v, k, eta)]
    go = r(B, H, NC, 16, 64).to(torch.bfloat16).to(dev)

    def step():
        out = linear_triton.TritonLinear.apply(*prm, qq, vv, kk, ee, 16)
        out.backward(go)

    ms2 = timeit(step)
    flop2 = 9 * 2 * 16 * 64 * 64 * NC * H * B  # fwd 3 U_L + bwd 6 U_L; recompute not counted
    print(json.dumps({"kernel": "ttt_linear fwd+bwd (TritonLinear.apply + backward)", "shape": [B, H, NC, 16, 64], "ms": ms2,
                      "tokens_per_s": B * NC * 16 / ms2 * 1e3, "tflops": flop2 / ms2 / 1e9, "us_per_minibatch": ms2 * 1e3 / NC}))


def bench_gate(L=18048, E=3072, B=1):
    x = torch.randn(B, L, E, device=dev, dtype=torch.bfloat16)
    s = torch.randn(B, L, E, device=dev, dtype=torch.bfloat16)
    a = [0.1 * torch.ones(E, device=dev) for _ in range(2)]
    ms = timeit(lambda: seq_block.GatedResidual.apply(x, s, a[0], a[1], 498, 1, False, True))
    bytes_ = 4 * x.numel() * 2  # read res, s; write out, rev
    print(json.dumps({"kernel": "gate_fwd_kernel<false,true>", "shape": [B, L, E], "ms": ms, "GBps": bytes_ / ms / 1e6,
                      "frac_of_hbm_peak": bytes_ / ms / 1e6 / peaks["hbm_gbs"]}))


def bench_attention_bwd(T=18048, H=48, B=1):
    q, k, v, go = (torch.randn(B, T, H, 64, device=dev).to(torch.bfloat16) for _ in range(4))
    q, k, v = (t.requires_grad_(True) for t in (q, k, v))

    def step():
        attention.sdpa_bthd(q, k, v).backward(go)

    ms = timeit(step, iters=5)
    qh, kh, vh = (t.detach().permute(0, 2, 1, 3).requires_grad_(True) for t in (q, k, v))
    goh = go.permute(0, 2, 1, 3)
    ms_lib = timeit(lambda: F.scaled_dot_product_attention(qh, kh, vh, is_causal=False).backward(goh), iters=5)
    flop = 3.5 * 4.0 * T * T * 64 * H * B  # forward 1x + backward 2.5x
    print(json.dumps({"kernel": "attn_fwd_kernel + attn_bwd_kernel<0>,<1> (sdpa_bthd fwd+bwd)", "shape": [B, T, H, 64], "ms": ms,
                      "tflops": flop / ms / 1e9, "frac_of_bf16_peak": flop / ms / 1e9 / peaks["bf16_tflops"],
                      "library_sdpa_ms": ms_lib, "speed_vs_library": ms_lib / ms}))


def bench_process_input(L=18048, H=48, B=1):
    g = torch.Generator().manual_seed(0)
    xq, xk, xv = (torch.randn(B, L, H * 64, generator=g).to(torch.bfloat16).to(dev) for _ in range(3))
    logit = torch.randn(B, L, H, generator=g).to(dev)
    cos, sin = torch.rand(L, 32, generator=g).to(dev), torch.rand(L, 32, generator=g).to(dev)
    lw, lb = torch.ones(H, 64, device=dev), torch.zeros(H, 64, device=dev)
    ms = timeit(lambda: process_input.prepare(xq, xk, xv, logit, cos, sin, lw, lb, 498, 64, 0.1))
    bytes_ = 6 * xq.numel() * 2 + logit.numel() * 4 + B * H * L * 2  # read q,k,v + write XQ,XK,XV (+ logits, eta)
    print(json.dumps({"kernel": "ttt_process_input_kernel", "shape": [B, L, H * 64], "ms": ms, "GBps": bytes_ / ms / 1e6,
                      "frac_of_hbm_peak": bytes_ / ms / 1e6 / peaks["hbm_gbs"]}))


def bench_output_norm(L=18048, H=48, B=1):
    x = torch.randn(B, H, L // 64, 64, 64, device=dev).to(torch.bfloat16)
    w, b_ = torch.ones(H * 64, device=dev), torch.zeros(H * 64, device=dev)
    ms = timeit(lambda: process_input.output_norm(x, w, b_))
    bytes_ = 2 * x.numel() * 2
    print(json.dumps({"kernel": "ttt_output_norm_kernel", "shape": [B, L, H * 64], "ms": ms, "GBps": bytes_ / ms / 1e6,
                      "frac_of_hbm_peak": bytes_ / ms / 1e6 / peaks["hbm_gbs"]}))


def bench_block(frames=13, H=48, B=1, TL=498):
    """SeqModelingBlock level (dit.py:163-266) at CogVideoX-5B dims, 3-second video: local attention over the one segment
    + bidirectional gated TTT-MLP layer (forward and reversed direction), forward + backward, random bf16 weights.
    Everything between the library Linears runs in this repo's kernels."""
    E, tpf, CS = H * 64, 1350, 64
    L = TL + frames * tpf
    g = torch.Generator().manual_seed(0)
    rb = lambda *s_: (0.02 * torch.randn(*s_, generator=g)).to(torch.bfloat16).to(dev).requires_grad_(True)
    rf = lambda *s_: (0.02 * torch.randn(*s_, generator=g)).to(dev).requires_grad_(True)
    P = {}
    for n in ("wq", "wk", "wv", "wo"):
        P[n + ".weight"], P[n + ".bias"] = rb(E, E), rb(E)
    P["learnable_ttt_lr_weight"], P["learnable_ttt_lr_bias"] = rb(H, 1, E), rb(H, 1)
    P["ttt_norm_weight"] = torch.ones(H, 64, device=dev, requires_grad=True); P["ttt_norm_bias"] = torch.zeros(H, 64, device=dev, requires_grad=True)
    P["post_norm.weight"] = torch.ones(E, device=dev, requires_grad=True); P["post_norm.bias"] = torch.zeros(E, device=dev, requires_grad=True)
    P["W1"], P["b1"], P["W2"], P["b2"] = rf(H, 64, 256), torch.zeros(H, 1, 256, device=dev, requires_grad=True), rf(H, 256, 64), torch.zeros(H, 1, 64, device=dev, requires_grad=True)
    A = {}
    for n in ("q", "k", "v", "o"):
        A[n + ".weight"], A[n + ".bias"] = rb(E, E), rb(E)
    for n in ("q_norm", "k_norm"):
        A[n + ".weight"], A[n + ".bias"] = torch.ones(64, device=dev, dtype=torch.bfloat16, requires_grad=True), torch.zeros(64, device=dev, dtype=torch.bfloat16, requires_grad=True)
    gates = [torch.full((E,), 0.1, device=dev, requires_grad=True) for _ in range(4)]
    cos_t, sin_t = torch.rand(frames * tpf, 32, generator=g).to(dev), torch.rand(frames * tpf, 32, generator=g).to(dev)
    sin_a, cos_a = torch.rand(frames * tpf, 64, generator=g).to(dev), torch.rand(frames * tpf, 64, generator=g).to(dev)
    vid = torch.randn(B, frames * tpf, E, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
    txt = torch.randn(B, TL, E, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
    go = torch.randn(B, L, E, generator=g).to(torch.bfloat16).to(dev)
    layer = lambda x: ttt_layer.ttt_layer_forward(x, P, cos_t, sin_t, TL, CS, 0.1, 16, kind="ttt_mlp")

    def step():
        a = attention.local_attention(vid, txt, A, H, TL, tpf, 1, frames, 0, sin_a, cos_a)
        y = seq_block.ssm_forward(a.contiguous(), layer, TL, 1, False, *gates)
        y.backward(go)

    ms = timeit(step, iters=5, warm=2)
    print(json.dumps({"kernel": "SeqModelingBlock fwd+bwd (local attention + bidirectional gated TTT-MLP layer), 5B dims, 3 s",
                      "shape": [B, L, E], "ms": ms, "tokens_per_s": B * L / ms * 1e3}))


def bench_adaln(L=18048, E=3072, B=1):
    """adaLN shell kernels (csrc/adaln.cu) forward and backward vs the HBM roofline: ln_affine reads x and writes out (4 bytes
    per element), its backward reads x, g and writes gx (6); gate_add reads x, y, writes out (6), backward reads g, y, writes
    dy (6)."""
    from ttt_video_dit_b200 import transformer_layer as TLm
    x = torch.randn(B, L, E, device=dev).to(torch.bfloat16).requires_grad_(True)
    y = torch.randn(B, L, E, device=dev).to(torch.bfloat16).requires_grad_(True)
    A = (1 + 0.1 * torch.randn(B, 2, E, device=dev)).requires_grad_(True)
    C = (0.1 * torch.randn(B, 2, E, device=dev)).requires_grad_(True)
    go = torch.randn(B, L, E, device=dev).to(torch.bfloat16)
    n = x.numel()
    for name, fwd, fb, bb in (("ln_affine", lambda: TLm.LnAffine.apply(x, A, C, 498, 1e-6), 4, 6),
                              ("gate_add", lambda: TLm.GateAdd.apply(x, y, A, 498), 6, 6)):
        with torch.no_grad():
            ms_f = timeit(fwd)
        def fwdbwd():
            o = fwd()
            o.backward(go)
            x.grad = None; y.grad = None; A.grad = None; C.grad = None
        ms_fb = timeit(fwdbwd)
        ms_b = ms_fb - ms_f
        print(json.dumps({"kernel": name, "shape": [B, L, E], "fwd_ms": ms_f, "fwd_GBps": fb * n / ms_f / 1e6,
                          "fwd_frac_of_hbm_peak": fb * n / ms_f / 1e6 / peaks["hbm_gbs"], "bwd_ms": ms_b,
                          "bwd_GBps": bb * n / ms_b / 1e6, "bwd_frac_of_hbm_peak": bb * n / ms_b / 1e6 / peaks["hbm_gbs"]}))


def bench_prologue(T=18048, H=48, B=1):
    """Attention prologue (csrc/attn_prologue.cu): q/k LayerNorm + RoPE, 8 bytes per element forward (read + write of q and k)
    and 12 backward (x, dy in; dx out), vs the same math as eager torch ops (what local_attention ran before)."""
    q = torch.randn(B, T, H, 64, device=dev).to(torch.bfloat16).requires_grad_(True)
    k = torch.randn(B, T, H, 64, device=dev).to(torch.bfloat16).requires_grad_(True)
    w = [torch.ones(64, device=dev, requires_grad=True), torch.zeros(64, device=dev, requires_grad=True),
         torch.ones(64, device=dev, requires_grad=True), torch.zeros(64, device=dev, requires_grad=True)]
    sin, cos = torch.rand(T, 64, device=dev), torch.rand(T, 64, device=dev)
    go = torch.randn(B, T, H, 64, device=dev).to(torch.bfloat16)
    f = lambda: attention.QKNormRope.apply(q, k, *w, sin, cos, 498, 1e-6)
    with torch.no_grad():
        ms_f = timeit(f)
    def fb():
        a, b_ = f()
        torch.autograd.backward((a, b_), (go, go))
        q.grad = None; k.grad = None
    ms_fb = timeit(fb)
    n = 2 * q.numel()

    def eager():
        outs = []
        for x_, gw, gb in ((q, w[0], w[1]), (k, w[2], w[3])):
            yv = F.layer_norm(x_, (64,), gw.to(x_.dtype), gb.to(x_.dtype), 1e-6)
            c_, s_ = cos[:T - 498].to(yv.dtype)[None, :, None, :], sin[:T - 498].to(yv.dtype)[None, :, None, :]
            v_ = yv[:, 498:]
            r = v_.reshape(*v_.shape[:-1], 32, 2)
            r = torch.stack((-r[..., 1], r[..., 0]), dim=-1).flatten(-2)
            outs.append(torch.cat([yv[:, :498], v_ * c_ + r * s_], dim=1))
        return outs
    def eager_fb():
        a, b_ = eager()
        torch.autograd.backward((a, b_), (go, go))
        q.grad = None; k.grad = None
    ms_e = timeit(eager_fb)
    print(json.dumps({"kernel": "qk_norm_rope (attention prologue)", "shape": [B, T, H, 64], "fwd_ms": ms_f, "fwd_GBps": 4 * n / ms_f / 1e6,
                      "fwd_frac_of_hbm_peak": 4 * n / ms_f / 1e6 / peaks["hbm_gbs"], "fwd_bwd_ms": ms_fb,
                      "bwd_frac_of_hbm_peak": 6 * n / (ms_fb - ms_f) / 1e6 / peaks["hbm_gbs"], "eager_torch_fwd_bwd_ms": ms_e}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["attention", "attention_bwd", "linear", "gate", "process_input", "output_norm", "adaln", "prologue", "block"]
    if "adaln" in which:
        bench_adaln()
    if "prologue" in which:
        bench_prologue()
    if "block" in which:
        bench_block()
    if "output_norm" in which:
        bench_output_norm()
    if "attention_bwd" in which:
        bench_attention_bwd()
    if "process_input" in which:
        bench_process_input()
    if "attention" in which:
        bench_attention()
    if "linear" in which:
        bench_linear()
    if "gate" in which:
        bench_gate()
