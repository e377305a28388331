"""Small, fixed workload for ncu captures (one call of every kernel family; never a bench number)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import __graft_entry__

__graft_entry__.build()
from oracle import ttt_oracle as O  # input generator only
from ttt_video_dit_b200 import attention, linear_triton, mlp_tk

which = sys.argv[1:] or ["mlp", "linear", "attention"]
dev = "cuda"
bf = lambda t: t.to(torch.bfloat16).to(dev)
if "mlp" in which:  # 3 checkpoint groups of 16 steps, 48 heads
    d = O.make_inputs(1, 48, 48, seed=1)
    prm = [d[k].to(dev).requires_grad_(True) for k in ("ln_w", "ln_b", "W1", "b1", "W2", "b2")]
    q, v, k = [bf(d[n]).requires_grad_(True) for n in ("XQ", "XV", "XK")]
    e = bf(d["eta"])[:, :, :, -1, :].clone().requires_grad_(True)
    for _ in range(2):
        out = mlp_tk.ttt_mlp_op(*prm, q, v, k, e, 16)
        out.backward(bf(d["dOut"]))
    torch.cuda.synchronize()
if "linear" in which:
    d = O.make_inputs(1, 48, 256, CS=16, seed=1, base_lr=1.0, linear=True)
    prm = [d[k].to(dev).requires_grad_(True) for k in ("ln_w", "ln_b", "W1", "b1")]
    q, v, k, e = [bf(d[n]).requires_grad_(True) for n in ("XQ", "XV", "XK", "eta")]
    for _ in range(2):
        out = linear_triton.TritonLinear.apply(*prm, q, v, k, e, 16)
        out.backward(bf(d["dOut"]))
    torch.cuda.synchronize()
if "attention" in which:
    q, k, v, go = (torch.randn(1, 4096, 48, 64, device=dev).to(torch.bfloat16) for _ in range(4))
    q, k, v = (t.requires_grad_(True) for t in (q, k, v))
    for _ in range(2):
        out = attention.sdpa_bthd(q, k, v)
        out.backward(go)
    torch.cuda.synchronize()
print("ncu target done", which)
