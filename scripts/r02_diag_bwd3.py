"""Diagnostic (GPU): dump the inputs of the ill-conditioned heads of the H=48 / NC=282 case (first 48 steps), the oracle's
carried state gradient entering step 47 and the kernel's gradients, for offline analysis with a rounding model."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import ttt_oracle as O
import test_gpu_full_shape as T
torch.set_num_threads(32)
H, NC, G, CUT = 48, 282, 16, 48
heads = [30, 12, 0]
d = T._bench_like_inputs(1, H, NC, seed=3)
state = [d[n] for n in ("W1", "b1", "W2", "b2")]
whole = T._forward(d, slice(0, NC), state, G)
g = T._backward(whole, d, slice(0, NC), G)
torch.cuda.synchronize()
c = lambda t: t.double().cpu()
sub = lambda t: c(t)[:, heads]
st48 = [ck[:, :, CUT // G] for ck in whole[1]]
ref_tail = O.ttt_mlp_primal_backward_chunked(*[sub(d[n][:, :, CUT:]) for n in ("XQ", "XK", "XV", "le")], c(d["ln_w"])[heads], c(d["ln_b"])[heads],
                                             *[sub(s) for s in st48], sub(d["dOut"][:, :, CUT:]), G)
out = dict(heads=heads, XQ=sub(d["XQ"][:, :, :CUT]).float(), XK=sub(d["XK"][:, :, :CUT]).float(), XV=sub(d["XV"][:, :, :CUT]).float(),
           le=sub(d["le"][:, :, :CUT]).float(), dOut=sub(d["dOut"][:, :, :CUT]).float(), ln_w=c(d["ln_w"])[heads], ln_b=c(d["ln_b"])[heads],
           state=[sub(s) for s in state], carry=[ref_tail[n] for n in ("dW1", "db1", "dW2", "db2")],
           k_dXK=sub(g[8][:, :, :CUT]).float(), k_dXV=sub(g[7][:, :, :CUT]).float(), k_dXQ=sub(g[6][:, :, :CUT]).float(),
           k_dW1=sub(g[2]), k_dW2=sub(g[4]), k_out=sub(whole[0][:, :, :CUT]).float())
torch.save(out, os.path.join(ROOT, "gpurun_out", "r02_illcond_heads.pt"))
print("saved", {k: (tuple(v.shape) if torch.is_tensor(v) else type(v).__name__) for k, v in out.items()})
