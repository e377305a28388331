"""Diagnostic (GPU): where does the backward's error at the benchmarked shape come from -- sequence length or head count?
Prints rel errors vs the chunked analytic oracle for a matrix of (H, NC, G) and a bitwise determinism check."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ttt_oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_full_shape as T

names = ["dln_w", "dln_b", "dW1", "db1", "dW2", "db2", "dXQ", "dXV", "dXK", "dlast_eta"]
c = lambda t: t.float().cpu()
torch.set_num_threads(32)
for H, NC, G, dt in [(2, 64, 16, "f32"), (2, 141, 16, "f32"), (2, 282, 16, "f32"), (2, 282, 16, "f64"), (2, 282, 4, "f32"), (2, 282, 282, "f32"),
                     (48, 64, 16, "f32"), (48, 282, 16, "f32")]:
    d = T._bench_like_inputs(1, H, NC, seed=3)
    state = [d[n] for n in ("W1", "b1", "W2", "b2")]
    fw = T._forward(d, slice(0, NC), state, G)
    g = T._backward(fw, d, slice(0, NC), G)
    g2 = T._backward(fw, d, slice(0, NC), G)
    torch.cuda.synchronize()
    same = all(torch.equal(a, b) for a, b in zip(g, g2))
    cv = (lambda t: t.double().cpu()) if dt == "f64" else c
    ref = O.ttt_mlp_primal_backward_chunked(cv(d["XQ"]), cv(d["XK"]), cv(d["XV"]), cv(d["le"]), cv(d["ln_w"]), cv(d["ln_b"]),
                                            *[cv(s) for s in state], cv(d["dOut"]), min(G, 16))
    errs = {n: round(O.rel_err(cv(a).reshape(ref[n].shape), ref[n]), 5) for n, a in zip(names, g)}
    # error of dXK per quarter of the sequence (early steps depend on the longest gradient chain)
    q = NC // 4
    per = [round(O.rel_err(cv(g[8])[:, :, i * q:(i + 1) * q], ref["dXK"][:, :, i * q:(i + 1) * q]), 5) for i in range(4)]
    print(json.dumps({"H": H, "NC": NC, "G": G, "oracle": dt, "bitwise_repeatable": same, "errs": errs, "dXK_by_quarter": per}), flush=True)
