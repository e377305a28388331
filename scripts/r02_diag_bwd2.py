"""Diagnostic (GPU): how the error of the carried state gradient develops along the reverse chain, per head."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import ttt_oracle as O
import test_gpu_full_shape as T
torch.set_num_threads(32)
H, NC, G = 48, 282, 16
d = T._bench_like_inputs(1, H, NC, seed=3)
state = [d[n] for n in ("W1", "b1", "W2", "b2")]
whole = T._forward(d, slice(0, NC), state, G)
c = lambda t: t.double().cpu()
relh = lambda a, b: [(float((a[:, h] - b[:, h]).norm() / (b[:, h].norm() + 1e-30))) for h in range(H)]
for cut in (272, 256, 224, 192, 128, 64, 16, 0):
    st = [ck[:, :, cut // G].contiguous() for ck in whole[1]]
    sl = slice(cut, NC)
    fw = T._forward(d, sl, st, G)
    g = T._backward(fw, d, sl, G)
    torch.cuda.synchronize()
    ref = O.ttt_mlp_primal_backward_chunked(*[c(d[n][:, :, sl]) for n in ("XQ", "XK", "XV", "le")], c(d["ln_w"]), c(d["ln_b"]),
                                            *[c(s) for s in st], c(d["dOut"][:, :, sl]), G)
    e1 = relh(c(g[2]), ref["dW1"]); e2 = relh(c(g[4]).reshape(ref["dW2"].shape), ref["dW2"]); eb = relh(c(g[5]).reshape(ref["db2"].shape), ref["db2"])
    srt = sorted(range(H), key=lambda h: -e1[h])
    print(json.dumps({"cut": cut, "steps": NC - cut, "dW1_all": round(O.rel_err(c(g[2]), ref["dW1"]), 5),
                      "dW1_worst": [(h, round(e1[h], 4)) for h in srt[:4]], "dW1_median": round(sorted(e1)[H // 2], 5),
                      "dW2_worst": round(max(e2), 4), "db2_worst": round(max(eb), 4),
                      "norm_dW1_worst_head": float(ref["dW1"][:, srt[0]].norm()), "norm_dW1_median": float(sorted(ref["dW1"][:, h].norm() for h in range(H))[H // 2])}), flush=True)
    if cut == 0:
        h = srt[0]
        per_step = [round(float((c(g[8])[:, h, t] - ref["dXK"][:, h, t]).norm() / ref["dXK"][:, h, t].norm()), 4) for t in range(0, 24)]
        print("worst head", h, "dXK rel err by step 0..23:", per_step)
        print("median head dXK by step:", [round(float((c(g[8])[:, srt[H // 2], t] - ref["dXK"][:, srt[H // 2], t]).norm() / ref["dXK"][:, srt[H // 2], t].norm()), 4) for t in range(0, 24)])
