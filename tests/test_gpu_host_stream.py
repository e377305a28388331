"""GPU: HostPipeline (double-buffered H2D prefetch / asynchronous D2H around the op) returns, batch by batch, exactly what
the op returns when called directly on device tensors -- the copies move, the results do not."""
import pytest
import torch

from oracle import ttt_oracle as O
from ttt_video_dit_b200 import mlp_tk
from ttt_video_dit_b200.host_stream import HostPipeline

pytestmark = pytest.mark.gpu


def make_batches(n, B, H, NC):
    out = []
    for i in range(n):
        d = O.make_inputs(B, H, NC, seed=400 + i)
        pin = lambda t: t.to(torch.bfloat16).contiguous().pin_memory()
        out.append((pin(d["XQ"]), pin(d["XK"]), pin(d["XV"]), pin(d["eta"][:, :, :, -1:, :])))
    return out, d


@pytest.mark.parametrize("n,depth", [(1, 2), (5, 2), (7, 3)])
def test_pipeline_equals_direct_calls(n, depth):
    B, H, NC, G = 1, 4, 9, 4
    batches, d = make_batches(n, B, H, NC)
    dev = torch.device("cuda")
    prm = [d[k].float().to(dev) for k in ("ln_w", "ln_b", "W1", "b1", "W2", "b2")]
    fn = lambda q, k, v, e: mlp_tk.ttt_mlp_op(*prm, q, v, k, e, G)

    direct = []
    for hb in batches:
        direct.append(fn(*[t.to(dev) for t in hb]).cpu())
    host_out = [torch.zeros(B, H, NC, 64, 64, dtype=torch.bfloat16).pin_memory() for _ in range(n)]
    pipe = HostPipeline(dev, depth=depth)
    assert pipe.run(iter(batches), fn, host_out) == n
    torch.cuda.synchronize()
    for i in range(n):
        assert torch.isfinite(host_out[i].float()).all()
        assert torch.equal(host_out[i], direct[i]), f"batch {i} differs"
    assert pipe.h2d_bytes == n * sum(t.numel() * 2 for t in batches[0])
    assert pipe.d2h_bytes == n * host_out[0].numel() * 2


def test_pipeline_reuse_and_unpinned_rejected():
    B, H, NC, G = 1, 2, 3, 2
    batches, d = make_batches(3, B, H, NC)
    dev = torch.device("cuda")
    prm = [d[k].float().to(dev) for k in ("ln_w", "ln_b", "W1", "b1", "W2", "b2")]
    fn = lambda q, k, v, e: mlp_tk.ttt_mlp_op(*prm, q, v, k, e, G)
    host_out = torch.zeros(B, H, NC, 64, 64, dtype=torch.bfloat16).pin_memory()
    pipe = HostPipeline(dev)
    for _ in range(2):  # second run reuses the staging slots and events
        pipe.run(iter(batches), fn, host_out)
        torch.cuda.synchronize()
        assert torch.equal(host_out, fn(*[t.to(dev) for t in batches[-1]]).cpu())
    assert pipe.run(iter([]), fn, host_out) == 0
    with pytest.raises(ValueError):
        pipe.run(iter([tuple(t.clone() for t in batches[0])]), fn, host_out)
