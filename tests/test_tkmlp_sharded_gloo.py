"""CPU, world_size 2 over gloo: TkMLP's and TritonLinear's head-sharded mode (``sharded_mode`` + ``local_map`` placements, reference
mlp_tk.py:297-404, linear_triton.py:262-362) hands the kernel boundary the LOCAL head shard of every DTensor argument and wraps the results back with
the reference's placements, forward and backward.  The kernel itself needs a GPU, so the two C-ABI entry points are replaced by
shape-recording stand-ins; what is tested is the DTensor plumbing around them."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.tensor import DTensor, Shard, distribute_tensor
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ttt_video_dit_b200 import mlp_tk

    B, H, NC, CS, F = 1, 4, 3, 64, 64
    seen = {}

    def fake_forward(ctx, lw, lb, W1, b1, W2, b2, XQ, XV, XK, last_eta, G):
        seen["fwd"] = [tuple(t.shape) for t in (lw, lb, W1, b1, W2, b2, XQ, XV, XK, last_eta)] + [G]
        assert not any(isinstance(t, DTensor) for t in (lw, lb, W1, XQ, last_eta))
        ctx.save_for_backward(XQ)
        return XQ * 2.0

    def fake_backward(ctx, g):
        (XQ,) = ctx.saved_tensors
        seen["bwd"] = tuple(g.shape)
        assert not isinstance(g, DTensor)
        Hl = XQ.shape[1]
        z = lambda *s: torch.zeros(*s)
        return (z(Hl, F), z(Hl, F), z(B, Hl, F, 4 * F), z(B, Hl, 1, 4 * F), z(B, Hl, 4 * F, F), z(B, Hl, 1, F),
                g * 2.0, z(*XQ.shape), z(*XQ.shape), z(B, Hl, NC, CS, 1))

    mlp_tk._forward_impl = fake_forward
    mlp_tk._backward_impl = fake_backward
    mesh = init_device_mesh("cpu", (world,))
    torch.manual_seed(0)
    full = dict(lw=torch.randn(H, F), lb=torch.randn(H, F), W1=torch.randn(B, H, F, 4 * F), b1=torch.randn(B, H, 1, 4 * F),
                W2=torch.randn(B, H, 4 * F, F), b2=torch.randn(B, H, 1, F), XQ=torch.randn(B, H, NC, CS, F),
                XV=torch.randn(B, H, NC, CS, F), XK=torch.randn(B, H, NC, CS, F), eta=torch.randn(B, H, NC, CS, CS))
    dt = {k: distribute_tensor(v, mesh, [Shard(0 if k in ("lw", "lb") else 1)]) for k, v in full.items()}
    dt["XQ"].requires_grad_(True)
    mlp_tk.TkMLP.sharded_mode = True
    try:
        out = mlp_tk.TkMLP.apply(dt["lw"], dt["lb"], dt["W1"], dt["b1"], dt["W2"], dt["b2"], dt["XQ"], dt["XV"], dt["XK"],
                                 dt["eta"], 2)
        assert isinstance(out, DTensor) and tuple(out.placements) == (Shard(1),) and tuple(out.shape) == (B, H, NC, CS, F)
        out.to_local().sum().backward()  # the upstream gradient arrives with the output's placement, as in the model
    finally:
        mlp_tk.TkMLP.sharded_mode = False
    g = dt["XQ"].grad
    assert isinstance(g, DTensor) and tuple(g.placements) == (Shard(1),)
    Hl = H // world
    assert seen["fwd"] == [(Hl, F), (Hl, F), (B, Hl, F, 4 * F), (B, Hl, 1, 4 * F), (B, Hl, 4 * F, F), (B, Hl, 1, F),
                           (B, Hl, NC, CS, F), (B, Hl, NC, CS, F), (B, Hl, NC, CS, F), (B, Hl, NC, CS, 1), 2], seen["fwd"]
    assert seen["bwd"] == (B, Hl, NC, CS, F)
    ok_out = torch.equal(out.full_tensor(), full["XQ"] * 2.0)
    ok_grad = torch.equal(g.full_tensor(), torch.full_like(full["XQ"], 2.0))

    # ---- TritonLinear (CS = 16, one-layer state), same protocol
    from ttt_video_dit_b200 import linear_triton
    CSl = 16
    seenl = {}

    def fake_linear_forward(XQ, XK, XV, last_eta, ln_w, ln_b, W1, b1, G, want_last=False):
        seenl["fwd"] = [tuple(t.shape) for t in (XQ, XK, XV, last_eta, ln_w, ln_b, W1, b1)] + [G]
        assert not any(isinstance(t, DTensor) for t in (XQ, last_eta, ln_w, W1))
        Bq, Hl = XQ.shape[:2]
        return XQ.float() * 3.0, (torch.zeros(Bq, Hl, 1, F, F), torch.zeros(Bq, Hl, 1, 1, F)), None

    def fake_linear_backward(XQ, XK, XV, last_eta, ln_w, ln_b, W1c, b1c, grad_out, G):
        seenl["bwd"] = tuple(grad_out.shape)
        Bq, Hl, NCl = XQ.shape[:3]
        z = lambda *sh: torch.zeros(*sh)
        return (z(Hl, F), z(Hl, F), z(Bq, Hl, F, F), z(Bq, Hl, 1, F), grad_out.float() * 3.0, z(*XQ.shape), z(*XQ.shape),
                z(Bq, Hl, NCl, CSl))

    linear_triton.linear_forward = fake_linear_forward
    linear_triton.linear_backward = fake_linear_backward
    fl = dict(lw=torch.randn(H, F), lb=torch.randn(H, F), W1=torch.randn(B, H, F, F), b1=torch.randn(B, H, 1, F),
              XQ=torch.randn(B, H, NC, CSl, F), XV=torch.randn(B, H, NC, CSl, F), XK=torch.randn(B, H, NC, CSl, F),
              eta=torch.randn(B, H, NC, CSl, CSl))
    dl = {k: distribute_tensor(v, mesh, [Shard(0 if k in ("lw", "lb") else 1)]) for k, v in fl.items()}
    dl["XQ"].requires_grad_(True)
    linear_triton.TritonLinear.sharded_mode = True
    try:
        outl = linear_triton.TritonLinear.apply(dl["lw"], dl["lb"], dl["W1"], dl["b1"], dl["XQ"], dl["XV"], dl["XK"], dl["eta"], 2)
        assert isinstance(outl, DTensor) and tuple(outl.placements) == (Shard(1),) and tuple(outl.shape) == (B, H, NC, CSl, F)
        outl.to_local().sum().backward()
    finally:
        linear_triton.TritonLinear.sharded_mode = False
    gl = dl["XQ"].grad
    assert isinstance(gl, DTensor) and tuple(gl.placements) == (Shard(1),)
    assert seenl["fwd"] == [(B, Hl, NC, CSl, F)] * 3 + [(B, Hl, NC, CSl), (Hl, F), (Hl, F), (B, Hl, F, F), (B, Hl, 1, F), 2], seenl["fwd"]
    assert seenl["bwd"] == (B, Hl, NC, CSl, F)
    ok_out = ok_out and torch.equal(outl.full_tensor(), fl["XQ"].to(torch.bfloat16).float() * 3.0)
    ok_grad = ok_grad and torch.equal(gl.full_tensor(), torch.full_like(fl["XQ"], 3.0))
    if rank == 0:
        torch.save((ok_out, ok_grad), ret)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mode_passes_local_head_shards():
    import tempfile
    ctx = mp.get_context("spawn")
    ret = os.path.join(tempfile.mkdtemp(), "ok.pt")
    port = 29100 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert torch.load(ret, weights_only=False) == (True, True)
