"""GPU (needs >= 2 devices): sequence-sharded forward scan with the NCCL state hand-off == the single-GPU scan, bit for
bit per head (same kernel, same state bits handed over), forward and reversed chain."""
import os
import sys
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, NC, H, direction, path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import ttt_oracle as O
    from ttt_video_dit_b200 import seq_shard
    dev = torch.device("cuda", rank)
    d = O.make_inputs(1, H, NC, seed=5)
    bf = lambda t: t.to(torch.bfloat16).to(dev).contiguous()
    q, k, v = bf(d["XQ"]), bf(d["XK"]), bf(d["XV"])
    le = bf(d["eta"][:, :, :, -1, :, None])
    if direction < 0:
        q, k, v, le = [t.flip(2).contiguous() for t in (q, k, v, le)]
    init = tuple(d[n].float().to(dev) for n in ("W1", "b1", "W2", "b2"))
    fn = seq_shard.cuda_scan_fn(d["ln_w"].to(dev), d["ln_b"].to(dev))
    # per-head-group ln slices: cuda_scan_fn receives the head slice in order; wrap to slice ln params
    groups = seq_shard.head_groups(H, 2)
    state = {"i": 0}

    def scan(q_, k_, v_, l_, st):
        g = groups[state["i"] % len(groups)]
        state["i"] += 1
        f = seq_shard.cuda_scan_fn(d["ln_w"][g].to(dev), d["ln_b"][g].to(dev))
        return f(q_, k_, v_, l_, st)
    ranges = seq_shard.partition_minibatches(NC, world)
    chain_pos = rank if direction > 0 else world - 1 - rank
    s, e = ranges[chain_pos]
    sl = lambda t: t[:, :, s:e].contiguous()
    out, fin = seq_shard.sharded_scan(scan, sl(q), sl(k), sl(v), sl(le), init, rank=rank, world=world, n_groups=2,
                                      direction=direction)
    torch.cuda.synchronize()
    if rank == 0:
        full_out, full_last = fn(q, k, v, le, init)   # single-GPU scan of the whole (possibly reversed) sequence
        torch.save((full_out.cpu(), [t.cpu() for t in full_last]), path + ".full")
    torch.save((out.cpu(), None if fin is None else [t.cpu() for t in fin], (s, e)), f"{path}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("direction", [+1, -1])
def test_two_gpu_chain_matches_single_gpu(direction):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, NC, H = 2, 9, 4
    path = os.path.join(tempfile.mkdtemp(), "shard")
    port = 29700 + os.getpid() % 200 + (1 if direction < 0 else 0)
    mp.spawn(_worker, args=(world, port, NC, H, direction, path), nprocs=world, join=True)
    full_out, full_last = torch.load(path + ".full", weights_only=False)
    for r in range(world):
        out, fin, (s, e) = torch.load(f"{path}.{r}", weights_only=False)
        assert torch.equal(out, full_out[:, :, s:e]), f"rank {r} output differs"
        if fin is not None:
            for a, b in zip(fin, full_last):
                assert torch.equal(a, b)


def _train_worker(rank, world, port, NC, H, M, path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import ttt_oracle as O
    from ttt_video_dit_b200 import seq_shard
    dev = torch.device("cuda", rank)
    bf = lambda t: t.to(torch.bfloat16).to(dev).contiguous()
    ds = [O.make_inputs(1, H, NC, seed=70 + m) for m in range(M)]
    d0 = ds[0]
    impl = seq_shard.CudaMLPRange(d0["ln_w"].to(dev), d0["ln_b"].to(dev), checkpoint_group_size=4)
    init = tuple(d0[n].float().to(dev) for n in ("W1", "b1", "W2", "b2"))
    s, e = seq_shard.partition_minibatches(NC, world)[rank]
    items = [tuple(bf(d[n][:, :, s:e]) for n in ("XQ", "XK", "XV")) + (bf(d["eta"][:, :, s:e, -1, :]),) for d in ds]
    gouts = [bf(d["dOut"][:, :, s:e]) for d in ds]
    stage = seq_shard.ShardedTTTMLP(impl, rank=rank, world=world)
    outs, finals = stage.forward(items, init)
    grads, d_init, dlw, dlb = stage.backward(gouts)
    torch.cuda.synchronize()
    cpu = lambda x: None if x is None else [t.cpu() if torch.is_tensor(t) else [u.cpu() for u in t] for t in x]
    torch.save(dict(outs=cpu(outs), grads=cpu(grads), d_init=cpu(d_init), dlw=dlw.cpu(), dlb=dlb.cpu(), range=(s, e)), f"{path}.{rank}")
    if rank == 0:  # single-GPU reference: the same kernels over the whole sequence of every item
        ref = []
        for d in ds:
            full = tuple(bf(d[n]) for n in ("XQ", "XK", "XV")) + (bf(d["eta"][:, :, :, -1, :]),)
            o, last, ctx = impl.forward(*full, init)
            g = impl.backward(ctx, bf(d["dOut"]), None)
            ref.append(dict(out=o.cpu(), g=[t.cpu() for t in g[:4]], d_in=[t.cpu() for t in g[4]], dlw=g[5].cpu(), dlb=g[6].cpu()))
        torch.save(ref, path + ".ref")
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_training_chain_matches_single_gpu():
    """Forward hand-off + backward hand-off (ttt_b200_mlp_backward_seeded) over NCCL, 3 sequences in flight: outputs and
    token gradients bit-identical to the un-sharded kernels, state gradient identical, LN-parameter sums to fp32 round-off."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import ttt_oracle as O
    world, NC, H, M = 2, 16, 4, 3  # ranges of 8 steps = two checkpoint groups of 4 each
    path = os.path.join(tempfile.mkdtemp(), "train")
    port = 29750 + os.getpid() % 200
    mp.spawn(_train_worker, args=(world, port, NC, H, M, path), nprocs=world, join=True)
    ref = torch.load(path + ".ref", weights_only=False)
    res = [torch.load(f"{path}.{r}", weights_only=False) for r in range(world)]
    for m in range(M):
        assert torch.equal(torch.cat([r["outs"][m] for r in res], dim=2), ref[m]["out"])
        for i, n in enumerate(("dXQ", "dXK", "dXV", "d_eta")):
            assert torch.equal(torch.cat([r["grads"][m][i] for r in res], dim=2), ref[m]["g"][i]), n
    assert res[1]["d_init"] is None
    for i in range(4):
        tot = sum(ref[m]["d_in"][i] for m in range(M))
        assert O.rel_err(res[0]["d_init"][i], tot) < 1e-6
    assert O.rel_err(res[0]["dlw"] + res[1]["dlw"], sum(r["dlw"] for r in ref)) < 1e-5
    assert O.rel_err(res[0]["dlb"] + res[1]["dlb"], sum(r["dlb"] for r in ref)) < 1e-5
