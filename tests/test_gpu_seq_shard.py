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
