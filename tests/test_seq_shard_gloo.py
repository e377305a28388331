"""CPU, world_size 2 and 3 over gloo: the sequence-sharded scan (partitioning, state hand-off order, head-group
pipeline, forward and reversed direction) reproduces the single-process scan.  The per-range scan is the oracle here
(tests may use it); on GPUs the same host logic drives the CUDA kernel (tests/test_gpu_seq_shard.py)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, NC, H, direction, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import ttt_oracle as O
    from ttt_video_dit_b200 import seq_shard
    d = O.make_inputs(1, H, NC, seed=11)
    le = d["eta"][:, :, :, -1, :, None]
    ranges = seq_shard.partition_minibatches(NC, world)
    # the reversed pass consumes the sequence from the end: rank r still owns range r, the chain runs world-1 -> 0
    s, e = ranges[rank]
    sl = lambda t: t[:, :, s:e].contiguous()
    q, k, v, l = sl(d["XQ"]), sl(d["XK"]), sl(d["XV"]), sl(le)
    if direction < 0:
        q, k, v, l = [t.flip(2) for t in (q, k, v, l)]

    def scan_fn(q_, k_, v_, l_, st):
        # which heads? recover from the state's provenance: match on ln slices via the closure counter
        h0 = scan_fn.next_h
        hn = q_.shape[1]
        scan_fn.next_h += hn
        lw, lb = d["ln_w"][h0:h0 + hn], d["ln_b"][h0:h0 + hn]
        out, _, last = O.ttt_mlp_primal_forward(q_, k_, v_, l_, lw, lb, *st, 1 << 30)
        return out, last
    scan_fn.next_h = 0
    if os.environ.get("TTT_TEST_TAKES_HEADS") == "1":  # the protocol cuda_scan_fn uses: heads=slice from sharded_scan
        def scan_fn(q_, k_, v_, l_, st, heads=None):
            assert heads is not None and heads.stop - heads.start == q_.shape[1]
            out, _, last = O.ttt_mlp_primal_forward(q_, k_, v_, l_, d["ln_w"][heads], d["ln_b"][heads], *st, 1 << 30)
            return out, last
        scan_fn.takes_heads = True
    init = (d["W1"], d["b1"], d["W2"], d["b2"])
    out, fin = seq_shard.sharded_scan(scan_fn, q, k, v, l, init, rank=rank, world=world, n_groups=2, direction=direction)
    gathered = [None] * world
    dist.all_gather_object(gathered, (out, fin))
    if rank == 0:
        outs = [g[0] for g in gathered]
        if direction < 0:
            outs = [o.flip(2) for o in outs]
        full = torch.cat(outs, dim=2)
        fins = [g[1] for g in gathered if g[1] is not None]
        torch.save((full, fins[0]), ret)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,NC,direction", [(2, 5, +1), (3, 7, +1), (2, 5, -1)])
def test_sharded_scan_matches_single_process(world, NC, direction):
    sys.path.insert(0, ROOT)
    from oracle import ttt_oracle as O
    from ttt_video_dit_b200 import seq_shard
    H = 3
    import tempfile
    ctx = mp.get_context("spawn")
    ret = os.path.join(tempfile.mkdtemp(), "out.pt")
    port = 29500 + (os.getpid() % 500) + world * 7 + (3 if direction < 0 else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, NC, H, direction, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    full, fin = torch.load(ret, weights_only=False)
    d = O.make_inputs(1, H, NC, seed=11)
    le = d["eta"][:, :, :, -1, :, None]
    q, k, v, l = d["XQ"], d["XK"], d["XV"], le
    if direction < 0:
        q, k, v, l = [t.flip(2) for t in (q, k, v, l)]
    ref, _, last = O.ttt_mlp_primal_forward(q, k, v, l, d["ln_w"], d["ln_b"], d["W1"], d["b1"], d["W2"], d["b2"], 1 << 30)
    if direction < 0:
        ref = ref.flip(2)
    assert O.rel_err(full, ref) < 1e-5
    for a, b in zip(fin, last):
        assert O.rel_err(a, b) < 1e-5


def test_sharded_scan_head_slice_protocol(monkeypatch):
    """A scan_fn flagged ``takes_heads`` receives the head slice of each pipeline group (what cuda_scan_fn relies on to pick
    its LayerNorm rows); spawned workers inherit the switch through the environment."""
    monkeypatch.setenv("TTT_TEST_TAKES_HEADS", "1")
    test_sharded_scan_matches_single_process(2, 6, +1)


def test_partition():
    from ttt_video_dit_b200 import seq_shard
    p = seq_shard.partition_minibatches(5487, 8)
    assert [e - s for s, e in p] == [686] * 7 + [685] and p[0][0] == 0 and p[-1][1] == 5487
    assert seq_shard.STATE_NUMEL * 4 == 132352
    assert [g.stop - g.start for g in seq_shard.head_groups(48, 8)] == [6] * 8
    # latency-bound regime: one group until a launch would oversubscribe the 148 SMs
    assert seq_shard.default_head_groups(1, 48) == 1 and seq_shard.default_head_groups(3, 48) == 1
    assert seq_shard.default_head_groups(8, 48) == 2 and seq_shard.default_head_groups(64, 48) == 20
