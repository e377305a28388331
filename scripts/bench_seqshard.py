#!/usr/bin/env python
"""Sequence-sharded forward scan throughput (BASELINE configs 4-5) -- one rank per GPU under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/bench_seqshard.py --nc 2630 --steps 10 --warmup 3 [--groups 8] [--direction 1]

Each rank holds its contiguous range of mini-batches (seq_shard.partition_minibatches) of one B=1, 48-head sequence and
runs seq_shard.sharded_scan with the CUDA scan; the only data-path traffic is the 132 352-B state per head at each range
boundary.  Prints one JSON line on rank 0: whole-sequence tokens/s (max over ranks of the CUDA-event time), the
single-GPU time of the same sequence for reference when it fits (--single), and the ideal pipeline bound
T_single * (groups + N - 1) / (groups * N).

NOT YET RUN ON HARDWARE (written after the round-1 GPU budget was spent); the pieces it composes are the ones
tests/test_gpu_seq_shard.py verifies on 2 GPUs.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nc", type=int, default=2630, help="mini-batches of the whole sequence (2630 = 30 s, 5487 = 63 s)")
    ap.add_argument("--heads", type=int, default=48)
    ap.add_argument("--groups", type=int, default=1, help="head groups of the hand-off pipeline (1: latency-bound regime)")
    ap.add_argument("--direction", type=int, default=1, choices=[1, -1])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--single", action="store_true", help="also time the un-sharded scan on rank 0")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import __graft_entry__
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        __graft_entry__.build()
    dist.barrier()
    from ttt_video_dit_b200 import seq_shard

    H, NC = args.heads, args.nc
    ranges = seq_shard.partition_minibatches(NC, world)
    pos = rank if args.direction > 0 else world - 1 - rank  # the reversed pass starts on the last rank
    s, e = ranges[pos]
    n = e - s
    gen = torch.Generator().manual_seed(1234 + pos)
    rn = lambda *sh: torch.randn(*sh, generator=gen)
    bf = lambda t: t.to(torch.bfloat16).to(dev).contiguous()
    q = bf(torch.nn.functional.normalize(rn(1, H, n, 64, 64), dim=-1))
    k = bf(torch.nn.functional.normalize(rn(1, H, n, 64, 64), dim=-1))
    v = bf(rn(1, H, n, 64, 64))
    le = bf((0.1 / 64) * torch.sigmoid(rn(1, H, n, 64, 1)) / 64)
    pg = torch.Generator().manual_seed(7)  # parameters identical on every rank
    ln_w = (1 + 0.1 * torch.randn(H, 64, generator=pg)).to(dev)
    ln_b = (0.1 * torch.randn(H, 64, generator=pg)).to(dev)
    init = ((0.02 * torch.randn(1, H, 64, 256, generator=pg)).to(dev), torch.zeros(1, H, 1, 256, device=dev),
            (0.02 * torch.randn(1, H, 256, 64, generator=pg)).to(dev), torch.zeros(1, H, 1, 64, device=dev))
    fn = seq_shard.cuda_scan_fn(ln_w, ln_b)

    def step():
        return seq_shard.sharded_scan(fn, q, k, v, le, init, rank=rank, world=world, n_groups=args.groups,
                                      direction=args.direction)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t) / args.steps

    single_ms = None
    if args.single and rank == 0:
        g2 = torch.Generator().manual_seed(99)
        rn2 = lambda *sh: torch.randn(*sh, generator=g2)
        fq = bf(torch.nn.functional.normalize(rn2(1, H, NC, 64, 64), dim=-1))
        fk = bf(torch.nn.functional.normalize(rn2(1, H, NC, 64, 64), dim=-1))
        fv = bf(rn2(1, H, NC, 64, 64))
        fl = bf((0.1 / 64) * torch.sigmoid(rn2(1, H, NC, 64, 1)) / 64)
        for _ in range(2):
            fn(fq, fk, fv, fl, init)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            fn(fq, fk, fv, fl, init)
        b.record()
        torch.cuda.synchronize()
        single_ms = a.elapsed_time(b) / args.steps
    if rank == 0:
        line = {"metric": "video-tokens/sec TTT-MLP layer-direction (fwd), sequence-sharded", "value": NC * 64 / (ms * 1e-3),
                "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "scaling": "strong", "config": {"workload": f"B=1, {H} heads, NC={NC} split over {world} ranks, {args.groups} head groups, "
                                                            f"direction {args.direction}", "handoff_bytes_per_boundary": H * seq_shard.STATE_NUMEL * 4},
                "single_gpu_ms": single_ms,
                "pipeline_bound_ms": None if single_ms is None else single_ms * (args.groups + world - 1) / (args.groups * world)}
        print(json.dumps(line))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
