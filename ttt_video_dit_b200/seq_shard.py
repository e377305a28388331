"""Sequence-sharded TTT scan across the GPUs of one box (north_star configs 4-5; NOT in the reference, which only
shards heads/batch -- SURVEY 2.3 "Sequence-sharding of the TTT scan with W hand-off: no").

The interleaved token sequence is cut into ``world`` contiguous ranges of mini-batches.  The recurrence is serial
across ranges, so the only data-path communication is the hand-off of the carried state
{W1[64,256], b1[256], W2[256,64], b2[64]} fp32 = 132 352 B per (batch, head) at each shard boundary: a point-to-point
chain (``torch.distributed`` send/recv: NCCL over NVLink on GPUs, gloo in the CPU tests).  Heads are independent, so
the chain can be pipelined over head groups: rank r works on group g while rank r+1 works on group g-1
(bubble = (world-1)/(groups+world-1)).  That only pays when a launch is throughput-bound: the scan is one latency chain
per (batch, head) and a B200 holds 148 of them at once, so for B*H <= 148 a group of 6 heads takes as long as all 48 and
the default is ONE group (``default_head_groups``).  For a single sequence the sharded scan therefore takes as long as
the un-sharded one (the recurrence is serial whichever GPU runs it): sequence sharding buys memory capacity and token
locality with the sequence-parallel attention, throughput comes from independent sequences (batch / data parallel).

``scan_fn(q, k, v, last_eta, state) -> (out, state_out)`` is injected: the product passes the CUDA kernel
(``cuda_scan_fn``); the CPU gloo tests pass the oracle so the partition / ordering / pipeline logic is covered without a
GPU.  A scan_fn with the attribute ``takes_heads = True`` is also given ``heads=slice`` -- the heads of the current
pipeline group -- so that it can pick its per-head parameters (``cuda_scan_fn`` does).
"""
from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist

STATE_SHAPES = ((64, 256), (1, 256), (256, 64), (1, 64))
STATE_NUMEL = sum(a * b for a, b in STATE_SHAPES)  # 33 088 floats = 132 352 B per (b, h)


def partition_minibatches(NC: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous ranges whose sizes differ by at most one, larger ones first (63 s: 5487 -> 7 x 686 + 685)."""
    base, rem = divmod(NC, world)
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((s, s + n))
        s += n
    return out


def head_groups(H: int, n_groups: int) -> List[slice]:
    n_groups = max(1, min(n_groups, H))
    base, rem = divmod(H, n_groups)
    out, s = [], 0
    for g in range(n_groups):
        n = base + (1 if g < rem else 0)
        out.append(slice(s, s + n))
        s += n
    return out


def pack_state(state: Sequence[torch.Tensor]) -> torch.Tensor:
    B, H = state[0].shape[:2]
    return torch.cat([t.reshape(B, H, -1).float() for t in state], dim=-1).contiguous()


def unpack_state(buf: torch.Tensor):
    B, H = buf.shape[:2]
    out, o = [], 0
    for a, b in STATE_SHAPES:
        out.append(buf[:, :, o:o + a * b].reshape(B, H, a, b).contiguous())
        o += a * b
    return tuple(out)


def default_head_groups(B: int, H: int, sms: int = 148) -> int:
    """Head groups of the hand-off pipeline.  One scan CTA per (batch, head) is a latency chain: a launch over 6 heads takes
    as long as one over 48, so splitting the heads only pays once a single launch would oversubscribe the SMs
    (B*H > sms).  Below that the whole range runs as one group -- the time of a range does not shrink with fewer heads,
    and every extra group would add a full range time to the chain."""
    return max(1, min(H, (B * H) // sms))


def sharded_scan(scan_fn: Callable, q, k, v, last_eta, init_state, *, rank: int, world: int, n_groups: int = None,
                 direction: int = +1, group=None):
    """Run this rank's range of the scan.  q,k,v: [B,H,NC_local,CS,F]; last_eta: [B,H,NC_local,CS,1];
    init_state: (W1,b1,W2,b2) used by the first rank of the chain only.  direction=+1: state flows rank 0 -> world-1
    (forward TTT pass); -1: world-1 -> 0 (the pass over the reversed sequence, whose first tokens live on the last rank).
    Returns (out [B,H,NC_local,CS,F], final_state or None) -- final_state only on the last rank of the chain."""
    B, H = q.shape[:2]
    if n_groups is None:
        n_groups = default_head_groups(B, H)
    chain = list(range(world)) if direction > 0 else list(range(world - 1, -1, -1))
    pos = chain.index(rank)
    prev_rank = chain[pos - 1] if pos > 0 else None
    next_rank = chain[pos + 1] if pos + 1 < world else None
    groups = head_groups(H, n_groups)
    out = torch.empty_like(q)
    recv_bufs, recv_work = [], []
    if prev_rank is not None:  # post every receive up front: the chain then runs at the pace of the slowest stage
        for g in groups:
            buf = torch.empty(B, g.stop - g.start, STATE_NUMEL, dtype=torch.float32, device=q.device)
            recv_bufs.append(buf)
            recv_work.append(dist.irecv(buf, src=prev_rank, group=group, tag=g.start))
    sends = []
    finals = []
    for gi, g in enumerate(groups):
        if prev_rank is None:
            st = tuple(t[:, g].contiguous() for t in init_state)
        else:
            recv_work[gi].wait()
            st = unpack_state(recv_bufs[gi])
        args = (q[:, g].contiguous(), k[:, g].contiguous(), v[:, g].contiguous(), last_eta[:, g].contiguous(), st)
        # per-head parameters (the LayerNorm weight / bias) live in the scan_fn: tell it which heads this call covers
        o, st_out = scan_fn(*args, heads=g) if getattr(scan_fn, "takes_heads", False) else scan_fn(*args)
        out[:, g] = o
        if next_rank is not None:
            buf = pack_state(st_out)
            sends.append((dist.isend(buf, dst=next_rank, group=group, tag=g.start), buf))
        else:
            finals.append(st_out)
    for w, _ in sends:
        w.wait()
    final_state = None
    if next_rank is None:
        final_state = tuple(torch.cat([f[i] for f in finals], dim=1) for i in range(4))
    return out, final_state


def cuda_scan_fn(ln_w, ln_b, checkpoint_group_size=1 << 30):
    """scan_fn backed by the sm_100a forward kernel (forward-only: one checkpoint group, final state exported)."""
    from . import test_time_training as tt
    ln_w, ln_b = ln_w.reshape(-1, 64), ln_b.reshape(-1, 64)  # [H, 64] (also accepts the [1, H, 1, 64] form)

    def fn(q, k, v, last_eta, st, heads=None):
        B, H, NC = q.shape[:3]
        lw_h, lb_h = (ln_w, ln_b) if heads is None else (ln_w[heads], ln_b[heads])
        if lw_h.shape[0] != H:
            raise RuntimeError(f"cuda_scan_fn: {lw_h.shape[0]} LayerNorm rows for {H} heads (pass heads= or pre-sliced parameters)")
        dev = q.device
        G = min(checkpoint_group_size, NC)
        K = (NC + G - 1) // G
        out = torch.empty_like(q)
        ck = [torch.empty(B, H, K, a, b, device=dev, dtype=torch.float32) for a, b in STATE_SHAPES]
        last = [torch.empty(B, H, a, b, device=dev, dtype=torch.float32) for a, b in STATE_SHAPES]
        lw = lw_h.reshape(1, -1, 1, 64).float().contiguous()
        lb = lb_h.reshape(1, -1, 1, 64).float().contiguous()
        tt.ttt_forward(q, k, v, last_eta, lw, lb, *[s.float().contiguous() for s in st], *ck, out, G, W_last=last)
        return out, tuple(last)
    fn.takes_heads = True
    return fn
