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


def chain_neighbors(rank: int, world: int, direction: int, group=None):
    """(previous, next) peer of chain position ``rank`` as GLOBAL ranks (what dist.isend / irecv take), None at the ends.
    ``rank`` / ``world`` are positions inside ``group`` (the sequence-parallel group of a larger job); messages of one
    pair are matched by issue order only -- NCCL ignores tags."""
    chain = list(range(world)) if direction > 0 else list(range(world - 1, -1, -1))
    pos = chain.index(rank)
    to_global = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    prev_rank = to_global(chain[pos - 1]) if pos > 0 else None
    next_rank = to_global(chain[pos + 1]) if pos + 1 < world else None
    return prev_rank, next_rank


def sharded_scan(scan_fn: Callable, q, k, v, last_eta, init_state, *, rank: int, world: int, n_groups: int = None,
                 direction: int = +1, group=None):
    """Run this rank's range of the scan.  q,k,v: [B,H,NC_local,CS,F]; last_eta: [B,H,NC_local,CS,1];
    init_state: (W1,b1,W2,b2) used by the first rank of the chain only.  direction=+1: state flows rank 0 -> world-1
    (forward TTT pass); -1: world-1 -> 0 (the pass over the reversed sequence, whose first tokens live on the last rank).
    Returns (out [B,H,NC_local,CS,F], final_state or None) -- final_state only on the last rank of the chain."""
    B, H = q.shape[:2]
    if n_groups is None:
        n_groups = default_head_groups(B, H)
    prev_rank, next_rank = chain_neighbors(rank, world, direction, group)
    groups = head_groups(H, n_groups)
    out = torch.empty_like(q)
    sends = []
    finals = []
    for gi, g in enumerate(groups):
        if prev_rank is None:
            st = tuple(t[:, g].contiguous() for t in init_state)
        else:
            # receive of group gi is posted when it is needed (not all up front): NCCL runs the p2p operations of one
            # communicator in issue order, so a receive posted early would sit in front of this rank's own sends
            buf = torch.empty(B, g.stop - g.start, STATE_NUMEL, dtype=torch.float32, device=q.device)
            dist.irecv(buf, src=prev_rank, group=group).wait()
            st = unpack_state(buf)
        args = (q[:, g].contiguous(), k[:, g].contiguous(), v[:, g].contiguous(), last_eta[:, g].contiguous(), st)
        # per-head parameters (the LayerNorm weight / bias) live in the scan_fn: tell it which heads this call covers
        o, st_out = scan_fn(*args, heads=g) if getattr(scan_fn, "takes_heads", False) else scan_fn(*args)
        out[:, g] = o
        if next_rank is not None:
            buf = pack_state(st_out)
            sends.append((dist.isend(buf, dst=next_rank, group=group), buf))
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


def make_pair_groups(ranks=None):
    """One two-rank process group per pair of chain neighbours (collective: every rank of the default group must call it
    with the same ``ranks``, the global ranks of the chain in order).  Returns {frozenset({a, b}): group}."""
    ranks = list(range(dist.get_world_size())) if ranks is None else list(ranks)
    return {frozenset((a, b)): dist.new_group([a, b]) for a, b in zip(ranks[:-1], ranks[1:])}


# ------------------------------------------------------------------------------------------------ training through the chain
class CudaMLPRange:
    """Forward + backward of ONE mini-batch range of one (micro-)batch on the sm_100a kernels: the forward writes the fp32
    checkpoints and exports the final state (the forward hand-off message); the backward takes the upstream gradient of that
    final state (ttt_b200_mlp_backward_seeded) and returns the gradient of the state it started from (the backward
    hand-off message)."""

    def __init__(self, ln_w, ln_b, checkpoint_group_size=16):
        self.ln_w, self.ln_b = ln_w.reshape(-1, 64), ln_b.reshape(-1, 64)
        self.G = int(checkpoint_group_size)

    def forward(self, q, k, v, last_eta, state):
        from . import test_time_training as tt
        B, H, NC = q.shape[:3]
        dev, f32 = q.device, torch.float32
        G = min(self.G, NC)
        K = (NC + G - 1) // G
        out = torch.empty_like(q)
        ck = [torch.empty(B, H, K, a, b, device=dev, dtype=f32) for a, b in STATE_SHAPES]
        last = [torch.empty(B, H, a, b, device=dev, dtype=f32) for a, b in STATE_SHAPES]
        lw = self.ln_w.reshape(1, H, 1, 64).float().contiguous()
        lb = self.ln_b.reshape(1, H, 1, 64).float().contiguous()
        le = last_eta.reshape(B, H, NC, 64, 1)
        tt.ttt_forward(q, k, v, le, lw, lb, *[s.float().contiguous() for s in state], *ck, out, G, W_last=last)
        return out, tuple(last), (q, k, v, le, lw, lb, ck, G)

    def backward(self, ctx, grad_out, d_state_out):
        from . import test_time_training as tt
        q, k, v, le, lw, lb, ck, G = ctx
        dlw, dlb, dW1, db1, dW2, db2, dq, dv, dk, de = tt.ttt_backward_simple(q, k, v, le, lw, lb, *ck, grad_out, G,
                                                                                dW_last=d_state_out)
        return dq, dk, dv, de, (dW1, db1, dW2, db2), dlw, dlb


class ShardedTTTMLP:
    """This rank's stage of the sequence-sharded TTT-MLP scan, trainable: forward state hand-off down the chain, gradient
    of the state handed back up the chain.  ``items`` are independent (micro-)batches -- different sequences -- each rank
    holding ITS mini-batch range of every one of them; because sends are asynchronous the ranks form a pipeline over the
    items (rank r works on item m while rank r+1 works on item m-1), so with M items in flight the serial chain is busy
    M / (M + world - 1) of the time (the hand-off itself is 6.35 MB per item and boundary).

    ``impl`` provides forward(q,k,v,last_eta,state) -> (out, state_out, ctx) and backward(ctx, grad_out, d_state_out) ->
    (dq, dk, dv, d_eta, d_state_in, d_ln_w, d_ln_b): CudaMLPRange in the product, the oracle in the CPU gloo tests."""

    def __init__(self, impl, *, rank: int, world: int, direction: int = +1, group=None, pair_groups=None):
        """``pair_groups`` (optional): {frozenset({global_rank_a, global_rank_b}): two-rank process group} from
        ``make_pair_groups``.  NCCL runs all point-to-point operations of ONE communicator in issue order on one stream, so on
        a shared group a rank's send to its successor queues behind its receive from its predecessor; a communicator per
        neighbour pair makes the two directions of a stage independent."""
        self.impl, self.group = impl, group
        self.prev, self.next = chain_neighbors(rank, world, direction, group)
        me = dist.get_rank() if dist.is_initialized() else rank
        pg = pair_groups or {}
        self._pg_prev = pg.get(frozenset((me, self.prev)), group) if self.prev is not None else None
        self._pg_next = pg.get(frozenset((me, self.next)), group) if self.next is not None else None
        self._ctx = []
        self._pending = []  # (work, buffer) of sends not yet known to be complete

    def _pg(self, peer):
        return self._pg_prev if peer == self.prev else self._pg_next

    def _recv_state(self, B, H, device, src):
        buf = torch.empty(B, H, STATE_NUMEL, dtype=torch.float32, device=device)
        dist.irecv(buf, src=src, group=self._pg(src)).wait()
        return unpack_state(buf)

    def _send_state(self, state, dst):
        buf = pack_state(state)
        self._pending.append((dist.isend(buf, dst=dst, group=self._pg(dst)), buf))

    def _drain(self):
        for w, _ in self._pending:
            w.wait()
        self._pending = []

    def forward(self, items, init_state, collect=True):
        """items: list of (q, k, v, last_eta) local ranges [B,H,NC_local,CS,F] / [B,H,NC_local,CS]; init_state: (W1,b1,W2,b2)
        [B,H,...], read by the first rank of the chain only.  Returns (outs, final_states or None): final_states (one per
        item) only on the last rank of the chain.  ``collect=False`` drops each item's output as soon as it is produced
        (benchmarks; a caller that consumes outputs item by item)."""
        outs, finals = [], []
        self._ctx = []
        for q, k, v, le in items:
            B, H = q.shape[:2]
            st = tuple(t.contiguous() for t in init_state) if self.prev is None else self._recv_state(B, H, q.device, self.prev)
            out, st_out, ctx = self.impl.forward(q, k, v, le, st)
            self._ctx.append(ctx)
            if collect:
                outs.append(out)
            del out
            if self.next is not None:
                self._send_state(st_out, self.next)
            else:
                finals.append(st_out)
        self._drain()
        return outs, (finals if self.next is None else None)

    def backward(self, grad_outs, d_final_states=None, collect=True):
        """grad_outs: one upstream gradient per item (layout of the outputs).  d_final_states: optional per-item upstream
        gradient of the final state, last rank of the chain only (zero at the reference's op boundary).  Returns
        (item_grads = [(dq, dk, dv, d_last_eta)], d_init_state or None, d_ln_w, d_ln_b): d_init_state (summed over items: the
        gradient of the shared initial-state parameters) only on the first rank of the chain; d_ln_* are this rank's
        partial sums (sum them over the ranks of the chain)."""
        item_grads, d_init, dlw, dlb = [], None, None, None
        for m, go in enumerate(grad_outs):
            B, H = go.shape[:2]
            if self.next is not None:
                d_out_state = self._recv_state(B, H, go.device, self.next)
            else:
                d_out_state = None if d_final_states is None else d_final_states[m]
            dq, dk, dv, de, d_in, a, b = self.impl.backward(self._ctx[m], go, d_out_state)
            self._ctx[m] = None
            if collect:  # collect=False: the token gradients of an item are dropped once its state gradient is on its way
                item_grads.append((dq, dk, dv, de))
            del dq, dk, dv, de
            dlw = a if dlw is None else dlw + a
            dlb = b if dlb is None else dlb + b
            if self.prev is not None:
                self._send_state(d_in, self.prev)
            else:
                d_init = tuple(d_in) if d_init is None else tuple(x + y for x, y in zip(d_init, d_in))
        self._drain()
        return item_grads, d_init, dlw, dlb
