"""CPU, gloo: the TRAINABLE sequence-sharded chain (seq_shard.ShardedTTTMLP) -- forward state hand-off down the chain,
state-gradient hand-off back up -- reproduces the single-process scan and its analytic backward, for several micro-batches
in flight, both directions, and inside a sub-group of a larger job (global-rank translation).  The per-range math is the
oracle here; on GPUs the same host logic drives the CUDA kernels (tests/test_gpu_seq_shard.py)."""
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, M = 2, 3  # heads, micro-batches (independent sequences) in flight


class OracleRange:
    """seq_shard range implementation on the oracle (fp32, the dtype of the hand-off message): same protocol as seq_shard.CudaMLPRange."""

    def __init__(self, O, ln_w, ln_b):
        self.O, self.ln_w, self.ln_b = O, ln_w, ln_b

    def forward(self, q, k, v, le, state):
        out, _, last = self.O.ttt_mlp_primal_forward(q, k, v, le.unsqueeze(-1), self.ln_w, self.ln_b, *state, 1 << 30)
        return out, last, (q, k, v, le, state)

    def backward(self, ctx, go, d_state_out):
        q, k, v, le, state = ctx
        r = self.O.ttt_mlp_primal_backward(q, k, v, le.unsqueeze(-1), self.ln_w, self.ln_b, *state, go, dW_last=d_state_out)
        return r["dXQ"], r["dXK"], r["dXV"], r["dlast_eta"], (r["dW1"], r["db1"], r["dW2"], r["db2"]), r["dln_w"], r["dln_b"]


def _inputs(O, NC):
    ds = [O.make_inputs(1, H, NC, seed=20 + m, dtype=torch.float32) for m in range(M)]
    for d in ds[1:]:  # one set of parameters, M independent sequences
        for n in ("ln_w", "ln_b", "W1", "b1", "W2", "b2"):
            d[n] = ds[0][n]
    return ds


def _worker(rank, nproc, port, NC, direction, subgroup, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=nproc)
    torch.set_num_threads(1)
    from oracle import ttt_oracle as O
    from ttt_video_dit_b200 import seq_shard
    group, members = None, list(range(nproc))
    pairs = seq_shard.make_pair_groups() if (direction < 0 and not subgroup) else None  # one communicator per neighbour pair
    if subgroup:  # chain over global ranks [1, 2] of a 3-process job: positions 0,1 inside the group
        members = [1, 2]
        group = dist.new_group(members)
    result = None
    if rank in members:
        pos, world = members.index(rank), len(members)
        ds = _inputs(O, NC)
        s, e = seq_shard.partition_minibatches(NC, world)[pos]
        items, gouts = [], []
        for d in ds:
            t = [d[n][:, :, s:e] for n in ("XQ", "XK", "XV")] + [d["eta"][:, :, s:e, -1, :], d["dOut"][:, :, s:e]]
            if direction < 0:
                t = [x.flip(2) for x in t]
            items.append(tuple(x.contiguous() for x in t[:4]))
            gouts.append(t[4].contiguous())
        d0 = ds[0]
        stage = seq_shard.ShardedTTTMLP(OracleRange(O, d0["ln_w"], d0["ln_b"]), rank=pos, world=world, direction=direction, group=group,
                                        pair_groups=pairs)
        outs, finals = stage.forward(items, (d0["W1"], d0["b1"], d0["W2"], d0["b2"]))
        grads, d_init, dlw, dlb = stage.backward(gouts)
        result = (pos, outs, finals, grads, d_init, dlw, dlb)
    gathered = [None] * nproc
    dist.all_gather_object(gathered, result)
    if rank == 0:
        torch.save([g for g in gathered if g is not None], ret)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nproc,NC,direction,subgroup", [(2, 5, +1, False), (3, 7, -1, False), (3, 4, +1, True)])
def test_sharded_training_chain_matches_single_process(nproc, NC, direction, subgroup):
    sys.path.insert(0, ROOT)
    from oracle import ttt_oracle as O
    ctx = mp.get_context("spawn")
    ret = os.path.join(tempfile.mkdtemp(), "out.pt")
    port = 29900 + (os.getpid() % 300) + nproc * 11 + (5 if direction < 0 else 0) + (2 if subgroup else 0)
    procs = [ctx.Process(target=_worker, args=(r, nproc, port, NC, direction, subgroup, ret)) for r in range(nproc)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    res = sorted(torch.load(ret, weights_only=False), key=lambda r: r[0])
    world = len(res)
    ds = _inputs(O, NC)
    d0 = ds[0]
    fix = (lambda t: t.flip(2)) if direction < 0 else (lambda t: t)
    tot = {n: 0 for n in ("dln_w", "dln_b", "dW1", "db1", "dW2", "db2")}
    for m, d in enumerate(ds):
        q, k, v, dO = [fix(d[n]) for n in ("XQ", "XK", "XV", "dOut")]
        le = fix(d["eta"][:, :, :, -1, :]).unsqueeze(-1)
        ref_out, _, ref_last = O.ttt_mlp_primal_forward(q, k, v, le, d0["ln_w"], d0["ln_b"], d0["W1"], d0["b1"], d0["W2"], d0["b2"], 1 << 30)
        ref = O.ttt_mlp_primal_backward(q, k, v, le, d0["ln_w"], d0["ln_b"], d0["W1"], d0["b1"], d0["W2"], d0["b2"], dO)
        chain = res if direction > 0 else res[::-1]  # chain order = order of the (possibly reversed) sequence
        cat = lambda f: torch.cat([f(r) for r in chain], dim=2)
        assert O.rel_err(cat(lambda r: r[1][m]), ref_out) < 1e-5
        for i, name in enumerate(("dXQ", "dXK", "dXV", "dlast_eta")):
            got = cat(lambda r: r[3][m][i])
            assert O.rel_err(got.reshape(ref[name].shape), ref[name]) < 1e-4, name
        last_rank = chain[-1]
        for a, b in zip(last_rank[2][m], ref_last):
            assert O.rel_err(a, b) < 1e-5
        for n in tot:
            tot[n] = tot[n] + ref[n]
    first = (res if direction > 0 else res[::-1])[0]
    assert all(r[4] is None for r in res if r is not first)
    for got, name in zip(first[4], ("dW1", "db1", "dW2", "db2")):
        assert O.rel_err(got, tot[name]) < 1e-4, name
    assert O.rel_err(sum(r[5] for r in res), tot["dln_w"]) < 1e-4
    assert O.rel_err(sum(r[6] for r in res), tot["dln_b"]) < 1e-4
