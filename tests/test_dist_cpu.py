"""world_size-2 gloo tests (CPU) of the multi-rank plumbing bench.py uses: batch sharding with
disjoint proof indices, the max-over-ranks timing reduction, and the gather-then-add combine of
partial commitments used when one large argument is sharded (EC addition is not a reduce op)."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from bulletproofspp_b200.sharding import shard_range, proof_offset, combine_partials
    from oracle.curve import Secp256k1 as G
    from oracle.transcript import get_points
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    # (1) batch sharding: disjoint, contiguous, covers everything
    lo, hi = shard_range(10, rank, world)
    # (2) timing: max over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # (3) one MSM sharded by contiguous ranges of terms; partial points gathered then added
    pts = get_points(G, "test points", 10)
    sc = [(i * 7919 + 13) % G.order for i in range(10)]
    part = G.msm(list(zip(sc[lo:hi], pts[lo:hi])))
    total = combine_partials(part, G, dist)
    q.put((rank, (lo, hi), proof_offset(rank, 64), t.item(), total))
    dist.destroy_process_group()


def test_two_rank_plumbing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from oracle.curve import Secp256k1 as G
    from oracle.transcript import get_points
    pts = get_points(G, "test points", 10)
    sc = [(i * 7919 + 13) % G.order for i in range(10)]
    full = G.msm(list(zip(sc, pts)))
    assert [r[1] for r in res] == [(0, 5), (5, 10)]
    assert [r[2] for r in res] == [0, 64]
    assert all(r[3] == 2.0 for r in res)
    assert all(r[4] == full for r in res)


def test_sharded_argument_plan_is_consistent():
    """host-side planning of bppp_nl_prove_sharded (bulletproofspp_b200/sweep.py): for every size and world the ranks'
    slices tile the vector, are equal powers of two, the local rounds never fold a slice below one element, and the
    gathered argument fits the window table (<= 4096 norm elements) whenever the argument is larger than that"""
    from bulletproofspp_b200.sharding import shard_range
    from bulletproofspp_b200.sweep import sharded_local_rounds
    from bulletproofspp_b200.workloads import sweep_rounds
    for e in range(6, 23):
        for world in (1, 2, 4, 8):
            N, ln = 1 << e, (1 << e) // world
            spans = [shard_range(N, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi - lo == ln for lo, hi in spans) and ln & (ln - 1) == 0
            lr = sharded_local_rounds(e, world)
            assert 0 <= lr <= sweep_rounds(e) and (ln >> lr) >= 1
            gathered = (ln >> lr) * world
            assert gathered <= max(4096, world) or e <= 12
            if e > 12:
                assert gathered == 4096
