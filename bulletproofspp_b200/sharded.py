"""One large NormLinear argument sharded over several GPUs.

The fold pairs ADJACENT elements (src/Bulletproof.hs:77-90), so rank r keeps the contiguous slice
[r*N/W, (r+1)*N/W) of the norm vector and of G and folds it locally for log2(N/W) rounds with no
exchange.  Per round the commitments are sums of per-rank partial MSMs: 2 points (128 bytes) per
rank are all-gathered and added (EC addition is not an NCCL reduction).  The linear part (M <= a few
hundred) and the scalar on g live on rank 0.  When a local slice is down to one element the stored
state is gathered and rank 0 finishes the remaining rounds.

`gather(obj) -> [obj of rank 0, ..., obj of rank W-1]` abstracts the collective: a closure over
torch.distributed.all_gather_object / all_gather under torchrun, or a local list when several
"ranks" are emulated on one GPU (tests).
"""
from . import lib as L
from .lib import NormLinearArgument, ARG_NL

R_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141


class Shard:
    """rank-local part of the argument"""

    def __init__(self, ctx, rank, world, N, g, G_slice, H, q, s, w_slice, l, c):
        assert N % world == 0 and (N // world) & (N // world - 1) == 0, "shards must be equal powers of two"
        self.ctx, self.rank, self.world, self.N = ctx, rank, world, N
        self.len0 = N // world
        own_lin = rank == 0
        self.arg = NormLinearArgument(ctx, ARG_NL, g, G_slice, H if own_lin else [], [q], [s if own_lin else 0],
                                      [w_slice], [l if own_lin else []], [c if own_lin else []])
        self.arg.set_shard(rank * self.len0)
        self.g = g

    def local_rounds(self, total_rounds):
        n = 0
        while n < total_rounds and (self.len0 >> (n + 1)) >= 1:
            n += 1
        return n


def add_points(ctx, pts):
    pts = [p for p in pts if p is not None]
    if not pts:
        return None
    return ctx.msm([(1, p) for p in pts])


def prove_sharded(shards, gather, total_rounds, oracle, q, M):
    """Runs the prover.  `shards`: the Shard objects this process drives (one under torchrun, W when
    emulating); `gather(list_of_local_values) -> list over all ranks`; `oracle(X, R) -> e`.
    Returns (responses newest first, final scalar s, final norm witness, final linear witness) on
    every rank."""
    ctx = shards[0].ctx
    k_local = shards[0].local_rounds(total_rounds)
    resp = []
    qc = q
    for r in range(k_local):
        parts = gather([sh.arg.round_commit() for sh in shards])          # [( [X], [R] ) per rank]
        X = add_points(ctx, [p[0][0] for p in parts])
        Rr = add_points(ctx, [p[1][0] for p in parts])
        e = oracle(X, Rr)
        resp.insert(0, (X, Rr))
        for sh in shards:
            sh.arg.round_fold([e])
        qc = qc * qc % R_ORDER
    finals = gather([sh.arg.final() for sh in shards])                     # (s, w, l) true terms per rank
    s_tot = sum(f[0][0] for f in finals) % R_ORDER
    w_all = [v for f in finals for v in f[1][0]]
    l_all = finals[0][2][0] if M else []
    if k_local == total_rounds:
        return resp, s_tot, w_all, l_all
    # remaining rounds on the gathered state (a handful of elements): true-term generators
    exps = gather([sh.arg.export() for sh in shards])                      # (nn, nl, points, c) per rank
    gens_n, gens_l, c_true = [], [], []
    for rk, (nn, nl, pts, cs) in enumerate(exps):
        n_n = len(finals[rk][1][0])
        inv_n = pow(nn[0], -1, R_ORDER)
        gens_n += ctx.msm_batch([[inv_n]] * n_n, [[p] for p in pts[0][:n_n]], shared_points=False) if n_n else []
        if rk == 0 and M:
            inv_l = pow(nl[0], -1, R_ORDER)
            m = len(pts[0]) - n_n
            gens_l = ctx.msm_batch([[inv_l]] * m, [[p] for p in pts[0][n_n:]], shared_points=False)
            c_true = [cv * inv_l % R_ORDER for cv in cs[0]]
    tail = NormLinearArgument(ctx, ARG_NL, shards[0].g, gens_n, gens_l, [qc], [s_tot], [w_all], [l_all], [c_true])
    for r in range(total_rounds - k_local):
        X, Rr = tail.round_commit()
        e = oracle(X[0], Rr[0])
        resp.insert(0, (X[0], Rr[0]))
        tail.round_fold([e])
    s, fw, fl = tail.final()
    tail.close()
    return resp, s[0], fw[0], fl[0]
