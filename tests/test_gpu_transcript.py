"""Device transcript: SHA-256, decimal `show`, hash-to-curve generators, `oracle` and `random`
on the GPU, byte-identical to the oracle (oracle/transcript.py, following app/Main.hs:64-87 and
src/ZKP.hs:90-101) and to the host C++ transcript."""
import ctypes as C
import hashlib

import pytest

pytestmark = pytest.mark.gpu

from oracle.curve import Secp256k1 as G
from oracle.field import Q, R
from oracle.transcript import BARE_DECIMAL, PREFIXED_P, ZKPT, get_points


def H(*a):
    return int.from_bytes(hashlib.sha256(repr(a).encode()).digest(), "big")


@pytest.mark.parametrize("policy,name", [(0, "exp"), (1, "even"), (2, "smaller")])
def test_device_generators_match_get_points(ctx, policy, name):
    from bulletproofspp_b200 import lib as L
    n = 300
    out = C.create_string_buffer(64 * n)
    ctx._ck(ctx.lib.bppp_get_points(ctx.h, b"test points", n, policy, out), "bppp_get_points")
    dev = L.bytes_to_points(out.raw[:64 * n])
    assert dev[:40] == get_points(G, "test points", 40, name)
    host = C.create_string_buffer(64 * n)
    assert ctx.lib.bppp_host_get_points(b"test points", n, policy, host) == 0
    assert out.raw == host.raw
    assert all(G.on_curve(p) for p in dev)
    # another seed, a count that needs more than one candidate chunk boundary to line up
    out2 = C.create_string_buffer(64 * 5)
    ctx._ck(ctx.lib.bppp_get_points(ctx.h, b"another basis seed with a longer name", 5, policy, out2), "bppp_get_points")
    assert L.bytes_to_points(out2.raw[:320]) == get_points(G, "another basis seed with a longer name", 5, name)


@pytest.mark.parametrize("fmt,ofmt", [(0, PREFIXED_P), (1, BARE_DECIMAL)])
def test_device_oracle_matches_transcript(ctx, gens, fmt, ofmt):
    """a prover-shaped sequence of `oracle` calls (130 commitments with three challenges, single commitments,
    then (X, R) rounds) for three proofs in lock-step; coordinates of every decimal length incl. tiny ones"""
    from bulletproofspp_b200 import lib as L
    B = 3
    base = gens(40)
    # points with short decimal renderings do not lie on the curve: the transcript only renders coordinates
    odd = [(7, 11), (0, 0), (10 ** 18, 10 ** 19 - 1), (Q - 1, 1), (10 ** 76, 10 ** 77 + 5), (2 ** 64, 2 ** 32 - 1)]
    calls = [(130, 3), (1, 3), (1, 1)] + [(2, 1)] * 9 + [(0, 2)]
    t = C.c_void_p()
    ctx._ck(ctx.lib.bppp_dtr_create(ctx.h, B, 200, fmt, C.byref(t)), "bppp_dtr_create")
    zks = [ZKPT(G, None, ofmt) for _ in range(B)]
    for ci, (npts, count) in enumerate(calls):
        pts = [[(odd + base)[H("p", ci, b, j) % 46] if j % 3 else base[H("q", ci, b, j) % 40] for j in range(npts)] for b in range(B)]
        raw = b"".join(b"".join(L.int_to_le(x) + L.int_to_le(y) for x, y in row) for row in pts)
        out = C.create_string_buffer(32 * B * count)
        ctx._ck(ctx.lib.bppp_dtr_oracle(t, raw if npts else None, npts, count, out), "bppp_dtr_oracle")
        got = L.bytes_to_ints(out.raw[:32 * B * count])
        for b in range(B):
            assert got[b * count:(b + 1) * count] == zks[b].oracle(pts[b], count), "call %d proof %d" % (ci, b)
    # reset: a fresh transcript again
    ctx.lib.bppp_dtr_reset(t)
    out = C.create_string_buffer(32 * B)
    row = [base[0], base[1]]
    raw = b"".join(L.point_to_bytes(p) for p in row) * B
    ctx._ck(ctx.lib.bppp_dtr_oracle(t, raw, 2, 1, out), "bppp_dtr_oracle")
    assert L.bytes_to_ints(out.raw[:32 * B]) == [ZKPT(G, None, ofmt).oracle(row, 1)[0]] * B
    ctx.lib.bppp_dtr_destroy(t)


def test_both_squeeze_kernels_agree_with_the_oracle(ctx, gens):
    """up to 296 hashes per launch take k_tr_squeeze_coop (a CTA per hash: schedule and rounds on separate warps),
    larger launches k_tr_squeeze (a thread per hash); 110 proofs with 1, 2 and 3 challenges per call cross the
    switch in both directions on one transcript, with bodies from one block to several tiles of 32 blocks"""
    from bulletproofspp_b200 import lib as L
    B = 110
    base = gens(40)
    calls = [(0, 1), (1, 3), (20, 3), (2, 1), (1, 2), (30, 2), (0, 3)]
    t = C.c_void_p()
    ctx._ck(ctx.lib.bppp_dtr_create(ctx.h, B, 60, 0, C.byref(t)), "bppp_dtr_create")
    zks = [ZKPT(G, None, PREFIXED_P) for _ in range(B)]
    for ci, (npts, count) in enumerate(calls):
        pts = [[base[H("k", ci, b, j) % 40] for j in range(npts)] for b in range(B)]
        raw = b"".join(b"".join(L.point_to_bytes(p) for p in row) for row in pts)
        out = C.create_string_buffer(32 * B * count)
        ctx._ck(ctx.lib.bppp_dtr_oracle(t, raw if npts else None, npts, count, out), "bppp_dtr_oracle")
        got = L.bytes_to_ints(out.raw[:32 * B * count])
        for b in range(B):
            assert got[b * count:(b + 1) * count] == zks[b].oracle(pts[b], count), "call %d proof %d" % (ci, b)
    ctx.lib.bppp_dtr_destroy(t)


def test_device_random_matches_zkpt(ctx):
    from bulletproofspp_b200 import lib as L
    seeds = ["default random seed", "default random seed#17", "s", "x" * 40]
    B, n0, count = len(seeds), 95, 1200                 # counters crossing 99 -> 100 -> 1000 digits
    arr = (C.c_char_p * B)(*[s.encode() for s in seeds])
    out = C.create_string_buffer(32 * B * count)
    ctx._ck(ctx.lib.bppp_dev_random(ctx.h, B, arr, n0, count, out), "bppp_dev_random")
    got = L.bytes_to_ints(out.raw[:32 * B * count])
    for b, s in enumerate(seeds):
        zk = ZKPT(G, s)
        zk.n = n0
        assert got[b * count:(b + 1) * count] == [zk.random() for _ in range(count)]


def test_device_transcript_and_round_loop_refuse_misuse(ctx, gens):
    """error behaviour of the round-2 entry points: capacity, stale handles, wrong shard -- an error code and a
    message, never a crash or a silent wrong answer"""
    from bulletproofspp_b200 import lib as L, sweep, workloads as W
    from bulletproofspp_b200.lib import ARG_NL
    lib = ctx.lib
    base = gens(8)
    t = C.c_void_p()
    ctx._ck(lib.bppp_dtr_create(ctx.h, 2, 3, 0, C.byref(t)), "bppp_dtr_create")
    raw = b"".join(L.point_to_bytes(p) for p in base[:2]) * 2
    ctx._ck(lib.bppp_dtr_absorb(t, raw, 2, 2), "bppp_dtr_absorb")
    assert lib.bppp_dtr_absorb(t, raw, 2, 2) != 0                       # 4 commitments > capacity 3
    assert b"capacity" in lib.bppp_last_error(ctx.h)
    out = C.create_string_buffer(64)
    assert lib.bppp_dtr_squeeze(t, 1, bytes([10]), None, out) != 0      # scalar index > 9
    assert lib.bppp_dtr_squeeze(t, 1, bytes([1]), bytes([5]), out) != 0  # a stage that does not exist yet
    bad = bytearray(raw)
    bad[0:32] = (2 ** 256 - 1).to_bytes(32, "little")                    # non-canonical coordinate
    lib.bppp_dtr_reset(t)
    assert lib.bppp_dtr_absorb(t, bytes(bad), 2, 2) != 0
    lib.bppp_dtr_destroy(t)
    # round loop: needs a transcript and a fresh handle
    e, M = 5, 6
    N = 1 << e
    points = W.sweep_generators(ctx, 1 + N + M)
    inp = W.sweep_inputs(ctx, e, M)
    gh, h, t2 = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx._ck(lib.bppp_gens_create(ctx.h, N, M, points[:64], points[64:64 * (1 + N)], points[64 * (1 + N):], C.byref(gh)), "gens")
    ctx._ck(lib.bppp_nl_create_gens(gh, ARG_NL, 1, inp["q"], inp["s"], inp["w"], inp["l"], inp["c"], C.byref(h)), "create")
    k = inp["rounds"]
    resp = C.create_string_buffer(128 * k)
    assert lib.bppp_nl_prove_device(h, k, resp, None, None, None, None) != 0          # no transcript
    ctx._ck(lib.bppp_dtr_create(ctx.h, 1, 1 + 2 * k, 0, C.byref(t2)), "bppp_dtr_create")
    ctx._ck(lib.bppp_nl_attach_transcript(h, t2), "attach")
    X, Rr = C.create_string_buffer(64), C.create_string_buffer(64)
    ctx._ck(lib.bppp_nl_round_commit(h, X, Rr), "round_commit")
    assert lib.bppp_nl_prove_device(h, k, resp, None, None, None, None) != 0          # a round was already taken
    comm = sweep.make_comm(ctx, 1, 0)
    assert lib.bppp_nl_prove_sharded(h, comm, k, 1, M, resp, None, None, None, None) != 0
    lib.bppp_nl_destroy(h)
    ctx._ck(lib.bppp_nl_create_gens(gh, ARG_NL, 1, inp["q"], inp["s"], inp["w"], inp["l"], inp["c"], C.byref(h)), "create")
    ctx._ck(lib.bppp_nl_attach_transcript(h, t2), "attach")
    ctx._ck(lib.bppp_nl_set_shard(h, N), "set_shard")                                  # rank 0 must start at element 0
    assert lib.bppp_nl_prove_sharded(h, comm, k, 1, M, resp, None, None, None, None) != 0
    assert lib.bppp_nl_prove_sharded(h, comm, k, k + 1, M, resp, None, None, None, None) != 0
    lib.bppp_nl_destroy(h)
    lib.bppp_comm_destroy(comm)
    lib.bppp_dtr_destroy(t2)
    lib.bppp_gens_destroy(gh)
