"""Synthetic norm-argument sweep (BASELINE.json config 5) as a measured workload: ONE
NormLinear argument of N = 2^e norm elements and M = 6 linear elements over the first 1 + N + M generators
of getPoints "test points", proved (proveBPM with the reference transcript, bppp_nl_prove: scalar folds,
generator folds and the X / R commitments of every round) and verified (challenge replay + tensor
expansion + one collapsed MSM of N + M + 2k + 2 terms, bppp_nl_verify_gens) through the C ABI.

    run(ctx, sizes, imad_wide_per_s, hbm_gbs) -> dict for bench.py's `norm_arg_sweep`
"""
import ctypes as C
import time

from . import workloads as W
from .lib import ARG_NL

MSM_KERNELS = ("k_pip_sort", "k_pip_accum", "k_pip_merge", "k_pip_reduce", "k_pip_horner", "k_msm_gens", "k_msm_gens_reduce",
               "k_jac_sum", "k_msm_bucket", "k_msm_finish", "k_gt_build")


def _prove(ctx, gens, inp, C0):
    lib, k = ctx.lib, inp["rounds"]
    h = C.c_void_p()
    ctx._ck(lib.bppp_nl_create_gens(gens, ARG_NL, 1, inp["q"], inp["s"], inp["w"], inp["l"], inp["c"], C.byref(h)), "bppp_nl_create_gens")
    resp, es = C.create_string_buffer(128 * k), C.create_string_buffer(32 * k)
    s_out, fw, fl = C.create_string_buffer(32), C.create_string_buffer(32 * 4), C.create_string_buffer(32)
    ctx.sync()
    ctx.timer_start()
    t0 = time.time()
    ctx._ck(lib.bppp_nl_prove(h, 1, 0, 1, C0, k, resp, es), "bppp_nl_prove")
    ctx._ck(lib.bppp_nl_final(h, s_out, fw, fl), "bppp_nl_final")
    ms = ctx.timer_stop()
    wall = time.time() - t0
    lib.bppp_nl_destroy(h)
    return dict(ms=ms, wall_ms=1e3 * wall, resp=resp.raw[:128 * k], es=es.raw[:32 * k], fw=fw.raw[:128], fl=fl.raw[:32])


def _prove_device(ctx, gens, inp, C0):
    """the same proof with the whole round loop on the device (bppp_nl_prove_device): the transcript (the initial
    commitment, then every round's X, R) is rendered and hashed there, rationalReduceScalar and the fold factors
    run in k_round_* -- one stream of launches, one synchronisation"""
    lib, k = ctx.lib, inp["rounds"]
    h, t = C.c_void_p(), C.c_void_p()
    ctx._ck(lib.bppp_nl_create_gens(gens, ARG_NL, 1, inp["q"], inp["s"], inp["w"], inp["l"], inp["c"], C.byref(h)), "bppp_nl_create_gens")
    ctx._ck(lib.bppp_dtr_create(ctx.h, 1, 1 + 2 * k, 0, C.byref(t)), "bppp_dtr_create")
    resp, es = C.create_string_buffer(128 * k), C.create_string_buffer(32 * k)
    s_out, fw, fl = C.create_string_buffer(32), C.create_string_buffer(32 * 4), C.create_string_buffer(32)
    ctx.sync()
    ctx.timer_start()
    t0 = time.time()
    ctx._ck(lib.bppp_dtr_absorb(t, C0, 1, 1), "bppp_dtr_absorb")
    ctx._ck(lib.bppp_nl_attach_transcript(h, t), "bppp_nl_attach_transcript")
    ctx._ck(lib.bppp_nl_prove_device(h, k, resp, es, s_out, fw, fl), "bppp_nl_prove_device")
    ms = ctx.timer_stop()
    wall = time.time() - t0
    lib.bppp_nl_destroy(h)
    lib.bppp_dtr_destroy(t)
    return dict(ms=ms, wall_ms=1e3 * wall, resp=resp.raw[:128 * k], es=es.raw[:32 * k], fw=fw.raw[:128], fl=fl.raw[:32])


def sharded_local_rounds(e, world):
    """fold locally until the gathered argument has at most 4096 norm elements (then it re-bases to tensor mode)"""
    ln = e - (world.bit_length() - 1)                         # log2 of the slice length
    return max(0, min(ln, e - 12))


def make_comm(ctx, world, rank, broadcast=None):
    """bppp_comm over NCCL: rank 0 draws the id, `broadcast(bytes_or_None) -> bytes` hands it to every rank"""
    lib = ctx.lib
    comm = C.c_void_p()
    if world == 1:
        ctx._ck(lib.bppp_comm_create(ctx.h, 1, 0, None, C.byref(comm)), "bppp_comm_create")
        return comm
    try:                                                      # the NCCL torch.distributed itself uses
        import nvidia.nccl
        import os
        path = os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so.2").encode()
    except Exception:
        path = None
    if lib.bppp_comm_load(path):
        raise RuntimeError("bppp_comm_load: " + lib.bppp_comm_last_error().decode())
    idb = C.create_string_buffer(128)
    if rank == 0 and lib.bppp_comm_unique_id(idb):
        raise RuntimeError("bppp_comm_unique_id: " + lib.bppp_comm_last_error().decode())
    raw = broadcast(idb.raw[:128] if rank == 0 else None)
    ctx._ck(lib.bppp_comm_create(ctx.h, world, rank, raw, C.byref(comm)), "bppp_comm_create")
    return comm


def _prove_sharded(ctx, comm, rank, world, inp, points, C0, local_rounds, barrier=None):
    """one argument over `world` ranks (bppp_nl_prove_sharded): this rank's contiguous slice of the norm vector and its
    generators, the linear part on rank 0; returns the same fields as _prove_device (identical on every rank)"""
    lib, k, N, M = ctx.lib, inp["rounds"], inp["N"], inp["M"]
    ln = N // world
    lo = rank * ln
    Mr = M if rank == 0 else 0
    gens, h, t = C.c_void_p(), C.c_void_p(), C.c_void_p()
    Hs = points[64 * (1 + N):64 * (1 + N + M)] if Mr else None
    ctx._ck(lib.bppp_gens_create(ctx.h, ln, Mr, points[:64], points[64 * (1 + lo):64 * (1 + lo + ln)], Hs, C.byref(gens)), "bppp_gens_create")
    ctx._ck(lib.bppp_nl_create_gens(gens, ARG_NL, 1, inp["q"], inp["s"], inp["w"][32 * lo:32 * (lo + ln)],
                                    inp["l"] if Mr else None, inp["c"] if Mr else None, C.byref(h)), "bppp_nl_create_gens")
    ctx._ck(lib.bppp_nl_set_shard(h, lo), "bppp_nl_set_shard")
    ctx._ck(lib.bppp_dtr_create(ctx.h, 1, 1 + 2 * k, 0, C.byref(t)), "bppp_dtr_create")
    resp, es = C.create_string_buffer(128 * k), C.create_string_buffer(32 * k)
    s_out, fw, fl = C.create_string_buffer(32), C.create_string_buffer(32 * 4), C.create_string_buffer(32)
    ctx.sync()
    if barrier:
        barrier()
    ctx.timer_start()
    t0 = time.time()
    ctx._ck(lib.bppp_dtr_absorb(t, C0, 1, 1), "bppp_dtr_absorb")
    ctx._ck(lib.bppp_nl_attach_transcript(h, t), "bppp_nl_attach_transcript")
    ctx._ck(lib.bppp_nl_prove_sharded(h, comm, k, local_rounds, M, resp, es, s_out, fw, fl), "bppp_nl_prove_sharded")
    ms = ctx.timer_stop()
    wall = time.time() - t0
    lib.bppp_nl_destroy(h)
    lib.bppp_dtr_destroy(t)
    lib.bppp_gens_destroy(gens)
    return dict(ms=ms, wall_ms=1e3 * wall, resp=resp.raw[:128 * k], es=es.raw[:32 * k], fw=fw.raw[:128], fl=fl.raw[:32], s=s_out.raw[:32])


def _verify(ctx, gens, inp, C0, pr):
    lib, k, N = ctx.lib, inp["rounds"], inp["N"]
    es = C.create_string_buffer(32 * k)
    ok = (C.c_int * 1)()
    zeros = bytes(32 * N)
    ctx.sync()
    ctx.timer_start()
    t0 = time.time()
    ctx._ck(lib.bppp_nl_challenges(1, 0, 1, C0, k, pr["resp"], es), "bppp_nl_challenges")
    ctx._ck(lib.bppp_nl_verify_gens(gens, ARG_NL, 1, k, inp["q"], bytes(32), zeros, inp["c"], es, pr["resp"], 4, 1, pr["fw"], pr["fl"],
                                    1, (1).to_bytes(32, "little"), C0, ok), "bppp_nl_verify_gens")
    ms = ctx.timer_stop()
    wall = time.time() - t0
    assert es.raw[:32 * k] == pr["es"], "verifier's challenges differ from the prover's"
    return dict(ms=ms, wall_ms=1e3 * wall, ok=bool(ok[0]))


def run_one(ctx, e, points, imad_wide, hbm_gbs, reps=2, M=6):
    lib = ctx.lib
    N = 1 << e
    P0 = 1 + N + M
    t0 = time.time()
    inp = W.sweep_inputs(ctx, e, M)
    gens = C.c_void_p()
    ctx._ck(lib.bppp_gens_create(ctx.h, N, M, points[:64], points[64:64 * (1 + N)], points[64 * (1 + N):64 * P0], C.byref(gens)),
            "bppp_gens_create")
    # the initial commitment C0 = s g + <w, G> + <l, H> (public vector 0): the one commitment in the transcript
    C0b = C.create_string_buffer(64)
    ctx._ck(lib.bppp_gens_msm_batch(gens, 1, P0, inp["s"] + inp["w"] + inp["l"], C0b), "bppp_gens_msm_batch")
    C0 = C0b.raw[:64]
    setup_s = time.time() - t0
    _prove(ctx, gens, inp, C0)                                   # warm-up (pools, lazily loaded kernels)
    hosts = [_prove(ctx, gens, inp, C0) for _ in range(reps)]     # reference arrangement: transcript + round constants on the host
    _prove_device(ctx, gens, inp, C0)
    proves = [_prove_device(ctx, gens, inp, C0) for _ in range(reps)]
    assert all(p["resp"] == hosts[0]["resp"] and p["fw"] == hosts[0]["fw"] and p["fl"] == hosts[0]["fl"] and p["es"] == hosts[0]["es"]
               for p in proves + hosts), "device round loop and host sequencing disagree"
    pr = min(proves, key=lambda p: p["ms"])
    ph = min(hosts, key=lambda p: p["ms"])
    _verify(ctx, gens, inp, C0, pr)
    verifies = [_verify(ctx, gens, inp, C0, pr) for _ in range(reps)]
    vr = min(verifies, key=lambda v: v["ms"])
    # a tampered final scalar must be rejected
    bad = dict(pr, fw=bytes([pr["fw"][0] ^ 1]) + pr["fw"][1:])
    rejected = not _verify(ctx, gens, inp, C0, bad)["ok"]
    # one more prove + verify with a CUDA-event pair around every launch: per-kernel times and algorithmic work
    ctx.profile_enable(True)
    ctx.profile_reset()
    pp = _prove_device(ctx, gens, inp, C0)
    rep_p = ctx.profile_report()["kernels"]
    ctx.profile_reset()
    _verify(ctx, gens, inp, C0, pp)
    rep_v = ctx.profile_report()["kernels"]
    ctx.profile_enable(False)
    lib.bppp_gens_destroy(gens)
    out = {"e": e, "N": N, "M": M, "rounds": inp["rounds"], "final": [4, 1], "setup_s": round(setup_s, 3),
           "prove_ms": round(pr["ms"], 3), "prove_wall_ms": round(pr["wall_ms"], 3),
           "prove_host_sequenced_ms": round(ph["ms"], 3), "verify_ms": round(vr["ms"], 3),
           "verify_wall_ms": round(vr["wall_ms"], 3), "verifies": all(v["ok"] for v in verifies), "rejects_tampered": rejected,
           "proofs_per_s": round(1e3 / (pr["ms"] + vr["ms"]), 3)}
    kms = {n: round(v["ms"], 3) for n, v in sorted(rep_p.items(), key=lambda kv: -kv[1]["ms"])}
    out["prove_kernels_ms"] = kms
    out["verify_kernels_ms"] = {n: round(v["ms"], 3) for n, v in sorted(rep_v.items(), key=lambda kv: -kv[1]["ms"])}
    roof = {}
    if "k_fold_dots" in rep_p and rep_p["k_fold_dots"]["ms"] > 0:
        kf = rep_p["k_fold_dots"]
        a = kf["work"] / (kf["ms"] * 1e-3) / 1e9
        roof["k_fold_dots"] = {"bound": "hbm", "unit": "GB/s", "achieved": round(a, 2), "peak": hbm_gbs, "frac": round(a / hbm_gbs, 4),
                               "ms": round(kf["ms"], 4), "launches": kf["launches"]}
        if kf["top_ms"] > 0:
            a1 = kf["top_work"] / (kf["top_ms"] * 1e-3) / 1e9
            roof["k_fold_dots"]["largest_launch"] = {"achieved": round(a1, 2), "frac": round(a1 / hbm_gbs, 4), "ms": round(kf["top_ms"], 4),
                                                     "bytes": kf["top_work"]}
        out["fold_scalar_GBps"] = roof["k_fold_dots"].get("largest_launch", roof["k_fold_dots"])["achieved"]
    if "k_pair_fold" in rep_p and rep_p["k_pair_fold"]["ms"] > 0:
        kp = rep_p["k_pair_fold"]
        a = kp["work"] / (kp["ms"] * 1e-3) / 1e12
        # issued IMAD.WIDE per folded point: 129 XYZZ doublings (6M + 3S) + ~65 mixed additions (8M + 2S; the joint sparse
        # form of a 129-bit pair is half dense) with fq::mul = 72 and fq::sqr = 43 IMAD.WIDE in this build's SASS; the
        # reference-unit count (half-length Shamir at 136 IMAD per multiplication) is 2.3x that
        issued = a * (129 * (6 * 72 + 3 * 43) + 65 * (8 * 72 + 2 * 43)) / (0.5 * 3.9e3 * 136.0)
        roof["k_pair_fold"] = {"bound": "imad", "unit": "TIMAD/s", "achieved": round(issued, 3), "peak": round(imad_wide / 1e12, 3),
                               "frac": round(issued / (imad_wide / 1e12), 4), "frac_alg": round(a / (imad_wide / 1e12), 4),
                               "achieved_alg": round(a, 3), "ms": round(kp["ms"], 3), "launches": kp["launches"],
                               "note": "frac = issued IMAD.WIDE (estimated from the schedule) / peak; frac_alg = reference-unit IMADs / peak"}
        out["fold_points_TIMADps"] = roof["k_pair_fold"]["achieved"]
        # bytes_points(k) = 96 (N_k + M_k) per round; all rounds ~ 192 (N + M)
        out["fold_points_GBps"] = round(192.0 * (N + M) / (kp["ms"] * 1e-3) / 1e9, 2)
    for side, rep in (("prove", rep_p), ("verify", rep_v)):
        ms = sum(rep[n]["ms"] for n in MSM_KERNELS if n in rep)
        work = sum(rep[n]["work"] for n in MSM_KERNELS if n in rep)
        if ms > 0 and work > 0:
            a = work / (ms * 1e-3) / 1e12
            roof["msm_" + side] = {"bound": "imad", "unit": "TIMAD/s", "achieved": round(a, 3), "peak": round(imad_wide / 1e12, 3),
                                   "frac": round(a / (imad_wide / 1e12), 4), "ms": round(ms, 3),
                                   "kernels": [n for n in MSM_KERNELS if n in rep]}
    out["rooflines"] = roof
    return out


def run_sharded(ctx, e, rank, world, dist=None, reps=3, M=6):
    """N = 2^e over `world` ranks (strong scaling): every rank derives the same generators and witness, proves its
    slice with bppp_nl_prove_sharded and asserts the result equals the unsharded bppp_nl_prove_device proof.
    `dist` = torch.distributed (initialised, NCCL) or None for one rank.  Times are the max over ranks."""
    import hashlib
    lib = ctx.lib
    N = 1 << e
    P0 = 1 + N + M
    points = W.sweep_generators(ctx, P0)
    inp = W.sweep_inputs(ctx, e, M)
    gens = C.c_void_p()
    ctx._ck(lib.bppp_gens_create(ctx.h, N, M, points[:64], points[64:64 * (1 + N)], points[64 * (1 + N):64 * P0], C.byref(gens)), "bppp_gens_create")
    C0b = C.create_string_buffer(64)
    ctx._ck(lib.bppp_gens_msm_batch(gens, 1, P0, inp["s"] + inp["w"] + inp["l"], C0b), "bppp_gens_msm_batch")
    C0 = C0b.raw[:64]
    _prove_device(ctx, gens, inp, C0)
    singles = [_prove_device(ctx, gens, inp, C0) for _ in range(reps)]
    lib.bppp_gens_destroy(gens)
    one = min(singles, key=lambda p: p["ms"])

    def bcast(raw):
        obj = [raw]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    comm = make_comm(ctx, world, rank, bcast if dist else None)
    lr = sharded_local_rounds(e, world)
    barrier = dist.barrier if dist else None
    _prove_sharded(ctx, comm, rank, world, inp, points, C0, lr, barrier)
    runs = [_prove_sharded(ctx, comm, rank, world, inp, points, C0, lr, barrier) for _ in range(reps)]
    lib.bppp_comm_destroy(comm)
    same = all(r[k] == one[k] for r in runs for k in ("resp", "es", "fw", "fl"))
    ms = [r["ms"] for r in runs]
    if dist:
        import torch
        t = torch.tensor(ms + [0.0 if same else 1.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, same = t[:-1].tolist(), t[-1].item() == 0.0
    assert same, "sharded proof differs from the unsharded one"
    best = min(ms)
    return {"workload": "one NormLinear argument, N = 2^%d, M = %d, sharded over the GPUs (bppp_nl_prove_sharded: NCCL all-gather of "
                        "256 B per rank and round inside the library)" % (e, M),
            "e": e, "N": N, "M": M, "rounds": inp["rounds"], "n_gpus": world, "scaling": "strong", "local_rounds": lr,
            "gathered_norm_length": ((N // world) >> lr) * world,
            "prove_ms": round(best, 3), "prove_ms_all": [round(x, 3) for x in ms], "single_gpu_prove_ms": round(one["ms"], 3),
            "speedup_vs_single_gpu": round(one["ms"] / best, 3), "bit_identical_to_unsharded": same,
            "proof_checksum": hashlib.sha256(one["resp"] + one["fw"] + one["fl"]).hexdigest()[:16]}


def run(ctx, sizes, imad_wide, hbm_gbs, reps=2):
    sizes = sorted(sizes)
    t0 = time.time()
    points = W.sweep_generators(ctx, 1 + (1 << sizes[-1]) + 6)          # the lists of the smaller sizes are prefixes
    gen_s = time.time() - t0
    res = {"generators": {"count": 1 + (1 << sizes[-1]) + 6, "derive_s": round(gen_s, 3),
                          "how": "getPoints \"test points\" (app/Main.hs:68-72) on the device, bppp_get_points"},
           "note": "one NormLinear argument per size, N = 2^e, M = 6; prove = bppp_dtr_absorb (initial commitment) + bppp_nl_prove_device "
                   "(round loop, transcript and round constants on the device, one synchronisation); prove_host_sequenced = the same proof "
                   "(asserted bit-identical) by bppp_nl_prove + bppp_nl_final with the transcript and round constants on the host, two "
                   "synchronisations per round; verify = bppp_nl_challenges + bppp_nl_verify_gens; *_ms = CUDA events on the library's stream around the whole "
                   "call sequence (witness / generators resident), best of %d; rooflines from one extra profiled pass; msm_* = algorithmic "
                   "IMADs of SURVEY 8(d) over all MSM kernels of the side" % reps,
           "sizes": []}
    for e in sizes:
        res["sizes"].append(run_one(ctx, e, points, imad_wide, hbm_gbs, reps))
    return res
