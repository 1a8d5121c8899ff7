"""Synthetic norm-argument sweep: ONE NormLinear argument of N = 2^e norm elements and
M = 6 linear elements, proved (fold mode: scalar folds, generator folds, X/R MSMs every round) and
verified (tensor expansion + one MSM of N+M+2k+2 terms) on the device, with the per-kernel
rooflines.  Witness scalars are SHA-derived; the relation s = |w|^2_q + <c,l> is made to hold so the
verifier accepts.  Generators: the first points of getPoints "test points" up to 4096, beyond that
device-generated multiples h_i*G of the secp256k1 base point (hash-to-curve in Python would take
minutes; parity does not depend on where generators come from).

    python tools/sweep.py [e ...]      prints one JSON line per e
"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import bulletproofspp_b200 as bp
from bulletproofspp_b200 import lib as L
from oracle.curve import Secp256k1 as G
from oracle.field import R
from oracle.transcript import get_points


def scalars(tag, n):
    out = bytearray()
    i = 0
    while len(out) < 32 * n:
        blk = hashlib.sha512(("%s%d" % (tag, i)).encode()).digest()
        out += (int.from_bytes(blk[:32], "little") % R).to_bytes(32, "little")
        out += (int.from_bytes(blk[32:], "little") % R).to_bytes(32, "little")
        i += 1
    return bytes(out[:32 * n])


def generators(ctx, n):
    if n <= 4096:
        return L.points_to_bytes(get_points(G, "test points", n))
    fb = C.c_void_p()
    ctx._ck(ctx.lib.bppp_fb_create(ctx.h, 1, L.point_to_bytes(G.gen), C.byref(fb)), "bppp_fb_create")
    out = C.create_string_buffer(64 * n)
    ctx._ck(ctx.lib.bppp_fb_msm_batch(fb, n, scalars("gen", n), out), "bppp_fb_msm_batch")
    ctx.lib.bppp_fb_destroy(fb)
    return out.raw[:64 * n]


def run(ctx, e, M=6, verify=True, profile=True):
    N = 1 << e
    k = e - 2                                             # NormArgument.hs:165-178 -> final (4, 1)
    t0 = time.time()
    pts = generators(ctx, 1 + N + M)
    g, Gb, Hb = pts[:64], pts[64:64 * (1 + N)], pts[64 * (1 + N):]
    q = scalars("q%d" % e, 1)
    w, l, c = scalars("w%d" % e, N), scalars("l%d" % e, M), scalars("c%d" % e, M)
    qi = L.le_to_int(q)
    wi, li, ci = L.bytes_to_ints(w), L.bytes_to_ints(l), L.bytes_to_ints(c)
    q2 = qi * qi % R
    acc, wt = 0, q2
    for x in wi:                                          # |w|^2_q = sum (q^2)^(i+1) w_i^2
        acc = (acc + wt * x % R * x) % R
        wt = wt * q2 % R
    s0 = (acc + sum(a * b for a, b in zip(ci, li))) % R
    t_setup = time.time() - t0
    if profile:
        ctx.profile_enable(True)
        ctx.profile_reset()
    t0 = time.time()
    arg = bp.NormLinearArgument.from_bytes(ctx, bp.ARG_NL, 1, N, M, g, Gb, Hb, q, L.int_to_le(s0), w, l, c)
    t_create = time.time() - t0
    es, xr = [], []
    per_round = []

    def snap():
        kk = ctx.profile_report()["kernels"] if profile else {}
        return {n: (v["ms"], v["work"]) for n, v in kk.items()}
    t0 = time.time()
    round_ms = []
    for r in range(k):
        tc = time.time()
        X, Rr = arg.round_commit_raw()
        tc = time.time() - tc
        ev = int.from_bytes(hashlib.sha256(X + Rr + bytes([r])).digest(), "big") % R
        es.insert(0, ev)
        xr.insert(0, (L.bytes_to_point(X), L.bytes_to_point(Rr)))
        before = snap() if profile and r < 4 else None
        tf = time.time()
        arg.round_fold(L.int_to_le(ev))
        round_ms.append((round(tc * 1e3, 3), round((time.time() - tf) * 1e3, 3)))
        if before is not None:
            after = snap()
            d = {n: (after[n][0] - before.get(n, (0, 0))[0], after[n][1] - before.get(n, (0, 0))[1]) for n in after}
            per_round.append(d)
    s, fw, fl = arg.final()
    t_prove = time.time() - t0
    arg.close()
    out = {"e": e, "N": N, "M": M, "rounds": k, "final": [len(fw[0]), len(fl[0])], "setup_s": round(t_setup, 2),
           "create_s": round(t_create, 3), "prove_s": round(t_prove, 4), "round_ms_commit_fold": round_ms}
    if verify:
        # initCom = the commitment C0 itself (public vector 0): C0 = s0*g + <w,G> + <l,H>
        sc = L.int_to_le(s0) + w + l
        C0 = C.create_string_buffer(64)
        ctx._ck(ctx.lib.bppp_msm(ctx.h, 1 + N + M, sc, pts, C0), "bppp_msm")
        t0 = time.time()
        ok = ctx.nl_verify(bp.ARG_NL, L.bytes_to_point(g), L.bytes_to_points(Gb), L.bytes_to_points(Hb), [qi], [0],
                           [[0] * N], [ci], [es], [xr], fw, fl, [[(1, L.bytes_to_point(C0.raw))]]) if N <= 1 << 14 else None
        if ok is None:                                    # large N: avoid Python list marshalling
            okv = (C.c_int * 1)()
            ctx._ck(ctx.lib.bppp_nl_verify(
                ctx.h, bp.ARG_NL, 1, N, M, k, g, Gb, Hb, q, bytes(32), bytes(32 * N), c, L.ints_to_bytes(es),
                b"".join(L.point_to_bytes(x) + L.point_to_bytes(r_) for x, r_ in xr), len(fw[0]), len(fl[0]),
                L.ints_to_bytes(fw[0]), L.ints_to_bytes(fl[0]), 1, L.int_to_le(1), C0.raw, okv), "bppp_nl_verify")
            ok = [bool(okv[0])]
        out["verify_s"] = round(time.time() - t0, 4)
        out["verifies"] = ok[0]
    if profile:
        rep = ctx.profile_report()
        ctx.profile_enable(False)
        kern = rep["kernels"]
        out["kernels_ms"] = {n: round(v["ms"], 3) for n, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}
        if "k_fold_dots" in kern:
            kf = kern["k_fold_dots"]
            out["fold_scalar_GBps"] = round(kf["work"] / (kf["ms"] * 1e-3) / 1e9, 2)
        if "k_pair_fold" in kern:
            kp = kern["k_pair_fold"]
            # generator fold: 96 bytes per input pair element (64 in + 32 out per point pair -> 96*(N_k+M_k))
            out["fold_points_TIMADps"] = round(kp["work"] / (kp["ms"] * 1e-3) / 1e12, 3)
        out["proofs_per_s"] = round(1.0 / (t_prove + out.get("verify_s", 0)), 3)
        # the first fold is the one that streams the full-length vectors: bytes / CUDA-event time
        if per_round:
            d = per_round[0]
            if "k_fold_dots" in d and d["k_fold_dots"][0] > 0:
                out["fold1_scalar_GBps"] = round(d["k_fold_dots"][1] / (d["k_fold_dots"][0] * 1e-3) / 1e9, 1)
                out["fold1_scalar_ms"] = round(d["k_fold_dots"][0], 4)
            if "k_pair_fold" in d and d["k_pair_fold"][0] > 0:
                out["fold1_points_TIMADps"] = round(d["k_pair_fold"][1] / (d["k_pair_fold"][0] * 1e-3) / 1e12, 3)
                out["fold1_points_GBps"] = round(96.0 * (N + M) / (d["k_pair_fold"][0] * 1e-3) / 1e9, 2)
                out["fold1_points_ms"] = round(d["k_pair_fold"][0], 3)
    return out


if __name__ == "__main__":
    ctx = bp.Context(0)
    prof = not os.environ.get("SWEEP_NOPROFILE")
    for e in [int(a) for a in sys.argv[1:]] or [10, 12, 14]:
        print(json.dumps(run(ctx, e, profile=prof)), flush=True)
