"""Per-kernel timing of the device argument on ONE lane (no concurrent streams): runs B synthetic
128by64-shaped NormLinear arguments through bppp_nl_* with the context's CUDA-event profiler on and
prints each kernel's average launch time and algorithmic rate.  Dev tool; not used by tests/bench.

  python tools/kernel_bench.py [B] [N] [M] [rounds] [reps]
"""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofspp_b200 as bp
from bulletproofspp_b200.lib import points_to_bytes, point_to_bytes
from oracle.curve import Secp256k1 as G
from oracle.transcript import get_points

R = G.order
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
M = int(sys.argv[3]) if len(sys.argv) > 3 else 261
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 9
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3


def rnd(tag, n):
    out = bytearray()
    i = 0
    while len(out) < 32 * n:
        blk = hashlib.sha512((tag + str(i)).encode()).digest()
        out += (int.from_bytes(blk[:32], "little") % R).to_bytes(32, "little")
        out += (int.from_bytes(blk[32:], "little") % R).to_bytes(32, "little")
        i += 1
    return bytes(out[:32 * n])


npts = min(1 + N + M, 1500)
pts = get_points(G, "test points", npts)
pts = [pts[i % npts] for i in range(1 + N + M)]
ctx = bp.Context(0)
wide, lo32 = ctx.measure_imad_peak()
g, Gb, Hb = point_to_bytes(pts[0]), points_to_bytes(pts[1:1 + N]), points_to_bytes(pts[1 + N:])
q, s, w, l, c = rnd("q", B), rnd("s", B), rnd("w", B * N), rnd("l", B * M), rnd("c", B * M)
es = [rnd("e%d" % r, B) for r in range(rounds)]


def run():
    arg = bp.NormLinearArgument.from_bytes(ctx, bp.ARG_NL, B, N, M, g, Gb, Hb, q, s, w, l, c)
    for r in range(rounds):
        arg.round_commit_raw()
        arg.round_fold(es[r])
    arg.final()
    arg.close()


run()                                    # warm-up (tables, pools)
ctx.profile_enable(True)
ctx.profile_reset()
t0 = time.time()
for _ in range(reps):
    run()
ctx.sync()
wall = time.time() - t0
rep = ctx.profile_report()
kern = [dict(v, name=n) for n, v in rep["kernels"].items() if v["launches"]]
tot = sum(k["ms"] for k in kern)
print("B=%d N=%d M=%d rounds=%d reps=%d  wall %.1f ms/run, kernel time %.1f ms/run, IMAD.WIDE peak %.2f T/s" % (
    B, N, M, rounds, reps, wall / reps * 1e3, tot / reps, wide / 1e12))
for k in sorted(kern, key=lambda k: -k["ms"]):
    line = "  %-20s launches %5d  avg %8.3f ms  share %5.1f%%" % (k["name"], k["launches"], k["ms"] / max(1, k["launches"]), 100 * k["ms"] / tot)
    if k.get("work"):
        rate = k["work"] / (k["ms"] * 1e-3)
        if k["name"] in ("k_msm_gens", "k_msm_bucket", "k_pair_fold", "k_fb_msm"):
            line += "  %.2f TIMAD/s (%.0f%% of peak)" % (rate / 1e12, 100 * rate / wide)
        else:
            line += "  %.1f GB/s" % (rate / 1e9)
    print(line)
print(json.dumps({"kernel_ms_per_argument": tot / reps / B}))
