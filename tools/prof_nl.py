"""Quick timing of the device NormLinear argument on synthetic witnesses (dev tool)."""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofspp_b200 as bp
from bulletproofspp_b200.lib import points_to_bytes, point_to_bytes
from oracle.curve import Secp256k1 as G
from oracle.transcript import get_points
R = G.order
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
M = int(sys.argv[3]) if len(sys.argv) > 3 else 261
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 9
t0 = time.time()
npts = min(1 + N + M, 2000)
pts = get_points(G, "test points", npts)
pts = [pts[i % npts] for i in range(1 + N + M)]
print("points %.1fs" % (time.time() - t0))
def rnd(tag, n):
    out = bytearray()
    i = 0
    while len(out) < 32 * n:
        blk = hashlib.sha512((tag + str(i)).encode()).digest() * 1
        out += (int.from_bytes(blk[:32], "little") % R).to_bytes(32, "little")
        out += (int.from_bytes(blk[32:], "little") % R).to_bytes(32, "little")
        i += 1
    return bytes(out[:32 * n])
ctx = bp.Context(0)
g, Gb, Hb = point_to_bytes(pts[0]), points_to_bytes(pts[1:1 + N]), points_to_bytes(pts[1 + N:])
q, s, w, l, c = rnd("q", B), rnd("s", B), rnd("w", B * N), rnd("l", B * M), rnd("c", B * M)
for rep in range(2):
    t0 = time.time()
    arg = bp.NormLinearArgument.from_bytes(ctx, bp.ARG_NL, B, N, M, g, Gb, Hb, q, s, w, l, c)
    t1 = time.time()
    tc = tf = 0
    for r in range(rounds):
        a = time.time()
        X, Rr = arg.round_commit_raw()
        b = time.time()
        e = rnd("e%d" % r, B)
        b2 = time.time()
        arg.round_fold(e)
        d = time.time()
        tc += b - a; tf += d - b2
        if rep: print(" round %d lens %s commit %.1f ms fold %.1f ms" % (r, arg.lengths(), (b - a) * 1e3, (d - b2) * 1e3))
    arg.final()
    arg.close()
    print("rep %d: create %.1f ms, commits %.1f ms, folds %.1f ms, total %.1f ms -> %.1f arguments/s (launches %d)" % (
        rep, (t1 - t0) * 1e3, tc * 1e3, tf * 1e3, (time.time() - t0) * 1e3, B / (time.time() - t0), ctx.launch_count()))
