"""dev tool: where the time of ONE 128by64 proof goes (batch = 1, the drop-in seams' size).
   python tools/latency_probe.py [--lut-gb 0]     (GPU box)
Prints wall time of prove and verify (median of 5) and the per-kernel CUDA-event times of one prove + verify."""
import argparse, os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BPPP_LANES", "1")
import bench
import bulletproofspp_b200 as bp

ap = argparse.ArgumentParser()
ap.add_argument("--lut-gb", type=float, default=0)
ap.add_argument("--batch", type=int, default=1)
a = ap.parse_args()
ctx = bp.Context(0)
st = bp.RangeProofSetup(ctx, bench.workload_schema())
st.set_device_transcript(True)
st.set_batch_verify(True)
if a.lut_gb > 0:
    st.enable_lut(a.lut_gb)
B = a.batch
one = bench.make_inputs(B, 0, st.n_inputs)
tp, tv = [], []
for _ in range(6):
    t0 = time.time()
    pr = st.prove_batch_raw(B, one[0], one[1], None, one[2])
    t1 = time.time()
    assert sum(st.verify_batch_raw(B, *pr)) == B
    t2 = time.time()
    tp.append(1e3 * (t1 - t0)); tv.append(1e3 * (t2 - t1))
print("batch %d: prove %.2f ms, verify %.2f ms (median of 5 after one warm-up)" % (B, statistics.median(tp[1:]), statistics.median(tv[1:])))
ctx.profile_enable(True); ctx.profile_reset()
pr = st.prove_batch_raw(B, one[0], one[1], None, one[2])
rp = ctx.profile_report(); ctx.profile_reset()
assert sum(st.verify_batch_raw(B, *pr)) == B
rv = ctx.profile_report()
for name, r in (("prove", rp), ("verify", rv)):
    k = r["kernels"]
    print("%s: %d launches, %.3f ms of kernels" % (name, sum(x["launches"] for x in k.values()), sum(x["ms"] for x in k.values())))
    for kn, x in sorted(k.items(), key=lambda kv: -kv[1]["ms"]):
        print("   %-22s %4d launches %8.3f ms  (%.1f us each)" % (kn, x["launches"], x["ms"], 1e3 * x["ms"] / max(1, x["launches"])))
