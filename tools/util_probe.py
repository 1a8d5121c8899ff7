"""dev tool: CPU / GPU utilisation of one bench step (who is the bottleneck?)"""
import os, sys, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bulletproofspp_b200 as bp
from bench import make_inputs, workload_schema
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = bp.Context(0)
setup = bp.RangeProofSetup(ctx, workload_schema())
ins = make_inputs(B, 0, setup.n_inputs)
def step():
    c, r, f = setup.prove_batch_raw(B, ins[0], ins[1], None, ins[2])
    return setup.verify_batch_raw(B, c, r, f)
step(); step()
util = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=utilization.gpu,clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [util.append(l.strip()) for l in p.stdout], daemon=True).start()
t0, c0 = time.time(), os.times()
for _ in range(2):
    tp = time.time(); c, r, f = setup.prove_batch_raw(B, ins[0], ins[1], None, ins[2]); tv = time.time(); ok = setup.verify_batch_raw(B, c, r, f); te = time.time()
    print("prove %.0f ms verify %.0f ms" % ((tv - tp) * 1e3, (te - tv) * 1e3))
t1, c1 = time.time(), os.times()
p.terminate()
wall = t1 - t0
cpu = (c1.user - c0.user) + (c1.system - c0.system)
print("proofs/s %.0f  wall %.2fs  cpu %.2fs -> %.1f cores busy of %d (sys %.2fs)" % (2 * B / wall, wall, cpu, cpu / wall, os.cpu_count(), c1.system - c0.system))
gu = [int(u.split(",")[0]) for u in util if u]
print("gpu util samples: mean %.0f%% min %d max %d  n=%d" % (sum(gu) / max(len(gu), 1), min(gu), max(gu), len(gu)), util[:3])
