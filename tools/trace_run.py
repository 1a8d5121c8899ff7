import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import bulletproofspp_b200 as bp
from bench import make_inputs, workload_schema
B = 4096
ctx = bp.Context(0)
setup = bp.RangeProofSetup(ctx, workload_schema())
ins = make_inputs(B, 0, setup.n_inputs)
c, r, f = setup.prove_batch_raw(B, ins[0], ins[1], None, ins[2]); setup.verify_batch_raw(B, c, r, f)
sys.stderr.write("=== MARK\n"); sys.stderr.flush()
t=time.time(); c, r, f = setup.prove_batch_raw(B, ins[0], ins[1], None, ins[2]); t1=time.time()
sys.stderr.write("=== MARK2\n"); sys.stderr.flush()
ok = setup.verify_batch_raw(B, c, r, f); t2=time.time()
print("prove %.0f verify %.0f" % ((t1-t)*1e3, (t2-t1)*1e3))
