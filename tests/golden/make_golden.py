"""Regenerates tests/golden/*.json with the CPU oracle (oracle/).  PARITY UNPINNED: the reference
ships no expected outputs, so these vectors pin the oracle against itself over time (and the CUDA
path against the oracle at full size without re-running the slow Python EC code); transcript
policy = PrefixedP, root policy = exp.  Usage: python tests/golden/make_golden.py [names...]"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from example_configs import EXAMPLES, batched  # noqa: E402
from oracle.curve import Secp256k1 as G  # noqa: E402
from oracle.rangeproof import load_schema, load_witness, prove, verify  # noqa: E402
from oracle.transcript import ZKPT  # noqa: E402

hx = lambda v: "%064x" % v
pt = lambda p: None if p is None else [hx(p[0]), hx(p[1])]


def make(name, schema, wit, seed=None):
    t0 = time.time()
    if seed is not None:
        schema = dict(schema, randomSeed=seed)
    setup = load_schema(schema, G)
    w = load_witness(setup, wit)
    trace = {}
    proof = prove(setup, ZKPT(G, setup.random_seed), w, trace)
    ok = verify(setup, ZKPT(G, None), proof)
    assert ok
    fin = proof["opening"].vec.get_witness()
    out = dict(name=name, arg=setup.arg, nrm_len=setup.nrm_len, lin_len=setup.lin_len, rounds=len(proof["responses"]),
               random_seed=setup.random_seed,
               challenges={k: hx(v) for k, v in trace["ch"].items()},
               round_challenges=[hx(r["e"]) for r in trace["rounds"]],
               round_scalars=[[hx(r["sX"]), hx(r["sR"])] for r in trace["rounds"]],
               coms=[pt(p) for p in proof["coms"]],
               responses=[[pt(x), pt(r)] for x, r in proof["responses"]],
               final_scalar=hx(proof["opening"].s), finals=[hx(v) for v in fin], verifies=ok)
    with open(os.path.join(HERE, name.replace("#", "_b") + ".json"), "w") as f:
        json.dump(out, f, indent=0)
    print("%s: N=%d M=%d rounds=%d  %.1fs" % (name, setup.nrm_len, setup.lin_len, out["rounds"], time.time() - t0))


if __name__ == "__main__":
    names = sys.argv[1:] or list(EXAMPLES) + ["128by64#1"]
    for n in names:
        if "#" in n:
            base, b = n.split("#")
            schema, wits, seeds = batched(base, int(b) + 1)
            make(n, schema, wits[int(b)], seeds[int(b)])
        else:
            make(n, *EXAMPLES[n])
