"""CLI of the synthetic norm-argument sweep (bulletproofspp_b200/sweep.py):

    python tools/sweep.py [e ...]      one JSON line per size N = 2^e
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bulletproofspp_b200 as bp
from bulletproofspp_b200 import sweep

if __name__ == "__main__":
    ctx = bp.Context(0)
    sizes = [int(a) for a in sys.argv[1:]] or [10, 12, 14]
    wide, _ = ctx.measure_imad_peak()
    hbm = 6546.6
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    res = sweep.run(ctx, sizes, wide, hbm)
    print(json.dumps({k: v for k, v in res.items() if k != "sizes"}), flush=True)
    for r in res["sizes"]:
        print(json.dumps(r), flush=True)
