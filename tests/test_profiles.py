"""profiles/ncu_summary.json (the counters bench.py quotes as `roofline.traffic` and `pipe_active_ncu`) must be what
tools/ncu_summarise.py extracts from the committed raw-page CSVs of the ncu captures -- no hand-edited numbers."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ncu_summary_matches_the_committed_captures():
    spec = importlib.util.spec_from_file_location("ncu_summarise", os.path.join(ROOT, "tools", "ncu_summarise.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
        committed = json.load(f)
    assert mod.summarise() == committed
    for name in ("k_msm_lut", "k_fold_dots", "k_pair_fold", "k_tr_squeeze_coop"):
        assert os.path.exists(os.path.join(ROOT, committed[name]["source"]))
    assert committed["k_msm_lut"]["dram_bytes_per_launch"] > 0 and 0 < committed["k_msm_lut"]["pipe_fmaheavy_active_pct"] < 100
