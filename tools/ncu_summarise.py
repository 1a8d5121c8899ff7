"""dev tool: profiles/ncu_summary.json from the raw-page CSVs of the committed `ncu --set full` captures.
   python tools/ncu_summarise.py            (rewrites profiles/ncu_summary.json)
For every kernel the newest capture wins (r2b_* over r2_*).  bench.py reads the summary for `roofline.traffic` and
`pipe_active_ncu` (a profiler cannot run inside the timed bench)."""
import csv, glob, json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def summarise():
    out = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r2*_ncu_full_*.csv"))):
        name = re.sub(r"^r2b?_ncu_full_", "", os.path.basename(path))[:-4]
        if "before" in name:
            continue
        rows = list(csv.reader(open(path)))
        d = dict(zip(rows[0], rows[2]))
        u = dict(zip(rows[0], rows[1]))
        f = lambda k: float(d[k]) if d.get(k) not in (None, "", "no data") else None
        def bytes_of(k):
            v = f(k)
            if v is None:
                return 0.0
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
        dur = f("gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[u["gpu__time_duration.sum"]]
        stalls = {}
        for k in d:
            m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio$", k)
            if m and f(k) is not None and f(k) >= 0.3:
                stalls[m.group(1)] = round(f(k), 2)
        out[name] = {
            "pipe_fmaheavy_active_pct": round(f("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"), 2),
            "issue_active_pct": round(f("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
            "warps_active_pct": round(f("sm__warps_active.avg.pct_of_peak_sustained_active"), 2),
            "dram_bytes_per_launch": int(bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")),
            "duration_us_under_ncu": round(dur, 2),
            "registers": int(f("launch__registers_per_thread")),
            "grid": int(f("launch__grid_size")),
            "block": int(f("launch__block_size")),
            "stalls_per_issue_ge_0.3": stalls,
            "source": "profiles/" + os.path.basename(path),
        }
    return out


if __name__ == "__main__":
    out = summarise()
    with open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w") as fo:
        json.dump(out, fo, indent=1, sort_keys=True)
        fo.write("\n")
    print(json.dumps({k: (v["pipe_fmaheavy_active_pct"], v["duration_us_under_ncu"], v["source"]) for k, v in out.items()}, indent=1))
