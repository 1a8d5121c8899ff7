"""dev tool: ASCII Gantt of `BPPP_TRACE=1` output (stderr lines "[bppp trace] what lane=i name t0 t1", ms).
   python tools/trace_view.py trace.err [ms_per_column]"""
import sys, collections
col = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
evs = []
for line in open(sys.argv[1]):
    if not line.startswith("[bppp trace]"):
        continue
    _, _, what, lane, name, t0, t1 = line.split()
    evs.append((int(lane.split("=")[1]), name, float(t0), float(t1)))
if not evs:
    sys.exit("no trace lines")
# split into calls: a gap is where the dump happened; simply cluster by time (prove / verify calls print separately)
evs.sort(key=lambda e: e[2])
T0 = evs[0][2]
sym = lambda n: ("h" if n.startswith("host") or n == "verify_host" else "r" if n == "round_hash" else
                 "C" if n == "nl_commit" else "F" if n == "nl_fold" else "M" if n.startswith("msm") else
                 "V" if n == "nl_verify" else "c" if n == "nl_create" else "f")
lanes = sorted({e[0] for e in evs})
end = max(e[3] for e in evs) - T0
ncol = int(end / col) + 1
tot = collections.Counter()
for ln in lanes:
    row = [" "] * ncol
    for (l, n, a, b) in evs:
        if l != ln:
            continue
        tot[sym(n)] += b - a
        for c in range(int((a - T0) / col), int((b - T0) / col) + 1):
            if c < ncol:
                row[c] = sym(n)
    print("%2d |%s|" % (ln, "".join(row)))
print("columns of %.0f ms; total %.0f ms; h=host phase r=round hash C=nl_commit F=nl_fold M=range-proof msm V=verify msm c/f=create/final" % (col, end))
print("lane-time by kind (ms):", {k: round(v) for k, v in tot.items()})
