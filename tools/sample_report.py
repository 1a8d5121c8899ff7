"""dev tool: flat profile from a BPPP_SAMPLE=file dump (see rp_host.cpp): resolves program counters inside
libbppp_b200.so against `nm` and prints the top symbols / libraries.   python tools/sample_report.py file [top]"""
import bisect, collections, subprocess, sys
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
libs, offs = collections.Counter(), collections.defaultdict(list)
total = 0
for line in open(path):
    lib, off, sym = line.split(None, 2)
    off, _, caller = off.partition("@")          # "pc@return-address" (the caller is 0 outside the library)
    total += 1
    libs[lib] += 1
    offs[lib].append((int(off, 16), sym.strip(), int(caller or "0", 16)))
print("samples", total)
for lib, n in libs.most_common(8):
    print("%6.1f%%  %s" % (100 * n / total, lib))
for lib in offs:
    if "libbppp_b200" not in lib:
        c = collections.Counter(s for _, s, _ in offs[lib])
        for s, n in c.most_common(6):
            if n / total > 0.01:
                print("   %5.1f%%  %s : %s" % (100 * n / total, lib.split("/")[-1], s))
        continue
    out = subprocess.run(["nm", "-C", "--defined-only", lib], capture_output=True, text=True).stdout
    syms = sorted((int(a, 16), name) for a, t, name in (l.split(None, 2) for l in out.splitlines() if len(l.split(None, 2)) == 3) if t.lower() in "tw")
    addrs = [a for a, _ in syms]
    c, pairs = collections.Counter(), collections.Counter()
    name_of = lambda off: (lambda i: syms[i][1].strip() if i >= 0 else "?")(bisect.bisect_right(addrs, off) - 1)
    for off, _, caller in offs[lib]:
        c[name_of(off)] += 1
        if caller:
            pairs[(name_of(off)[:60], name_of(caller)[:90])] += 1
    for s, n in c.most_common(top):
        print("   %5.1f%%  %s" % (100 * n / total, s[:150]))
    if pairs:
        print("callee <- caller (frame-pointer build):")
        for (a, b), n in pairs.most_common(top):
            print("   %5.1f%%  %s <- %s" % (100 * n / total, a, b))
