"""dev tool: every reference citation (`src/Foo.hs:12-34`, `app/Main.hs:68`, `README.md:187` ...) in the headers, sources,
oracle and docs must name a file of the reference tree with at least that many lines.
   python tools/check_citations.py [/root/reference]        (exit status 1 when a citation does not resolve)"""
import os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
pat = re.compile(r"((?:src|app|test)/[A-Za-z0-9_/]+\.hs|package\.yaml|README\.md):(\d+)(?:-(\d+))?")
short = re.compile(r"\b([A-Z][A-Za-z]+(?:/[A-Z][A-Za-z]+)*\.hs):(\d+)(?:-(\d+))?")      # `NormArgument.hs:113` without a directory
index = {}
for d, _, fs in os.walk(REF):
    for f in fs:
        if f.endswith(".hs"):
            index.setdefault(f, []).append(os.path.join(d, f))
lines_of = {}
def nlines(p):
    if p not in lines_of:
        with open(p, errors="replace") as f:
            lines_of[p] = sum(1 for _ in f)
    return lines_of[p]
bad = total = 0
for base in ("include", "bulletproofspp_b200", "oracle", "hs", "tests", "tools", "DESIGN.md", "INTEGRATION.md", "README.md", "bench.py", "__graft_entry__.py"):
    paths = [os.path.join(ROOT, base)] if os.path.isfile(os.path.join(ROOT, base)) else [os.path.join(d, f) for d, _, fs in os.walk(os.path.join(ROOT, base)) for f in fs
                                                                                     if f.endswith((".h", ".cu", ".cuh", ".hpp", ".cpp", ".py", ".md", ".hs", ".c", ".sh"))]
    for p in paths:
        if "_build" in p or "_ref" in p or p.endswith("check_citations.py"):
            continue
        text = open(p, errors="replace").read()
        seen = set()
        for m in list(pat.finditer(text)) + list(short.finditer(text)):
            name, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            if (name, a, b) in seen:
                continue
            seen.add((name, a, b))
            if "/" in name and not name[0].isupper() or name in ("package.yaml", "README.md"):
                cands = [os.path.join(REF, name)]
                if name == "README.md" and os.path.relpath(p, ROOT) in ("README.md",):
                    continue
            else:
                cands = index.get(os.path.basename(name), [])
                cands = [c for c in cands if c.endswith(name)]
            total += 1
            ok = any(os.path.exists(c) and nlines(c) >= max(a, b) for c in cands)
            if not ok:
                bad += 1
                print("%s: %s:%d-%d does not resolve (%s)" % (os.path.relpath(p, ROOT), name, a, b, "no such file" if not any(os.path.exists(c) for c in cands) else "file is shorter"))
print("%d citations checked, %d unresolved" % (total, bad))
sys.exit(1 if bad else 0)
