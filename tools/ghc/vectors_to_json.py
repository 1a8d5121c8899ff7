#!/usr/bin/env python
"""Convert the artefacts of dump_vectors.sh into tests/golden/ghc_<example>.json (see README.md).

points.bin   Data.Binary list of WideEncoding: 8-byte big-endian length, then x, y per point, each four
             big-endian Word64 with the FIRST word least significant (app/Main.hs:90-98, src/Encoding.hs:75-86)
proof.bin / commits.bin   kept as hex: the test compares them with the oracle's own encoder byte for byte
ghci.txt     BEGIN-/END- sections printed by dump_vectors.ghci
"""
import json
import os
import re
import sys


def field(b):
    w = [int.from_bytes(b[8 * i:8 * i + 8], "big") for i in range(4)]
    return w[0] + (w[1] << 64) + (w[2] << 128) + (w[3] << 192)


def section(text, name):
    m = re.search(r"BEGIN-%s\n(.*?)END-%s" % (name, name), text, re.S)
    return [l.strip() for l in m.group(1).splitlines() if l.strip()] if m else None


def main(out_dir):
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    ghci = open(os.path.join(out_dir, "ghci.txt")).read() if os.path.exists(os.path.join(out_dir, "ghci.txt")) else ""
    common = {}
    pts = section(ghci, "POINTS")
    if pts:
        common["points"] = [[hex(int(a)), hex(int(b))] for a, b in (l.split() for l in pts)]
    sh = section(ghci, "SHOW")
    if sh:
        common["show_sample"] = sh[0]                       # "P 1234..." or "1234..."
    orc = section(ghci, "ORACLE")
    if orc:
        common["oracle3"] = [hex(int(v)) for v in orc]
    for ex in sorted(os.listdir(out_dir)):
        d = os.path.join(out_dir, ex)
        if not os.path.isfile(os.path.join(d, "proof.bin")):
            continue
        rec = dict(common, example=ex)
        pb = open(os.path.join(d, "points.bin"), "rb").read()
        n = int.from_bytes(pb[:8], "big")
        rec["points_bin"] = [[hex(field(pb[8 + 64 * i:8 + 64 * i + 32])), hex(field(pb[8 + 64 * i + 32:8 + 64 * i + 64]))] for i in range(n)]
        rec["proof_bin"] = open(os.path.join(d, "proof.bin"), "rb").read().hex()
        rec["commits_bin"] = open(os.path.join(d, "commits.bin"), "rb").read().hex()
        so = open(os.path.join(d, "stdout.txt")).read()
        m = re.search(r"\((\d+),(\d+)\)", so)                  # print (nrmLen, linLen)  (app/Main.hs:296)
        if m:
            rec["nrm_len"], rec["lin_len"] = int(m.group(1)), int(m.group(2))
        rec["stdout"] = so[-2000:]
        path = os.path.join(root, "tests", "golden", "ghc_%s.json" % ex)
        json.dump(rec, open(path, "w"), indent=1)
        print("wrote", path)


if __name__ == "__main__":
    main(sys.argv[1])
