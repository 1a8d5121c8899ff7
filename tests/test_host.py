"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol
include/bppp_b200.h declares, the host arithmetic the kernels share (portable multiply path,
JSF / signed-digit recoding, half-GCD, group law, the pair-fold chain) and the host transcript
match the oracle, and without a GPU the compute entry points FAIL LOUDLY (no CPU fallback)."""
import ctypes as C
import hashlib
import os
import random
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from oracle.curve import Secp256k1 as G
from oracle.field import Q, R, rational_reduce_scalar
from oracle.transcript import ZKPT, get_points, input_blinds

le = lambda x: int(x).to_bytes(32, "little")
pb = lambda p: bytes(64) if p is None else le(p[0]) + le(p[1])


def unpt(b):
    x, y = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little")
    return None if x == 0 and y == 0 else (x, y)


@pytest.fixture(scope="module")
def lib():
    import bulletproofspp_b200 as bp
    return bp.load_library()


@pytest.fixture(scope="module")
def ht():
    path = os.path.join(ROOT, "bulletproofspp_b200", "lib", "libbppp_hosttest.so")
    if not os.path.exists(path):
        pytest.skip("host self-test library not built")
    return C.CDLL(path)


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "bppp_b200.h")).read()
    declared = set(re.findall(r"\b(bppp_[a-z0-9_]+)\s*\(", hdr))
    from bulletproofspp_b200.lib import EXPORTS
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bppp_abi_version() == 1


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import bulletproofspp_b200 as bp
    with pytest.raises(bp.BpppError):
        bp.Context(0)


def test_host_sha256_fr_points_and_oracle(lib):
    from bulletproofspp_b200 import lib as L
    rnd = random.Random(1)
    for n in [0, 1, 55, 56, 63, 64, 65, 119, 120, 128, 1000, 20000]:
        d = bytes(rnd.getrandbits(8) for _ in range(n))
        out = C.create_string_buffer(32)
        lib.bppp_host_sha256(d, n, out)
        assert out.raw == hashlib.sha256(d).digest()
    for op, f in [(0, lambda a, b: a * b % R), (1, lambda a, b: (a + b) % R), (2, lambda a, b: (a - b) % R),
                  (3, lambda a, b: pow(a, -1, R)), (4, lambda a, b: (-a) % R)]:
        for a, b in [(R - 1, R - 1), (1, 0)] + [(rnd.randrange(1, R), rnd.randrange(R)) for _ in range(50)]:
            out = C.create_string_buffer(32)
            lib.bppp_host_fr(op, le(a), le(b), out)
            assert L.le_to_int(out.raw[:32]) == f(a, b)
    pts = get_points(G, "test points", 24)
    out = C.create_string_buffer(64 * 24)
    lib.bppp_host_get_points(b"test points", 24, 0, out)
    assert L.bytes_to_points(out.raw[:64 * 24]) == pts
    small = [1, 9, 10, 10 ** 19 - 1, 10 ** 19, 10 ** 38, 2 ** 64, 2 ** 128 + 5, 2 ** 255, Q - 1]
    fake = [(v, small[(i + 3) % len(small)]) for i, v in enumerate(small)]
    for fmt, name in [(0, "PrefixedP"), (1, "BareDecimal")]:
        for batch in (pts[:7], fake):
            exp = ZKPT(G, None, name).oracle(batch, 3)
            out = C.create_string_buffer(96)
            lib.bppp_host_oracle(L.points_to_bytes(batch), len(batch), 3, fmt, out)
            assert L.bytes_to_ints(out.raw[:96]) == exp
    out = C.create_string_buffer(32)
    lib.bppp_input_blind(b"default random seed", 3, out)
    assert L.le_to_int(out.raw[:32]) == input_blinds("default random seed", 3)[2]


def test_rational_reduce_matches_reference_rule(lib, ht):
    rnd = random.Random(5)
    xs = [0, 1, 2, R - 1, R - 2, (R - 1) // 2, (R + 1) // 2, 2 ** 128, 2 ** 129, 2 ** 127, R - 2 ** 128, 2 ** 255]
    for x in xs + [rnd.randrange(R) for _ in range(2000)]:
        for L_ in (lib.bppp_rational_reduce, ht.ht_rational_reduce):
            a, b, an, bn = C.create_string_buffer(32), C.create_string_buffer(32), C.c_int(), C.c_int()
            L_(le(x), a, C.byref(an), b, C.byref(bn))
            av = int.from_bytes(a.raw, "little") * (-1 if an.value else 1)
            bv = int.from_bytes(b.raw, "little") * (-1 if bn.value else 1)
            assert (av, bv) == rational_reduce_scalar(x)


def test_jsf_and_signed_digit_recoding(ht):
    rnd = random.Random(6)
    for it in range(500):
        bits = rnd.choice([1, 5, 64, 129, 130, 200, 254])
        k0, k1 = rnd.getrandbits(bits), rnd.getrandbits(bits)
        d = C.create_string_buffer(264)
        n = ht.ht_jsf(le(k0), le(k1), d, 264)
        v0 = v1 = nz = 0
        for j in range(n - 1, -1, -1):
            u0, u1 = (d.raw[j] & 3) - 1, ((d.raw[j] >> 2) & 3) - 1
            v0, v1 = 2 * v0 + u0, 2 * v1 + u1
            nz += (u0 != 0 or u1 != 0)
        assert (v0, v1) == (k0, k1) and n <= bits + 1
        if bits >= 129:
            assert nz <= 0.62 * n                                   # joint density ~ 1/2
    for it in range(500):
        s = rnd.randrange(R) if it else R - 1
        for c in (4, 6, 9, 13, 16):
            w = 256 // c + 1
            out = (C.c_int * w)()
            assert ht.ht_signed_digits(le(s), c, w, out) == 0
            assert sum(out[j] << (c * j) for j in range(w)) == s and all(abs(x) <= 1 << (c - 1) for x in out)


def test_field_and_group_law_of_the_kernel_headers(ht):
    rnd = random.Random(8)
    for op, f, m in [(0, lambda a, b: a * b % Q, Q), (1, lambda a, b: (a + b) % Q, Q), (2, lambda a, b: (a - b) % Q, Q),
                     (3, lambda a, b: pow(a, -1, Q), Q), (10, lambda a, b: pow(a, -1, R), R)]:
        for a, b in [(m - 1, m - 1), (1, 0), (2 ** 32 + 977, m - 2)] + [(rnd.randrange(1, m), rnd.randrange(m)) for _ in range(40)]:
            out = C.create_string_buffer(32)
            ht.ht_field(op, le(a), le(b), out)
            assert int.from_bytes(out.raw, "little") == f(a, b)
    pts = get_points(G, "test points", 16)

    def call(op, a, b):
        o = C.create_string_buffer(64)
        ht.ht_ec(op, pb(a), pb(b), o)
        return unpt(o.raw)
    for a, b in zip(pts, pts[1:]):
        assert call(0, a, b) == G.add(a, b) and call(1, a, b) == G.add(a, a)
        assert call(2, a, b) == G.add(G.add(a, a), G.add(a, b))
    a = pts[0]
    assert call(0, a, a) == G.add(a, a) and call(0, a, G.neg(a)) is None and call(0, None, a) == a and call(0, a, None) == a


def test_pair_fold_chain_incl_degenerate_pairs(ht):
    rnd = random.Random(9)
    pts = get_points(G, "test points", 12)
    for it in range(40):
        ka, kb = rnd.getrandbits(129), rnd.getrandbits(129)
        an, bn = rnd.random() < .5, rnd.random() < .5
        pl, pr = pts[it % 11], pts[(it * 7 + 3) % 11]
        if it == 3: pr = None
        if it == 4: pl = None
        if it == 5: pr = pl
        if it == 6: pr = G.neg(pl)
        if it == 7: ka = 0
        if it == 8: kb = 0
        o = C.create_string_buffer(64)
        ht.ht_pair_fold(le(kb), int(bn), le(ka), int(an), pb(pl), pb(pr), o)
        assert unpt(o.raw) == G.msm([(-kb if bn else kb, pl), (-ka if an else ka, pr)])


def test_batched_transcript_paths_match_oracle(lib):
    """Round challenges via oracle_rounds (all rounds at once, suffix hashing) and oracle_pair (two
    transcripts per task), and the bulk RNG (two one-block hashes at a time) against the oracle's
    plain ZKPT, for both `show` policies and odd / even counts."""
    from bulletproofspp_b200 import lib as L
    pts = get_points(G, "test points", 40)
    lib.bppp_host_transcript.argtypes = [C.c_char_p, C.c_int, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p]
    for fmt, name in [(0, "PrefixedP"), (1, "BareDecimal")]:
        for seed, n_random, npts, rounds in [("default random seed", 7, 5, 4), ("s", 1, 0, 1), ("seed#12", 10, 3, 9),
                                             ("a much longer random seed string than one SHA block can take with a counter", 3, 2, 3)]:
            z = ZKPT(G, seed, name)
            exp = [z.random() for _ in range(n_random)]
            if npts:
                z.oracle(pts[:npts], 1)
            rp = [(pts[10 + 2 * r], pts[11 + 2 * r]) for r in range(rounds)]
            es = [z.oracle([x, y], 1)[0] for x, y in rp]
            out = C.create_string_buffer(32 * (n_random + 2 * rounds))
            rc = lib.bppp_host_transcript(seed.encode(), fmt, n_random, L.points_to_bytes(pts[:npts]), npts,
                                          L.points_to_bytes([p for xy in rp for p in xy]), rounds, out)
            assert rc == 0
            got = L.bytes_to_ints(out.raw)
            assert got[:n_random] == exp
            assert got[n_random:n_random + rounds] == es
            assert got[n_random + rounds:] == es


def test_host_job_scheduler_runs_every_item_once(lib):
    """The priority job scheduler that runs the lanes' host phases (rp_host.cpp): many concurrent
    submitters, jobs of 1..300 items, repeated."""
    for lanes, jobs, items in [(1, 3, 1), (4, 20, 7), (16, 30, 64), (8, 10, 300), (16, 200, 3)]:
        assert lib.bppp_host_scheduler_selftest(lanes, jobs, items) == 0
    assert lib.bppp_host_scheduler_selftest(0, 1, 1) == -1


def test_glv_constants_and_split_of_the_pippenger_kernel():
    """The endomorphism constants of csrc/pippenger.cuh, parsed from the source, and its scalar split
    restated with big integers: lambda^3 = 1 (mod r), beta^3 = 1 (mod q), lambda G = (beta Gx, Gy),
    k = k1 + k2 lambda with |k1|, |k2| < 2^128 for edge values and random scalars."""
    import random
    from oracle.curve import Secp256k1 as G
    from oracle.field import Q, R
    src = open(os.path.join(ROOT, "bulletproofspp_b200", "csrc", "pippenger.cuh")).read()

    def const(name):
        m = re.search(r"#define %s glv_const\(([^)]*)\)" % name, src)
        limbs = [int(x.strip().rstrip("u"), 16) for x in m.group(1).split(",")]
        return sum(v << (32 * i) for i, v in enumerate(limbs))
    g1, g2, mb1, b2, lam, beta = (const(n) for n in ("GLV_G1", "GLV_G2", "GLV_MB1", "GLV_B2", "GLV_LAMBDA", "GLV_BETA"))
    assert pow(lam, 3, R) == 1 and lam != 1 and pow(beta, 3, Q) == 1 and beta != 1
    assert G.mul(lam, G.gen) == (beta * G.gen[0] % Q, G.gen[1])
    assert (b2 - mb1 * lam) % R == 0                      # (a1, b1) = (b2, -mb1) is a lattice vector
    assert g1 == (b2 * (1 << 384) + R // 2) // R and g2 == (mb1 * (1 << 384) + R // 2) // R

    def split(k):                                         # k_pip_glv, statement by statement
        c1 = (k * g1 + (1 << 383)) >> 384
        c2 = (k * g2 + (1 << 383)) >> 384
        t1, t2 = c1 * mb1, c2 * b2
        assert c1 < 1 << 128 and c2 < 1 << 128 and t1 < 1 << 256 and t2 < 1 << 256
        neg2, m2 = t1 < t2, abs(t1 - t2)
        prod = m2 * lam % R
        k1 = (k + prod) % R if neg2 else (k - prod) % R
        neg1 = k1 > (R - 1) // 2
        return (R - k1 if neg1 else k1), neg1, m2, neg2
    rnd = random.Random(7)
    ks = [1, 2, R - 1, R - 2, (1 << 128) - 1, 1 << 128, 1 << 255, R // 2, R // 2 + 1, lam, R - lam, 255]
    ks += [rnd.randrange(R) for _ in range(20000)]
    for k in ks:
        m1, s1, m2, s2 = split(k)
        assert m1 < 1 << 128 and m2 < 1 << 128
        assert ((-m1 if s1 else m1) + (-m2 if s2 else m2) * lam) % R == k


def test_haskell_shim_binds_only_declared_entry_points_with_the_right_arity():
    """hs/ and INTEGRATION.md: every `foreign import ccall` names a function include/bppp_b200.h declares, and the
    Haskell type has as many arguments as the C prototype (the shim cannot be compiled here: no GHC)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "bppp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|void|const char\*|size_t)\s+(bppp_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    assert len(protos) > 80
    checked = 0
    for rel in ("hs/Bulletproof/B200/FFI.hs", "INTEGRATION.md"):
        text = open(os.path.join(root, rel)).read()
        for m in re.finditer(r'foreign import ccall (?:safe|unsafe) "(\w+)"\s+\w+\s*::\s*((?:[^\n]|\n\s+->)*)', text):
            name, ty = m.group(1), m.group(2)
            ty = re.sub(r"--.*", "", ty)
            assert name in protos, "%s binds %s, which the header does not declare" % (rel, name)
            depth, arrows = 0, 0
            for i, ch in enumerate(ty):
                depth += ch == "("
                depth -= ch == ")"
                arrows += depth == 0 and ty[i:i + 2] == "->"
            assert arrows == protos[name], "%s: %s has %d arguments in Haskell, %d in C" % (rel, name, arrows, protos[name])
            checked += 1
    assert checked >= 40
