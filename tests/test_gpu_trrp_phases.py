"""The TypedReciprocal scalar phases on the device (bppp_trrp_*, include/bppp_b200.h) called
directly through the C ABI and compared, phase by phase, with the oracle's restatement of
makePhase2s / makeErrorTerms / makePublicConsts (TypedReciprocal.hs:165-263, 399-444) on a typed
configuration with shared, inline, has-bit and assumed ranges.  The challenges are arbitrary hash
outputs (not transcript derived), different per proof."""
import ctypes as C
import hashlib
import struct

import pytest

pytestmark = pytest.mark.gpu

from example_configs import EXAMPLES
from oracle import rangeproof as orp
from oracle.curve import Secp256k1 as G
from oracle.field import R, batch_inverse

vp, sz, u8p = C.c_void_p, C.c_size_t, C.c_char_p


def h(*a):
    return int.from_bytes(hashlib.sha256(repr(a).encode()).digest(), "big") % R


def test_device_phases_match_oracle_functions(ctx):
    from bulletproofspp_b200 import lib as L
    lib = ctx.lib
    lib.bppp_trrp_create.argtypes = [vp, sz, u8p, u8p, u8p, sz, sz, C.POINTER(vp)]
    lib.bppp_trrp_destroy.argtypes = [vp]
    lib.bppp_trrp_destroy.restype = None
    lib.bppp_trrp_phase1.argtypes = [vp, sz, u8p, u8p, u8p]
    lib.bppp_trrp_phase2.argtypes = [vp, u8p, u8p, sz, u8p, u8p]
    lib.bppp_trrp_phase3.argtypes = [vp, u8p, u8p, u8p]
    lib.bppp_trrp_commit_bl.argtypes = [vp, u8p, u8p]
    lib.bppp_trrp_phase4.argtypes = [vp, u8p, u8p]
    lib.bppp_nl_create_trrp.argtypes = [vp, u8p, u8p, u8p, u8p, C.POINTER(vp)]
    lib.bppp_trrp_verify_pub.argtypes = [vp, sz, u8p, u8p]
    schema, wit0 = EXAMPLES["typed_nl"]
    st = orp.load_schema(schema, G)
    N, M, P0 = st.nrm_len, st.lin_len, 1 + st.nrm_len + st.lin_len
    # static entry descriptors, as bppp_rp_setup derives them from the verifier's Phase-1 template
    tmpl = st.ph1s_verifier()
    assert len(tmpl) == N
    desc, eb, es = b"", [], []
    for p in tmpl:
        if p[0] == "T":
            flags, ind, bi, b, s = 1 | (2 if p[2] else 0) | (4 if p[3] else 0), p[1], -1, 0, 0
        elif p[0] == "I":
            flags, ind, bi, b, s = 8 | (16 if p[6] % R else 0), p[1], st.sorted_bases.index(p[2]), p[3] % R, p[6] % R
        else:
            flags, ind, bi, b, s = 0, p[1], st.sorted_bases.index(p[2]), p[3] % R, 0
        desc += struct.pack("<Iiii", flags, ind, bi, 0)
        eb.append(b); es.append(s)
    gh, th = vp(), vp()
    ctx._ck(lib.bppp_gens_create(ctx.h, N, M, L.point_to_bytes(st.g), L.points_to_bytes(st.gs), L.points_to_bytes(st.hs),
                                 C.byref(gh)), "bppp_gens_create")
    ctx._ck(lib.bppp_trrp_create(gh, N, desc, L.ints_to_bytes(eb), L.ints_to_bytes(es), len(st.rds), len(st.sorted_bases),
                                 C.byref(th)), "bppp_trrp_create")
    B = 2
    wits = [orp.load_witness(st, wit0) for _ in range(B)]

    def row(w):                              # commitment scalars of an RPW over [g | gs | hs]
        r = [0] * P0
        r[0] = w.sc % R
        for i, v in enumerate(w.nrm[:N]):
            r[1 + i] = v % R
        for i, v in enumerate(w.lin[:M]):
            r[1 + N + i] = v % R
        return r
    sclin = lambda w: [w.sc % R] + [(w.lin[i] % R if i < len(w.lin) else 0) for i in range(M)]
    dms, ms, amounts = [], [], []
    for b, w in enumerate(wits):
        ds, mi = [], []
        for p in w["ph1s"]:
            ds.append(p[4] if p[0] in "IS" else p[5]); mi.append(p[5] if p[0] == "I" else 0)
        ms_shared = [m for _, mm in w["base_mss"] for m in mm]
        dms.append(orp.RPW(h("dm", b), [h("dml", b, i) for i in range(6)] + ms_shared, ds))
        ms.append(orp.RPW(h("m", b), [h("ml", b, i) for i in range(6)], mi))
        amounts.append([v % R for v, _, _ in w["inputs"]])
    out2 = C.create_string_buffer(64 * 2 * B)
    ctx._ck(lib.bppp_trrp_phase1(th, B, b"".join(L.ints_to_bytes(row(x)) for pair in zip(dms, ms) for x in pair),
                                 b"".join(L.ints_to_bytes(a) for a in amounts), out2), "phase1")
    got = L.bytes_to_points(out2.raw)
    for b in range(B):
        assert got[2 * b] == G.msm(st.com_terms(dms[b])) and got[2 * b + 1] == G.msm(st.com_terms(ms[b]))
    # ---- phase 2
    ch2, exp = [], []
    for b, w in enumerate(wits):
        e, x, r0 = h("e", b), h("x", b), h("r0", b)
        e_inv, r0_inv = batch_inverse([e, r0])
        ph2s = orp.make_phase2s(True, e, e_inv, x, st.base_map(x), w["ph1s"])
        err7 = r0_inv * (-sum(2 * o.r * o.c for o in ph2s)) % R
        rw = orp.RPW(h("rsc", b), [h("rl", b, i) for i in range(4)] + [0, 0], [o.r for o in ph2s])
        ch2.append([e, e_inv, x, r0_inv]); exp.append(dict(e=e, e_inv=e_inv, x=x, r0=r0, ph2s=ph2s, err7=err7, rw=rw))
    rcom, err7 = C.create_string_buffer(64 * B), C.create_string_buffer(32 * B)
    ctx._ck(lib.bppp_trrp_phase2(th, b"".join(L.ints_to_bytes(c) for c in ch2),
                                 b"".join(L.ints_to_bytes(sclin(x["rw"])) for x in exp), 4, rcom, err7), "phase2")
    for b, x in enumerate(exp):
        assert L.bytes_to_ints(err7.raw)[b] == x["err7"]
        x["rw"].lin[4] = x["err7"]
        assert L.bytes_to_points(rcom.raw)[b] == G.msm(st.com_terms(x["rw"]))
    assert any(x["err7"] for x in exp)       # the configuration has inline symbols: the c_i path is live
    # ---- phase 3: error terms over the norm entries (no shared-multiplicity part)
    ch3, bls = [], []
    for b, x in enumerate(exp):
        q, xq = h("q", b), h("xq", b)
        q0 = st.q_powers(q, 1)[0]
        bl = [h("bl", b, i) for i in range(N)]
        x.update(q=q, xq=xq, q0=q0, bls=bl)
        x["errs"] = orp.make_error_terms(x["e"], xq, [], [], list(zip(x["ph2s"], st.q_powers(q, N), bl)))
        ch3.append([q0, xq]); bls.append(bl)
    errs = C.create_string_buffer(32 * 6 * B)
    ctx._ck(lib.bppp_trrp_phase3(th, b"".join(L.ints_to_bytes(c) for c in ch3), b"".join(L.ints_to_bytes(v) for v in bls),
                                 errs), "phase3")
    got = L.bytes_to_ints(errs.raw)
    for b, x in enumerate(exp):
        assert got[6 * b:6 * b + 6] == x["errs"]
    # ---- blinding commitment: scalar + linear slots from the host, norm part = the phase-3 blinders
    blc = C.create_string_buffer(64 * B)
    blws = [orp.RPW(h("blsc", b), [h("bll", b, i) for i in range(M)], exp[b]["bls"]) for b in range(B)]
    ctx._ck(lib.bppp_trrp_commit_bl(th, b"".join(L.ints_to_bytes(sclin(w)) for w in blws), blc), "commit_bl")
    for b in range(B):
        assert L.bytes_to_points(blc.raw)[b] == G.msm(st.com_terms(blws[b]))
    # ---- phase 4: public constants + combined witness
    ch4 = []
    for b, x in enumerate(exp):
        t = h("t", b)
        q0_inv = pow(x["q0"], -1, R)
        x.update(t=t, q0_inv=q0_inv)
        x["pub"] = orp.make_public_consts(x["e"], x["e_inv"], x["x"], x["xq"], x["q0"], q0_inv, t, st.has_types, st.rds,
                                          st.pub_vt, x["ph2s"])
        ch4.append([t, q0_inv])
    sums = C.create_string_buffer(32 * 3 * B)
    ctx._ck(lib.bppp_trrp_phase4(th, b"".join(L.ints_to_bytes(c) for c in ch4), sums), "phase4")
    got = L.bytes_to_ints(sums.raw)
    nlh = vp()
    zeros = L.ints_to_bytes([0] * (B * M))
    ctx._ck(lib.bppp_nl_create_trrp(th, L.ints_to_bytes([x["q"] for x in exp]), L.ints_to_bytes([0] * B), zeros, zeros,
                                    C.byref(nlh)), "bppp_nl_create_trrp")
    fs, fw, fl = C.create_string_buffer(32 * B), C.create_string_buffer(32 * B * N), C.create_string_buffer(32 * B * M)
    ctx._ck(lib.bppp_nl_final(nlh, fs, fw, fl), "bppp_nl_final")
    lib.bppp_nl_destroy(nlh)
    w_dev = L.bytes_to_ints(fw.raw)
    for b, x in enumerate(exp):
        t, t5 = x["t"], pow(x["t"], 5, R)
        ts0, sq2, sv = got[3 * b:3 * b + 3]
        # pub.sc = z + ts0 + 2 t^5 (sum q2 + 1/e sum v): recover z from the oracle's own formula
        mins = [0 if rd.is_assumed else rd.min % R for rd in st.rds]
        x2 = x["x"] * x["x"] % R
        z = -2 * t5 * sum(a * pow(x2, i + 1, R) for i, a in enumerate(mins))
        if st.has_types:
            rs = batch_inverse([(x["e"] + tt) % R for _, tt, _ in st.pub_vt])
            z -= 2 * t5 * x["x"] * (sum((-1 if io else 1) * r * (v % R) for (io, _, v), r in zip(st.pub_vt, rs)) % R)
        assert (z + ts0 + 2 * t5 * (sq2 + x["e_inv"] * sv)) % R == x["pub"].sc % R
        want = [(p + bl + t * m + t * t * d + pow(t, 3, R) * r) % R
                for p, bl, m, d, r in zip(x["pub"].nrm, x["bls"], ms[b].nrm, dms[b].nrm, x["rw"].nrm)]
        assert w_dev[b * N:(b + 1) * N] == want
    # ---- verifier: the same public constants from public data only
    chv = [[x["e"], x["e_inv"], x["x"], x["xq"], x["q0"], x["q0_inv"], x["t"], 0] for x in exp]
    vs = C.create_string_buffer(32 * 3 * B)
    ctx._ck(lib.bppp_trrp_verify_pub(th, B, b"".join(L.ints_to_bytes(c) for c in chv), vs), "verify_pub")
    assert L.bytes_to_ints(vs.raw) == got    # ts0, sum q2, sum v do not depend on the witness
    lib.bppp_trrp_destroy(th)
    lib.bppp_gens_destroy(gh)
