"""Range-proof layer (oracle; test infrastructure only).

Literal restatement of
  src/RangeProof/Internal.hs        (RPWitness, commitRPW, blinding helpers, makePolyTerms)
  src/RangeProof/TypedReciprocal.hs (reciprocal range proof + typed conservation of money)
  src/RangeProof/Binary.hs          (binary range proof)
  src/RangeProof.hs                 (RangeProof ZKP wrapper)
  app/Parse.hs, app/Main.hs         (schema / witness loading, setup wiring)
The argument itself is oracle.bulletproof; hashing is oracle.transcript.
"""
import json
from collections import Counter

from .field import R, inv, batch_inverse, powers1, powers2
from . import bulletproof as bp
from .transcript import ZKPT, get_points, input_blinds


# ------------------------------------------------------------------ RPWitness
class RPW:
    """`RPWitness` (Internal.hs:22-41): scalar, linear list, norm list; a vector space with
    zero-padding addition."""

    def __init__(self, sc=0, lin=(), nrm=()):
        self.sc, self.lin, self.nrm = sc % R, [v % R for v in lin], [v % R for v in nrm]

    def __add__(self, o):
        def z(a, b):
            L = max(len(a), len(b))
            return [((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % R for i in range(L)]
        return RPW(self.sc + o.sc, z(self.lin, o.lin), z(self.nrm, o.nrm))

    def scale(self, s):
        return RPW(self.sc * s, [s * v for v in self.lin], [s * v for v in self.nrm])


def sum_v(ws):
    acc = RPW()
    for w in ws:
        acc = acc + w
    return acc


def rpw_terms(w, g, hs, gs, G):
    """`commitRPW` opening (Internal.hs:43-48; `dotWith` pads with 0 / identity)."""
    t = [(w.sc, g)]
    t += [(w.lin[i] if i < len(w.lin) else 0, hs[i] if i < len(hs) else G.zero) for i in range(max(len(hs), len(w.lin)))]
    t += [(w.nrm[i] if i < len(w.nrm) else 0, gs[i] if i < len(gs) else G.zero) for i in range(max(len(gs), len(w.nrm)))]
    return t


def integer_log(b, n):
    """src/Utils.hs:91-92."""
    r = 0
    while n >= b:
        n //= b
        r += 1
    return r


def insert_at(n, x, xs):
    return xs[:n] + [x] + xs[n:]


def pad_right(n, x, xs):
    return (list(xs) + [x] * n)[:n]


def blind_witness(zk, n, k, ls, ns):
    """Internal.hs:134-142."""
    n_bls = 2 * n - 1 if k == 1 else 2 * n - k + 1
    bls = pad_right(2 * n + 1, 0, insert_at(2 * n - k, 0, [zk.random() for _ in range(n_bls)]))
    return RPW(bls[0], bls[1:] + list(ls), ns)


def blind_err_witness(zk, n, es, ls, ns):
    """Internal.hs:145-152."""
    bls = pad_right(2 * n + 1, 0, insert_at(n, 0, [zk.random() for _ in range(n + 1)]) + list(es))
    return RPW(bls[0], bls[1:] + list(ls), ns)


def scale_errs(n, k, xs):
    """Internal.hs:118-121."""
    ys, zs = xs[:n + 1], xs[n + 1:]
    return ys + [k * a % R for a in zs[:n - 2]] + zs[n - 2:]


def sum_diagonals(xss):
    """Internal.hs:105-111."""
    m = {}
    for a, xs in enumerate(xss):
        for b, x in enumerate(xs):
            m[a + b] = (m.get(a + b, 0) + x) % R
    return [m[k] for k in sorted(m)]


def blind_blinding_term(bl, tC, r0, r0i, r1, r1i, errs, wits, input_bl):
    """Internal.hs:157-195."""
    assert bl.sc == 0
    blT, bls_lin, bls_nrm = bl.lin[0], bl.lin[1:], bl.nrm
    rs_inv = r0i * r1i % R
    n = len(wits)
    wits1, wit_err = wits[:n - 1], wits[n - 1]
    wit_err1 = [wit_err.sc] + pad_right(2 * n, 0, wit_err.lin[:n + 1])
    wit_rows = [[w.sc] + w.lin[:2 * n] for w in wits1]
    rows = [[r[0], r[1]] + [(-v) % R for v in r[2:]] for r in wit_rows + [wit_err1]]
    errs1 = [(-v) % R for v in [(errs[0] - tC * blT) % R] + [rs_inv * v % R for v in errs[1:]]]

    def add_consts(a, b, r):
        return [(a * r[0] + b * r[1]) % R] + r[2:]
    table = [insert_at(2 * n - 1, 0, row) for row in
             [errs1] + [scale_errs(n, r1i, add_consts(rs_inv, rs_inv * tC % R, r)) for r in rows]]
    sd = sum_diagonals(table)
    sd = sd[:2 * n - 1] + sd[2 * n:]                         # removeAt (2n-1)
    bl_errs = scale_errs(n, r1, sd[:2 * n])
    bl_errs[-1] = (bl_errs[-1] - 2 * input_bl) % R
    return RPW(-bl_errs[0], [blT] + bl_errs[1:] + bls_lin, bls_nrm)


def make_poly_terms(ws, tss):
    """Internal.hs:65-75, for the two-row case used by Binary.hs:188."""
    def wdot(a, b):
        return sum(w * x * y for w, x, y in zip(ws, a, b)) % R
    t0, t1 = tss
    return [wdot(t0, t0), 2 * wdot(t0, t1) % R, wdot(t1, t1)]


# ------------------------------------------------------ TypedReciprocal (TRRP)
class RangeData:
    """TypedReciprocal.hs:73-115 (`makeRangeData`)."""

    def __init__(self, char, base, mn, mx, is_shared, is_output, is_assumed):
        if not (mx > mn and base > 1 and mx - mn < char):
            raise ValueError("Invalid range")
        b = base
        n1 = integer_log(b, mx - mn - 1)
        self.base, self.min, self.max = b, mn, mx
        self.is_shared, self.is_output, self.is_assumed = is_shared, is_output, is_assumed
        self.has_bit = ((mx - mn - 1) % (b - 1)) != 0
        tail = [b ** (n1 - i) for i in range(1, n1 + 1)]
        if not self.has_bit:
            bs = [(mx - mn - b ** n1) // (b - 1)] + tail
        elif mx - mn < 2 * b ** n1:
            bs = [mx - mn - b ** n1] + tail
        else:
            bn1 = 1 + (mx - mn) // (2 * (b - 1)) - (b ** n1 - 1) // (b - 1)
            bs = [mx - mn - bn1 * (b - 1) - b ** n1, bn1] + tail
        self.base_coeffs = [] if is_assumed else bs


def digits(rd, n):
    """TypedReciprocal.hs:120-122."""
    out = []
    bases = ([2] if rd.has_bit else []) + [rd.base] * len(rd.base_coeffs)
    for b, base in zip(rd.base_coeffs, bases):
        d = min(base - 1, n // b)
        n -= d * b
        out.append(d)
    return out


def make_phase1s(ind, rd, n):
    """TypedReciprocal.hs:128-152.  n = None for the verifier (`makePhase1sVer`).  Entries:
    ('I', ind, base, b, d, m, s) | ('S', ind, base, b, d)."""
    if rd.is_assumed:
        return [], None
    prover = n is not None
    if prover:
        n_adj = (n - rd.min) % R
        if not (0 <= n_adj < rd.max - rd.min):
            raise ValueError("witness out of range")
        ds = digits(rd, n_adj)
    else:
        ds = [None] * len(rd.base_coeffs)
    base = rd.base
    if prover:
        cnt = Counter(ds[1:] if rd.has_bit else ds)
        ms = ([ds[0]] if rd.has_bit else []) + [cnt.get(i, 0) for i in range(1, base)]
    else:
        ms = [None] * (base - 1 + (1 if rd.has_bit else 0))
    ns = ([1] if rd.has_bit else []) + list(range(1, base))
    bs = rd.base_coeffs
    bases = ([2] if rd.has_bit else []) + [base] * max(len(bs), len(ms))
    if rd.is_shared:
        return [("S", ind, bb, b % R, d) for bb, b, d in zip(bases, bs, ds)], ms
    L = max(len(bs), len(ds), len(ms), len(ns))
    zero = 0 if prover else None
    bs2, ns2 = pad_right(L, 0, bs), pad_right(L, 0, ns)
    ds2, ms2 = pad_right(L, zero, ds), pad_right(L, zero, ms)
    return [("I", ind, bases[i], bs2[i] % R, ds2[i], ms2[i], ns2[i]) for i in range(L)], None


class Ph2:
    __slots__ = ("isT", "d", "m", "u", "v", "r", "c")


def make_phase2s(prover, e, e_inv, x, base_map, ph1s):
    """TypedReciprocal.hs:165-195."""
    ds, ss, ps, vs, out = [], [], [], [], []
    for p in ph1s:
        o = Ph2()
        xp = pow(x, 2 * (p[1] + 1), R)
        if p[0] == "T":
            _, _, io, ia, v, t = p
            xpp = (-x) % R if io else x
            ds.append((e + t) % R if prover else None)
            ss.append(0)
            ps.append(v)
            o.isT, o.d, o.m, o.u, o.v = True, t, 0, (0 if ia else xp), xpp
        elif p[0] == "I":
            _, _, base, b, d, m, s = p
            xpp = base_map[base]
            ds.append((e + d) % R if prover else None)
            ss.append(0 if s == 0 else (e + s) % R)
            ps.append(1)
            o.isT, o.d, o.m, o.u, o.v = False, d, m, xp * b % R, xpp
        else:
            _, _, base, b, d = p
            xpp = base_map[base]
            ds.append((e + d) % R if prover else None)
            ss.append(0)
            ps.append(1)
            o.isT, o.d, o.m, o.u, o.v = False, d, 0, xp * b % R, xpp
        vs.append(xpp)
        out.append(o)
    rs = [pp * di % R for pp, di in zip(ps, batch_inverse(ds))] if prover else [None] * len(out)
    cs = [v * (0 if s == 0 else (e_inv - s)) % R for v, s in zip(vs, batch_inverse(ss))]
    for o, r, c in zip(out, rs, cs):
        o.r, o.c = r, c
    return out


def make_shared_coeffs(e, e_inv, m_bases, base_map):
    """TypedReciprocal.hs:204-206."""
    xs, ss = [], []
    for b in m_bases:
        for s in range(1, b):
            xs.append(base_map[b])
            ss.append((e + s) % R)
    return [x * (e_inv - si) % R for x, si in zip(xs, batch_inverse(ss))]


def make_error_terms(e, xq, shared_cs, bls_ms, ph3s):
    """TypedReciprocal.hs:217-233."""
    aug = 2 * sum(a * b for a, b in zip(shared_cs, bls_ms)) % R
    tot = [0, 0, 0, aug, 0, 0]
    for o, q2, bl in ph3s:
        d, m, u, v, r, c = o.d, o.m, o.u, o.v, o.r, o.c
        rC = xq * (u + q2) % R if o.isT else u
        dC = (v + q2 * e) % R
        qd, qr = (q2 * d + dC) % R, (q2 * r + rC) % R
        errs = [q2 * bl * bl,
                2 * q2 * m * bl,
                q2 * m * m + 2 * bl * qd,
                2 * (bl * qr + m * qd),
                (q2 * d * d + 2 * d * dC) + 2 * (bl * c + m * qr),
                (q2 * r * r + 2 * r * rC) + 2 * c * d]
        tot = [(a + b) % R for a, b in zip(tot, errs)]
    return tot


def make_public_consts(e, e_inv, x, xq, q0, q0_inv, t, has_types, rds, pub_vt, ph2s):
    """TypedReciprocal.hs:236-263."""
    mins = [0 if rd.is_assumed else rd.min % R for rd in rds]
    t5 = pow(t, 5, R)
    z = -2 * t5 * sum(a * b for a, b in zip(mins, powers1(x * x % R, len(mins))))
    if has_types:
        pub_rs = batch_inverse([(e + tt) % R for _, tt, _ in pub_vt])
        pub_sum = sum((-1 if io else 1) * r * (v % R) for (io, _, v), r in zip(pub_vt, pub_rs)) % R
        z -= 2 * t5 * x * pub_sum
    t2, t3, t4 = t * t % R, pow(t, 3, R), pow(t, 4, R)
    ts0, ts1 = 0, []
    for o, q2, qi2 in zip(ph2s, powers1(q0, len(ph2s)), powers1(q0_inv, len(ph2s))):
        if o.isT:
            rC, p2C = xq * (qi2 * o.u + 1) % R, 0
        else:
            rC, p2C = qi2 * o.u % R, (2 * q2 + 2 * e_inv * o.v) % R
        p = (t2 * (e + qi2 * o.v) + t3 * rC + t4 * (qi2 * o.c % R)) % R
        ts0 += q2 * p * p + t5 * p2C
        ts1.append(p)
    return RPW(z + ts0, [], ts1)


def input_coeffs_trrp(has_types, assumed, x, q0):
    """TypedReciprocal.hs:325-328."""
    xp = [0 if a else v for a, v in zip(assumed, powers1(x * x % R, len(assumed)))]
    if has_types:
        xp = [(a + b) % R for a, b in zip(powers1(q0, len(xp)), xp)]
    return xp


def make_bp_coeffs(has_types, xq, r0, r1, t, cs):
    """TypedReciprocal.hs:391-396."""
    rs = r0 * r1 % R
    t3 = pow(t, 3, R)
    return [(-xq) % R if has_types else 0, rs * t % R, rs * t * t % R, rs * t3 % R,
            r0 * pow(t, 4, R) % R, rs * pow(t, 6, R) % R] + [2 * t3 * c % R for c in cs]


class SetupTRRP:
    """`setup` (TypedReciprocal.hs:332-359)."""
    kind = "TRRP"
    num_rp_coms = 4

    def __init__(self, G, arg, ps, has_types, pub_vt, rds):
        self.G, self.arg, self.has_types, self.pub_vt, self.rds = G, arg, has_types, pub_vt, rds
        live = [rd for rd in rds if not rd.is_assumed]
        any_has_bit = any(rd.has_bit for rd in live)
        any_shared_has_bit = any(rd.has_bit and rd.is_shared for rd in live)
        shared = sorted(rd.base for rd in live if rd.is_shared)
        allb = sorted(rd.base for rd in live)
        self.m_bases = sorted(set(([2] if any_shared_has_bit else []) + shared))
        self.sorted_bases = sorted(set(([2] if any_has_bit else []) + allb))
        self.nrm_len = sum(len(rd.base_coeffs) + (1 if has_types else 0) for rd in rds)
        self.lin_len = 6 + sum(b - 1 for b in self.m_bases)
        self.h, self.g = ps[0], ps[1]
        rest = ps[2:]
        if len(rest) < self.lin_len + self.nrm_len:
            raise ValueError("setup failed: not enough points")
        self.hs = rest[:self.lin_len]
        self.gs = rest[self.lin_len:self.lin_len + self.nrm_len]

    @staticmethod
    def points_needed(has_types, rds):
        live = [rd for rd in rds if not rd.is_assumed]
        m = sorted(set(([2] if any(rd.has_bit and rd.is_shared for rd in live) else []) +
                       [rd.base for rd in live if rd.is_shared]))
        return 2 + 6 + sum(b - 1 for b in m) + sum(len(rd.base_coeffs) + (1 if has_types else 0) for rd in rds)

    def base_map(self, x):
        return dict(zip(self.sorted_bases, powers2(pow(x, 3, R), x * x % R, len(self.sorted_bases))))

    def q_powers(self, q, n):
        return bp.q_powers(self.arg, q, n)

    def com_terms(self, w):
        return rpw_terms(w, self.g, self.hs, self.gs, self.G)

    def psv(self, q, cs, w):
        return bp.PSV(w.sc, self.g, bp.NormLinear.make(self.arg, self.G, q, cs, w.nrm, self.gs, w.lin, self.hs))

    def rounds(self):
        return bp.optimal_witness_size(self.arg, self.nrm_len, self.lin_len)[0]

    def info(self):
        return (4, self.nrm_len, self.lin_len)

    # -- witness (TypedReciprocal.hs:373-388)
    def witness(self, inputs):
        """inputs: [(value, type, blind)] as field elements (the PedersenScalarPair of app/Main.hs:287)."""
        vs = [v % R for v, _, _ in inputs]
        ts = [t % R for _, t, _ in inputs]
        if self.has_types:
            sums = {}
            for io, t, v in self.pub_vt:
                sums[t % R] = (sums.get(t % R, 0) + (-v if io else v)) % R
            for t, v, rd in zip(ts, vs, self.rds):
                sums[t] = (sums.get(t, 0) + (-v if rd.is_output else v)) % R
            if any(sums.values()):
                raise ValueError("unbalanced types")
        ph1ss, mss = [], []
        for i, (rd, v) in enumerate(zip(self.rds, vs)):
            a, b = make_phase1s(i, rd, v)
            ph1ss.append(a)
            mss.append(b)
        types = [("T", i, rd.is_output, rd.is_assumed, v, t) for i, (rd, v, t) in enumerate(zip(self.rds, vs, ts))]
        ph1s = (types if self.has_types else []) + [p for l in ph1ss for p in l]
        bm = {}
        for rd, ms in zip(self.rds, mss):                       # baseMss (:363-367)
            if ms is None:
                continue
            ents = [(2, [ms[0]]), (rd.base, ms[1:])] if rd.has_bit else [(rd.base, ms)]
            for b, m in ents:
                bm[b] = [(p + q) % R for p, q in zip(bm[b], m)] if b in bm else list(m)
        return dict(inputs=inputs, ph1s=ph1s, base_mss=sorted(bm.items()))

    def ph1s_verifier(self):
        ph1ss = [make_phase1s(i, rd, None)[0] for i, rd in enumerate(self.rds)]
        types = [("T", i, rd.is_output, rd.is_assumed, None, None) for i, rd in enumerate(self.rds)]
        return (types if self.has_types else []) + [p for l in ph1ss for p in l]

    # -- prover (TypedReciprocal.hs:399-444)
    def prove_rp(self, zk, wit, trace=None):
        G = self.G
        is_as = [rd.is_assumed for rd in self.rds]
        m_bases = [b for b, _ in wit["base_mss"]]
        ms_shared = [m for _, ms in wit["base_mss"] for m in ms]
        ds, ms_inline = [], []
        for p in wit["ph1s"]:
            if p[0] == "I":
                ds.append(p[4]); ms_inline.append(p[5])
            elif p[0] == "S":
                ds.append(p[4]); ms_inline.append(0)
            else:
                ds.append(p[5]); ms_inline.append(0)
        n_wits = [RPW(v, [t, bl], []) for v, t, bl in wit["inputs"]]
        dm_wit = blind_witness(zk, 3, 2, ms_shared, ds)
        m_wit = blind_witness(zk, 3, 1, [], ms_inline)
        coms1 = G.msm_many([self.com_terms(w) for w in n_wits + [dm_wit, m_wit]])
        n_coms, dm_com, m_com = coms1[:-2], coms1[-2], coms1[-1]
        e, x, r0 = zk.oracle([dm_com, m_com] + n_coms, 3)
        e_inv, r0_inv = batch_inverse([e, r0])
        base_map = self.base_map(x)
        ph2s = make_phase2s(True, e, e_inv, x, base_map, wit["ph1s"])
        err7 = r0_inv * (-sum(2 * o.r * o.c for o in ph2s)) % R
        r_wit = blind_err_witness(zk, 3, [err7], [], [o.r for o in ph2s])
        r_com = G.msm(self.com_terms(r_wit))
        q, xq, r1 = zk.oracle([r_com], 3)
        q0 = self.q_powers(q, 1)[0]
        q_inv, q0_inv, r1_inv = batch_inverse([q, q0, r1])
        shared_cs = make_shared_coeffs(e, e_inv, m_bases, base_map)
        tC = xq if self.has_types else 0
        bls_lin = [zk.random() for _ in range(self.lin_len - 5)]
        bls_nrm = [zk.random() for _ in range(self.nrm_len)]
        bls_ms = bls_lin[1:]
        n_wit_sum = sum_v(w.scale(c) for c, w in zip(input_coeffs_trrp(self.has_types, is_as, x, q0), n_wits))
        assert len(n_wit_sum.lin) == 2
        input_bl = n_wit_sum.lin[1]
        L = len(ph2s)
        ph3s = list(zip(ph2s, self.q_powers(q, L), bls_nrm))
        errs = make_error_terms(e, xq, shared_cs, bls_ms, ph3s)
        bl_wit = blind_blinding_term(RPW(0, bls_lin, bls_nrm), tC, r0, r0_inv, r1, r1_inv, errs,
                                     [m_wit, dm_wit, r_wit], input_bl)
        bl_com = G.msm(self.com_terms(bl_wit))
        t = zk.oracle([bl_com], 1)[0]
        pub = make_public_consts(e, e_inv, x, xq, q0, q0_inv, t, self.has_types, self.rds, self.pub_vt, ph2s)
        w = (pub + bl_wit + m_wit.scale(t) + dm_wit.scale(t * t) + r_wit.scale(pow(t, 3, R))
             + n_wit_sum.scale(2 * pow(t, 5, R)))
        coms = [bl_com, r_com, dm_com, m_com] + n_coms
        cs = make_bp_coeffs(self.has_types, xq, r0, r1, t, shared_cs)
        ch = dict(e=e, x=x, r0=r0, q=q, xq=xq, r1=r1, q0=q0, t=t)
        if trace is not None:
            trace.update(ch=ch, wits=dict(n=n_wits, dm=dm_wit, m=m_wit, r=r_wit, bl=bl_wit), pub=pub, wit=w, cs=cs)
        return coms, self._bp_setup(q, cs, pub, coms, ch), self.psv(q, cs, w)

    def init_open(self, coms, ch):
        """`TranscriptTRRP.openWith` (TypedReciprocal.hs:279-282)."""
        bl, r, dm, m = coms[:4]
        t = ch["t"]
        ss = [2 * pow(t, 5, R) * c % R for c in
              input_coeffs_trrp(self.has_types, [rd.is_assumed for rd in self.rds], ch["x"], ch["q0"])]
        return list(zip(ss, coms[4:])) + [(1, bl), (t, m), (t * t % R, dm), (pow(t, 3, R), r)]

    def _bp_setup(self, q, cs, pub, coms, ch):
        return dict(basis=self.psv(q, cs, RPW()), init=self.init_open(coms, ch),
                    pub=self.psv(q, cs, pub), rounds=self.rounds())

    # -- verifier (TypedReciprocal.hs:447-467)
    def verify_rp(self, zk, coms):
        bl_com, r_com, dm_com, m_com = coms[:4]
        e, x, r0 = zk.oracle([dm_com, m_com] + coms[4:], 3)
        q, xq, r1 = zk.oracle([r_com], 3)
        q0 = self.q_powers(q, 1)[0]
        t = zk.oracle([bl_com], 1)[0]
        e_inv, q_inv, q0_inv = batch_inverse([e, q, q0])
        base_map = self.base_map(x)
        ph2s = make_phase2s(False, e, e_inv, x, base_map, self.ph1s_verifier())
        pub = make_public_consts(e, e_inv, x, xq, q0, q0_inv, t, self.has_types, self.rds, self.pub_vt, ph2s)
        cs = make_bp_coeffs(self.has_types, xq, r0, r1, t, make_shared_coeffs(e, e_inv, self.m_bases, base_map))
        ch = dict(e=e, x=x, r0=r0, q=q, xq=xq, r1=r1, q0=q0, t=t)
        return self._bp_setup(q, cs, pub, coms, ch)

    def decode_opening(self, nrm_scs, lin_scs):
        """`decodeProof'` tail (src/RangeProof.hs:78-80): makeNormLinearBP 1 [] nrm [] lin []."""
        return bp.PSV(0, self.G.zero, bp.NormLinear.make(self.arg, self.G, 1, [], nrm_scs, [], lin_scs, []))


# --------------------------------------------------------------------- Binary
class BinRangeData:
    """Binary.hs:37-54."""

    def __init__(self, char, mn, mx, is_output, is_assumed):
        if not (mx > mn and mx - mn < char):
            raise ValueError("Invalid range")
        n1 = integer_log(2, mx - mn - 1)
        self.min, self.max, self.is_output, self.is_assumed = mn, mx, is_output, is_assumed
        self.base_coeffs = [(mx - mn) - 2 ** n1] + [2 ** (n1 - i) for i in range(1, n1 + 1)]


def make_digits(rd, n):
    """Binary.hs:56-69."""
    if rd.is_assumed:
        return []
    n_adj = (n - rd.min) % R
    if not (0 <= n_adj < rd.max - rd.min):
        raise ValueError("witness out of range")
    bn = rd.base_coeffs[0]
    n1 = integer_log(2, rd.max - rd.min - 1)
    dn, n2 = (1, n_adj - bn) if n_adj > bn else (0, n_adj)
    bits = [int(c) for c in bin(n2)[2:]] if n2 else []
    return [dn] + [0] * (n1 - len(bits)) + bits


def input_coeffs_brp(cons, is_os, is_as, x):
    """Binary.hs:128-130."""
    return [((0 if a else x2) + (((-x) % R if o else x) if cons else 0)) % R
            for o, a, x2 in zip(is_os, is_as, powers1(x * x % R, len(is_os)))]


class SetupBRP:
    """`setupBRP` (Binary.hs:143-156)."""
    kind = "BRP"
    num_rp_coms = 2

    def __init__(self, G, arg, ps, cons, rds, net_pub):
        self.G, self.arg, self.cons, self.rds, self.net_pub = G, arg, cons, rds, net_pub
        self.nrm_len = sum(len(rd.base_coeffs) for rd in rds)
        self.lin_len = 2
        if len(ps) < 4 + self.nrm_len:
            raise ValueError("setup failed: not enough points")
        self.h, self.g, self.h0, self.h1 = ps[:4]
        self.hs = [self.h0, self.h1]
        self.gs = ps[4:4 + self.nrm_len]

    @staticmethod
    def points_needed(rds):
        return 4 + sum(len(rd.base_coeffs) for rd in rds)

    def q_powers(self, q, n):
        return bp.q_powers(self.arg, q, n)

    def com_terms(self, w):
        return rpw_terms(w, self.g, self.hs, self.gs, self.G)

    def psv(self, q, r, t, w):
        return bp.PSV(w.sc, self.g, bp.NormLinear.make(self.arg, self.G, q, [0, r * t % R], w.nrm, self.gs,
                                                        w.lin, self.hs))

    def info(self):
        return (2, self.nrm_len, 2)

    def witness(self, inputs):
        """inputs: [(value, blind)] (PedersenScalar, app/Main.hs:315).  Binary.hs:161-168."""
        vs = [v % R for v, _ in inputs]
        v_sum = sum(-v if rd.is_output else v for rd, v in zip(self.rds, vs))
        if not (self.cons and (self.net_pub + v_sum) % R == 0):
            raise ValueError("invalid witness (Binary.hs:165-167 needs conserved + balanced)")
        return dict(inputs=inputs, ds=[d for rd, v in zip(self.rds, vs) for d in make_digits(rd, v)])

    def public_consts(self, x, q0, q0_inv):
        """Binary.hs:73-97."""
        bss = []
        for xi, rd in zip(powers1(x * x % R, len(self.rds)), self.rds):
            if not rd.is_assumed:
                bss += [xi * (b % R) % R for b in rd.base_coeffs]
        mins = [0 if rd.is_assumed else rd.min % R for rd in self.rds]
        net = (-x) * self.net_pub % R if self.cons else 0
        z = (-2) * (net + sum(a * b for a, b in zip(mins, powers1(x * x % R, len(mins))))) % R
        half = inv(2)
        q2, q2i, s, nrm = q0, q0_inv, z, []
        for bx in bss:
            p = (-half + bx * q2i) % R
            s = (s + q2 * p * p) % R
            q2, q2i = q2 * q0 % R, q2i * q0_inv % R
            nrm.append(p)
        return RPW(s, [], nrm)

    def prove_rp(self, zk, wit, trace=None):
        """`proveBRPM` (Binary.hs:171-203)."""
        G = self.G
        n_wits = [RPW(v, [bl], []) for v, bl in wit["inputs"]]
        s_bl, l_bl0 = zk.random(), zk.random()
        d_wit = RPW(s_bl, [l_bl0, 0], wit["ds"])
        coms1 = G.msm_many([self.com_terms(w) for w in n_wits + [d_wit]])
        n_coms, d_com = coms1[:-1], coms1[-1]
        q, x, r = zk.oracle([d_com] + n_coms, 3)
        r_inv = inv(r)
        q0 = self.q_powers(q, 1)[0]
        q0_inv = inv(q0)
        pub_wit = self.public_consts(x, q0, q0_inv)
        bls_nrm = [zk.random() for _ in range(self.nrm_len)]
        bl_bl = zk.random()
        bl0, bl1, _ = make_poly_terms(self.q_powers(q, self.nrm_len), [bls_nrm, (d_wit + pub_wit).nrm])
        bl_wit = RPW(bl0, [bl_bl, r_inv * (s_bl - bl1) % R], bls_nrm)
        bl_com = G.msm(self.com_terms(bl_wit))
        t = zk.oracle([bl_com], 1)[0]
        coms = [bl_com, d_com] + n_coms
        is_os = [rd.is_output for rd in self.rds]
        is_as = [rd.is_assumed for rd in self.rds]
        bp_rounds = integer_log(2, self.nrm_len) - 1                   # Binary.hs:195 (prover rule)
        pub1 = RPW(t * pub_wit.sc, [], pub_wit.nrm)
        w1 = pub1 + d_wit + sum_v(w.scale(c) for c, w in zip(input_coeffs_brp(self.cons, is_os, is_as, x), n_wits)).scale(2 * t)
        bp_wit = self.psv(q, r, t, bl_wit + w1.scale(t))
        ch = dict(q=q, x=x, r=r, t=t, q0=q0)
        if trace is not None:
            trace.update(ch=ch, wits=dict(n=n_wits, d=d_wit, bl=bl_wit), pub=pub1.scale(t))
        setup = dict(basis=self.psv(q, r, t, RPW()), init=self.init_open(coms, ch),
                     pub=self.psv(q, r, t, pub1.scale(t)), rounds=bp_rounds)
        return coms, setup, bp_wit

    def init_open(self, coms, ch):
        """`TranscriptBRP.openWith` (Binary.hs:106-110)."""
        t = ch["t"]
        xs = [2 * t * t * c % R for c in input_coeffs_brp(self.cons, [rd.is_output for rd in self.rds],
                                                          [rd.is_assumed for rd in self.rds], ch["x"])]
        return list(zip(xs, coms[2:])) + [(1, coms[0]), (t, coms[1])]

    def verify_rp(self, zk, coms):
        """`verifyBRPM` (Binary.hs:205-220)."""
        q, x, r = zk.oracle([coms[1]] + coms[2:], 3)
        q0 = self.q_powers(q, 1)[0]
        q0_inv = inv(q0)
        t = zk.oracle([coms[0]], 1)[0]
        pw = self.public_consts(x, q0, q0_inv)
        pub = RPW(t * pw.sc, [], pw.nrm)
        ch = dict(q=q, x=x, r=r, t=t, q0=q0)
        rounds = bp.optimal_witness_size(self.arg, self.nrm_len, 2)[0]
        return dict(basis=self.psv(q, r, t, RPW()), init=self.init_open(coms, ch),
                    pub=self.psv(q, r, t, pub.scale(t)), rounds=rounds)

    def decode_opening(self, nrm_scs, lin_scs):
        return bp.PSV(0, self.G.zero, bp.NormLinear.make(self.arg, self.G, 1, [], nrm_scs, [], lin_scs, []))


# ------------------------------------------------------ RangeProof ZKP wrapper
def prove(setup, zk, wit, trace=None):
    """`RangeProof.proveM` (src/RangeProof.hs:95-97) -> dict(coms, responses (newest first),
    opening (final PSV))."""
    coms, bp_setup, bp_wit = setup.prove_rp(zk, wit, trace)
    rounds_trace = [] if trace is not None else None
    final, resps = bp.prove_bpm(setup.G, zk, bp_setup["rounds"], bp_wit, rounds_trace)
    if trace is not None:
        trace["rounds"] = rounds_trace
        trace["bp_wit"] = bp_wit
    return dict(coms=coms, responses=resps, opening=final)


def verify(setup, zk, proof):
    """`RangeProof.verifyM` (src/RangeProof.hs:99-101)."""
    s = setup.verify_rp(zk, proof["coms"])
    ok, _ = bp.verify_bpm(setup.G, zk, s["init"], proof["responses"], s["pub"], s["basis"], proof["opening"])
    return ok


# ----------------------------------------------------------- schema / witness
def approx_log_w(n):
    """app/Parse.hs:193-197."""
    l = integer_log(2, n)
    return l // integer_log(2, l)


def load_schema(schema, G, root_policy="exp", points=None):
    """app/Parse.hs:97-172 + the setup wiring of app/Main.hs:255-335.  `schema` is a dict or a
    path.  Returns the setup object (SetupTRRP | SetupBRP) with `.random_seed`."""
    if not isinstance(schema, dict):
        with open(schema) as f:
            schema = json.load(f)
    arg = {"ip": "IP", "innerproduct": "IP", "nl": "NL", "normlinear": "NL"}[schema.get("argument", "IP").lower()]
    typed, con, binary = schema.get("typed", False), schema.get("conserved", False), schema.get("binary", False)
    pubs = [(p["amount"], p.get("type", 0), p.get("isOutput", False)) for p in schema.get("public", [])]
    seed = schema.get("basisSeed", "test points")
    rds = []
    for r in schema["ranges"]:
        cnt, mn, mx = r.get("count", 1), r.get("min", 0), r.get("max", 2 ** 64)
        io, ia = r.get("isOutput", False), r.get("isAssumed", False)
        if binary:
            rds += [BinRangeData(R, mn, mx, io, ia) for _ in range(cnt)]
        else:
            base = r.get("base", approx_log_w(mx - mn))
            rds += [RangeData(R, base, mn, mx, r.get("isShared", False), io, ia) for _ in range(cnt)]
    if binary:
        need = SetupBRP.points_needed(rds)
        ps = points if points is not None else get_points(G, seed, need, root_policy)
        net = sum(-v if io else v for v, _, io in pubs)
        s = SetupBRP(G, arg, ps, con, rds, net)
    else:
        has_types = typed or con
        need = SetupTRRP.points_needed(has_types, rds)
        ps = points if points is not None else get_points(G, seed, need, root_policy)
        s = SetupTRRP(G, arg, ps, has_types, [(io, t, v) for v, t, io in pubs], rds)
    s.random_seed = schema.get("randomSeed", "default random seed")
    s.basis_seed = seed
    return s


def load_witness(setup, witness):
    """app/Main.hs:268-276: values, types, blinds (hash-derived unless given)."""
    if not isinstance(witness, list):
        with open(witness) as f:
            witness = json.load(f)
    gen = input_blinds(setup.random_seed, len(witness))
    vals = [(w["amount"] % R, w.get("type", 0) % R, (w["blind"] % R) if "blind" in w else gen[i])
            for i, w in enumerate(witness)]
    if setup.kind == "BRP":
        return setup.witness([(v, b) for v, _, b in vals])
    return setup.witness(vals)


def run_example(schema, witness, G, fmt="PrefixedP", trace=None):
    """prove + verify like `BulletproofsPP-exe prove` then `verify` (app/Main.hs:188-211)."""
    setup = load_schema(schema, G)
    wit = load_witness(setup, witness)
    proof = prove(setup, ZKPT(G, setup.random_seed, fmt), wit, trace)
    ok = verify(setup, ZKPT(G, None, fmt), proof)
    return setup, proof, ok
