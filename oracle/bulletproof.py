"""The recursive norm / weighted-inner-product argument (oracle; test infrastructure only).

Literal *stored-form* restatement (lazy normalisation `nrmlz''`, half-length
`rationalReduceScalar` generator folds) of
  src/Bulletproof.hs                      (driver, BPCollection [] instance, round counting)
  src/Bulletproof/NormArgument.hs         (NL: Norm, Linear, NormLinear)
  src/Bulletproof/InnerProductArgument.hs (IP: InnerProduct, Linear, Norm, NormLinear)
generic over a group back-end `G` (oracle.curve.Secp256k1 or Toy).

Vectors are Python lists (the reference's `ArgColl = []`, app/Main.hs:101).
"""
from .field import R, inv, batch_inverse, rational_reduce_scalar, powers1


# ------------------------------------------------------------ BPCollection []
def halves(xs, d):
    """Adjacent pairs (x_2i, x_2i+1), odd tail padded with `d` (src/Bulletproof.hs:77-90)."""
    return [(xs[i], xs[i + 1] if i + 1 < len(xs) else d) for i in range(0, len(xs), 2)]


def tensor(bs, es, qs):
    """`tensor'` for lists (src/Bulletproof.hs:94-95): `es` newest-first, `qs` in round order."""
    ts = [1]
    for j, e in enumerate(reversed(es)):
        q = qs[j]
        ts = [q * t % R for t in ts] + [e * t % R for t in ts]
    return [b * t % R for b in bs for t in ts]


def contract(xs, ys):
    """`contract'` (src/Bulletproof.hs:97): dot xs with consecutive chunks of ys."""
    n = len(xs)
    return [sum(x * y for x, y in zip(xs, ys[i:i + n])) % R for i in range(0, len(ys), n)]


def round_reduce(n):
    return n // 2 + n % 2


def round_reduce_by(n, k):
    for _ in range(k):
        n = round_reduce(n)
    return n


def number_rounds_reduce(n):
    """src/Bulletproof.hs:300-304."""
    r = 0
    while n >= 5:
        n = round_reduce(n)
        r += 1
    return r, n


def number_rounds_reduce2(n):
    """`numberRoundsReduce'` (src/Bulletproof.hs:306-308)."""
    r, n2 = number_rounds_reduce(n)
    return (r + 1, round_reduce(n2)) if n2 > 2 else (r, n2)


def _iter_sq(q, n):
    out = []
    for _ in range(n):
        out.append(q)
        q = q * q % R
    return out


# ------------------------------------------------------------------ NL: Linear
class NLLinear:
    """`Linear` of NormArgument.hs:34-81: frame normalisation n, entries (c, x, g)."""
    make_es = staticmethod(lambda e: (e % R, (e * e - 1) % R))

    def __init__(self, G, cs, xs, gs, n=1):
        L = max(len(cs), len(xs), len(gs))
        self.G, self.n = G, n % R
        self.cs = list(cs) + [0] * (L - len(cs))
        self.xs = list(xs) + [0] * (L - len(xs))
        self.gs = list(gs) + [G.zero] * (L - len(gs))

    def opening(self):
        return list(zip(self.xs, self.gs))

    def eval_scalar(self):
        return sum(c * x for c, x in zip(self.cs, self.xs)) % R

    def get_witness(self):
        return [self.n * x % R for x in self.xs]

    def _pairs(self):
        Z = self.G.zero
        return halves(list(zip(self.cs, self.xs, self.gs)), (0, 0, Z))

    def make_scalars_coms(self):
        """NormArgument.hs:56-59."""
        sX = sR = 0
        oX, oR = [], []
        for (cL, xL, gL), (cR, xR, gR) in self._pairs():
            sX += cL * xR + cR * xL
            sR += cR * xR
            oX += [(xR, gL), (xL, gR)]
            oR += [(xR, gR)]
        return sX % R, oX, sR % R, oR

    def _ratio(self, e):
        return e

    def collapse(self, e):
        """NormArgument.hs:64-71."""
        a, b = rational_reduce_scalar(self._ratio(e))
        a0, b0 = a % R, b % R
        b0i = inv(b0)
        cs, xs, gp = [], [], []
        for (cL, xL, gL), (cR, xR, gR) in self._pairs():
            cs.append((b0 * cL + a0 * cR) % R)
            xs.append((b0i * xL + e * b0i % R * xR) % R)
            gp.append((gL, gR))
        gs = self.G.pair_ip_many(b, a, gp)                  # collapsePoints b' a' gL gR
        return type(self)(self.G, cs, xs, gs, self.n * b0)

    def _tensor_es(self, es):
        return es

    def expand_challenges(self, es, pub, basis):
        """NormArgument.hs:73-81 -> (sc, verifier opening over the original generators)."""
        es = self._tensor_es(es)
        ones = [1] * len(es)
        exp_es = tensor([1], es, ones)
        cs2 = contract(exp_es, pub.cs)
        vs = [self.n * x % R for x in self.xs]
        sc = sum(c * v for c, v in zip(cs2, vs)) % R
        ts = tensor(vs, es, ones)
        out = []
        for i, (p, g) in enumerate(zip(pub.xs, basis.gs)):
            out.append(((p - (ts[i] if i < len(ts) else 0)) % R, g))
        return sc, out


# -------------------------------------------------------------------- NL: Norm
class NLNorm:
    """`Norm` of NormArgument.hs:86-148."""
    make_es = staticmethod(lambda e: (e % R, (e * e - 1) % R))

    def __init__(self, G, q, xs, gs, n=1, q_inv=None):
        L = max(len(xs), len(gs))
        self.G, self.q, self.n = G, q % R, n % R
        self.q_inv = inv(q) if q_inv is None else q_inv
        self.xs = list(xs) + [0] * (L - len(xs))
        self.gs = list(gs) + [G.zero] * (L - len(gs))

    def opening(self):
        return list(zip(self.xs, self.gs))

    def eval_scalar(self):
        """:110-111  n^2 * sum (q^2)^(i+1) x_i^2."""
        ws = powers1(self.q * self.q % R, len(self.xs))
        return self.n * self.n % R * sum(w * x * x for w, x in zip(ws, self.xs)) % R

    def get_witness(self):
        return [self.n * x % R for x in self.xs]

    def make_scalars_coms(self):
        """:113-118."""
        q, qi, n = self.q, self.q_inv, self.n
        q4 = pow(q, 4, R)
        s, sX, sR = 1, 0, 0
        oX, oR = [], []
        for (xL, gL), (xR, gR) in halves(list(zip(self.xs, self.gs)), (0, self.G.zero)):
            sX += s * xL % R * xR
            sR += s * xR % R * xR
            oX += [(q * xR % R, gL), (qi * xL % R, gR)]
            oR += [(xR, gR)]
            s = s * q4 % R
        return (2 * n * n * pow(q, 3, R) * sX) % R, oX, (n * n * q4 * sR) % R, oR

    def collapse(self, e):
        """:123-129."""
        q, qi = self.q, self.q_inv
        a, b = rational_reduce_scalar(e * qi % R)
        b0 = b % R
        b0i = inv(b0)
        xs, gp = [], []
        for (xL, gL), (xR, gR) in halves(list(zip(self.xs, self.gs)), (0, self.G.zero)):
            xs.append((b0i * xL + e * q % R * b0i % R * xR) % R)
            gp.append((gL, gR))
        gs = self.G.pair_ip_many(b, a, gp)
        return NLNorm(self.G, q * q % R, xs, gs, self.n * b0 % R * qi, qi * qi % R)

    def expand_challenges(self, es, pub, basis):
        """:131-145."""
        q = pub.q
        vs = [self.n * x % R for x in self.xs]
        qF = q
        for _ in es:
            qF = qF * qF % R
        ws = powers1(qF * qF % R, len(vs))
        sc = sum(w * v * v for w, v in zip(ws, vs)) % R
        ts = tensor(vs, es, _iter_sq(q, len(es)))
        out = []
        for i, (p, g) in enumerate(zip(pub.xs, basis.gs)):
            out.append(((p - (ts[i] if i < len(ts) else 0)) % R, g))
        return sc, out


# -------------------------------------------------------------- IP: InnerProduct
class IPInner:
    """`InnerProduct` of InnerProductArgument.hs:32-127: entries (x, g, y, h), two
    normalisations nx (frame) and ny, scale s."""
    make_es = staticmethod(lambda e: (inv(e), e % R))

    def __init__(self, G, s, q, body, nx=1, ny=1, q_inv=None):
        self.G, self.s, self.q, self.nx, self.ny = G, s % R, q % R, nx % R, ny % R
        self.q_inv = inv(q) if q_inv is None else q_inv
        self.body = list(body)           # [(x, g, y, h)]

    def opening(self):
        out = []
        for x, g, y, h in self.body:
            out += [(x, g), (y, h)]
        return out

    def eval_scalar(self):
        """:60-63."""
        ws = powers1(self.q, len(self.body))
        return self.s * self.nx % R * self.ny % R * sum(
            x * y % R * w for (x, _, y, _), w in zip(self.body, ws)) % R

    def _pairs(self):
        return halves(self.body, (0, self.G.zero, 0, self.G.zero))

    def make_scalars_coms(self):
        """:70-81.  Returned frames carry (t*nx) but commitments ignore normalisation."""
        q, qi = self.q, self.q_inv
        q2 = q * q % R
        w, sL, sR = 1, 0, 0
        oL, oR = [], []
        for (xL, gL, yL, hL), (xR, gR, yR, hR) in self._pairs():
            sL += w * xL % R * yR
            sR += w * xR % R * yL
            oL += [(qi * xL % R, gR), (yR, hL)]
            oR += [(q * xR % R, gL), (yL, hR)]
            w = w * q2 % R
        k = self.s * self.nx % R * self.ny % R
        return (k * q % R * sL) % R, oL, (k * q2 % R * sR) % R, oR

    def collapse(self, e):
        """:86-101."""
        q, qi = self.q, self.q_inv
        ei = inv(e)
        a, b = rational_reduce_scalar(qi * ei % R)
        b0i = inv(b % R)
        c, d = rational_reduce_scalar(e)
        d0i = inv(d % R)
        prs = self._pairs()
        gs = self.G.pair_ip_many(b, a, [(L[1], Rr[1]) for L, Rr in prs])
        hs = self.G.pair_ip_many(d, c, [(L[3], Rr[3]) for L, Rr in prs])
        body = []
        for ((xL, gL, yL, hL), (xR, gR, yR, hR)), g2, h2 in zip(prs, gs, hs):
            body.append(((b0i * (xL + e * q % R * xR)) % R, g2, (d0i * (yL + ei * yR)) % R, h2))
        return IPInner(self.G, self.s, q * q % R, body, self.nx * (b % R) % R * qi,
                       self.ny * (d % R), qi * qi % R)

    def expand_challenges(self, esY, pub, basis):
        """:103-124."""
        q = pub.q
        s = pub.s
        qF = q
        for _ in esY:
            qF = qF * qF % R
        esX = [inv(e) for e in esY]
        vsX = [self.nx * b[0] % R for b in self.body]
        vsY = [self.ny * b[2] % R for b in self.body]
        ws = powers1(qF, len(vsX))
        sc = s * sum(w * x % R * y for w, x, y in zip(ws, vsX, vsY)) % R
        tsX = tensor(vsX, esX, _iter_sq(q, len(esY)))
        tsY = tensor(vsY, esY, [1] * len(esY))
        out = []
        for i, ((pX, _, pY, _), (_, g, _, h)) in enumerate(zip(pub.body, basis.body)):
            eX = tsX[i] if i < len(tsX) else 0
            eY = tsY[i] if i < len(tsY) else 0
            out += [((pX - eX) % R, g), ((pY - eY) % R, h)]
        return sc, out


class IPNorm(IPInner):
    """`Norm` of InnerProductArgument.hs:190-231: inner product after a basis change."""

    @classmethod
    def make(cls, G, r, ss, gs):
        """`makeNorm r` (:194-206): q = r^4, pairs (s0,g0),(s1,g1) -> x = s0/2r + s1/2,
        y = -s0/2r + s1/2, g' = g1 + r*g0, h' = g1 - r*g0."""
        L = max(len(ss), len(gs))
        ss = list(ss) + [0] * (L - len(ss))
        gs = list(gs) + [G.zero] * (L - len(gs))
        half = inv(2)
        r2i = inv(2 * r % R) if r % R else 0
        body = []
        for (s0, g0), (s1, g1) in halves(list(zip(ss, gs)), (0, G.zero)):
            p = G.mul(r, g0)
            body.append(((r2i * s0 + half * s1) % R, G.add(g1, p),
                         (-r2i * s0 + half * s1) % R, G.sub(g1, p)))
        return cls(G, 4, pow(r, 4, R), body)

    def collapse(self, e):
        c = IPInner.collapse(self, e)
        c.__class__ = IPNorm
        return c

    def get_witness(self):
        """:222-223  (nx*x - ny*y, nx*x + ny*y) per element."""
        out = []
        for x, _, y, _ in self.body:
            out += [(self.nx * x - self.ny * y) % R, (self.nx * x + self.ny * y) % R]
        return out


class IPLinear(NLLinear):
    """`Linear` of InnerProductArgument.hs:132-181."""
    make_es = staticmethod(lambda e: (inv(e), e % R))

    def make_scalars_coms(self):
        """:149-152."""
        sL = sR = 0
        oL, oR = [], []
        for (cL, xL, gL), (cR, xR, gR) in self._pairs():
            sL += cR * xL
            sR += cL * xR
            oL += [(xL, gR)]
            oR += [(xR, gL)]
        return sL % R, oL, sR % R, oR

    def _ratio(self, e):
        return inv(e)                       # :158  rationalReduceScalar (recip e)

    def _tensor_es(self, es):
        return [inv(e) for e in es]         # :172


# ------------------------------------------------------------------ NormLinear
class NormLinear:
    """`NormLinear` = BPCompose (Norm) (Linear) with scalar `s` (src/Bulletproof.hs:225-273;
    NormArgument.hs:153-178; InnerProductArgument.hs:239-267).  kind: 'NL' | 'IP'."""

    def __init__(self, kind, G, s, norm, lin):
        self.kind, self.G, self.s, self.norm, self.lin = kind, G, s % R, norm, lin

    @classmethod
    def make(cls, kind, G, q, cs, nrm, gs, lin, hs, s=1):
        """`makeNormLinearBP' s q cs nss ngs lss lgs` (NormArgument.hs:162; InnerProductArgument.hs:248)."""
        if kind == "NL":
            return cls(kind, G, s, NLNorm(G, q, nrm, gs), NLLinear(G, cs, lin, hs))
        return cls(kind, G, s, IPNorm.make(G, q, nrm, gs), IPLinear(G, cs, lin, hs))

    def make_es(self, e):
        return self.norm.make_es(e)

    def opening(self):
        return self.norm.opening() + self.lin.opening()

    def eval_scalar(self):
        return self.s * (self.norm.eval_scalar() + self.lin.eval_scalar()) % R

    def make_scalars_coms(self):
        a = self.norm.make_scalars_coms()
        b = self.lin.make_scalars_coms()
        return (a[0] + b[0]) % R, a[1] + b[1], (a[2] + b[2]) % R, a[3] + b[3]

    def get_witness(self):
        return [self.s * w % R for w in self.norm.get_witness() + self.lin.get_witness()]

    def collapse(self, e):
        return NormLinear(self.kind, self.G, self.s, self.norm.collapse(e), self.lin.collapse(e))

    def expand_challenges(self, es, pub, basis):
        sa, oa = self.norm.expand_challenges(es, pub.norm, basis.norm)
        sb, ob = self.lin.expand_challenges(es, pub.lin, basis.lin)
        return (sa + sb) % R, oa + ob

    def lengths(self):
        if self.kind == "NL":
            return len(self.norm.xs), len(self.lin.xs)
        return len(self.norm.body), len(self.lin.xs)


def q_powers(kind, q, n):
    """`qPowers'` of the Weighted instances: NL `powers' (q^2)` (NormArgument.hs:147-148);
    IP norm `powers' (-(q^2))` (InnerProductArgument.hs:230-231)."""
    return powers1(q * q % R if kind == "NL" else (-q * q) % R, n)


def optimal_witness_size(kind, n_len, l_len):
    """NormArgument.hs:165-178 / InnerProductArgument.hs:253-267 -> (rounds, (nrm, lin))."""
    if kind == "NL":
        nR, n1 = number_rounds_reduce(n_len)
        lR, l1 = number_rounds_reduce(l_len)
        r = max(nR, lR)
        n2, l2 = round_reduce_by(n1, r - nR), round_reduce_by(l1, r - lR)
        if n2 + l2 > 5:
            return r + 1, (round_reduce(n2), round_reduce(l2))
        return r, (n2, l2)
    n_even = (n_len + n_len % 2) // 2
    nR, n1 = number_rounds_reduce2(n_even)
    lR, l1 = number_rounds_reduce(l_len)
    r = max(nR, lR)
    n2, l2 = round_reduce_by(n1, r - nR), round_reduce_by(l1, r - lR)
    if 2 * n2 + l2 > 5:
        return r + 1, (2 * round_reduce(n2), round_reduce(l2))
    return r, (2 * n2, l2)


# --------------------------------------------------------- prover / verifier
class PSV:
    """`PedersenScalarVector` (src/Commitment.hs:487-500): scalar `s` on generator `g` + vector."""

    def __init__(self, s, g, vec):
        self.s, self.g, self.vec = s % R, g, vec


def prove_round(G, zk, com, trace=None):
    """`proveRoundM` (src/Bulletproof.hs:346-355)."""
    c = com.vec
    as_, oa, bs_, ob = c.make_scalars_coms()
    ac, bc = G.msm_many([[(as_, com.g)] + oa, [(bs_, com.g)] + ob])
    e = zk.oracle([ac, bc])[0]
    e0, e1 = c.make_es(e)
    sc = (com.s + e0 * as_ + e1 * bs_) % R
    if trace is not None:
        trace.append(dict(sX=as_, sR=bs_, X=ac, R=bc, e=e))
    return PSV(sc, com.g, c.collapse(e)), (ac, bc)


def prove_bpm(G, zk, n, com, trace=None):
    """`proveBPM` (src/Bulletproof.hs:357-359): responses NEWEST FIRST."""
    resps = []
    for _ in range(n):
        com, xr = prove_round(G, zk, com, trace)
        resps.insert(0, xr)
    return com, resps


def verify_bpm(G, zk, init_open, rs, pub, basis, opening):
    """`verifyBPM` (src/Bulletproof.hs:370-378).  `init_open` = `openToList initCom`;
    `pub`, `basis`, `opening` are PSVs; returns (ok, the single MSM's (scalar, point) list)."""
    es = []
    for a, b in reversed(rs):                       # foldrM: oldest round first
        es.insert(0, zk.oracle([a, b])[0])
    sc, chs = opening.vec.expand_challenges(es, pub.vec, basis.vec)
    terms = [((pub.s - sc) % R, basis.g)] + chs + list(init_open)
    for e, (a, b) in zip(es, rs):                   # verifyWith (:362-368)
        e0, e1 = opening.vec.make_es(e)
        terms += [(e0, a), (e1, b)]
    return G.msm(terms) == G.zero, terms
