"""Group back-ends for the oracle (test infrastructure only).

`Secp256k1` is the real group; values at the interface are affine tuples
``(x, y)`` or ``None`` for the identity (the reference's `A x y` / `O`).
`Toy` is F_r acting on itself -- the reference's `WrapV` fake group
(src/Utils.hs:117-133, src/Commitment.hs:87-92).

The reference's own MSM (`innerProduct`, src/Commitment.hs:325-335) is restated
literally in `straus_reference` (256-row bit-serial Straus over batch-normalised
bases with the *incomplete* projective mixed add of :156-169); `Secp256k1.msm`
is a fast, complete equivalent (same group element).  `oracle/c/` holds the C
restatement used for timing.
"""
from .field import Q, R, GX, GY, batch_inverse, reduce_scalar

_P = Q


# ---------------------------------------------------------------- Jacobian core
def _jdbl(P):
    X1, Y1, Z1 = P
    if Z1 == 0 or Y1 == 0:
        return (1, 1, 0)
    A = X1 * X1 % _P
    B = Y1 * Y1 % _P
    C = B * B % _P
    D = 2 * ((X1 + B) * (X1 + B) - A - C) % _P
    E = 3 * A % _P
    X3 = (E * E - 2 * D) % _P
    Y3 = (E * (D - X3) - 8 * C) % _P
    Z3 = 2 * Y1 * Z1 % _P
    return (X3, Y3, Z3)


def _jadd(P, Qp):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Qp
    if Z1 == 0:
        return Qp
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % _P
    Z2Z2 = Z2 * Z2 % _P
    U1 = X1 * Z2Z2 % _P
    U2 = X2 * Z1Z1 % _P
    S1 = Y1 * Z2 * Z2Z2 % _P
    S2 = Y2 * Z1 * Z1Z1 % _P
    H = (U2 - U1) % _P
    r = (S2 - S1) % _P
    if H == 0:
        return _jdbl(P) if r == 0 else (1, 1, 0)
    HH = H * H % _P
    HHH = H * HH % _P
    V = U1 * HH % _P
    X3 = (r * r - HHH - 2 * V) % _P
    Y3 = (r * (V - X3) - S1 * HHH) % _P
    Z3 = Z1 * Z2 * H % _P
    return (X3, Y3, Z3)


def _jmadd(P, A):
    """Jacobian + affine (complete)."""
    if A is None:
        return P
    X1, Y1, Z1 = P
    if Z1 == 0:
        return (A[0], A[1], 1)
    x2, y2 = A
    Z1Z1 = Z1 * Z1 % _P
    U2 = x2 * Z1Z1 % _P
    S2 = y2 * Z1 * Z1Z1 % _P
    H = (U2 - X1) % _P
    r = (S2 - Y1) % _P
    if H == 0:
        return _jdbl(P) if r == 0 else (1, 1, 0)
    HH = H * H % _P
    HHH = H * HH % _P
    V = X1 * HH % _P
    X3 = (r * r - HHH - 2 * V) % _P
    Y3 = (r * (V - X3) - Y1 * HHH) % _P
    Z3 = Z1 * H % _P
    return (X3, Y3, Z3)


def _to_affine(P):
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, -1, _P)
    zi2 = zi * zi % _P
    return (X * zi2 % _P, Y * zi2 * zi % _P)


def _to_affine_many(Ps):
    zi = batch_inverse([p[2] for p in Ps], _P)
    out = []
    for (X, Y, Z), z in zip(Ps, zi):
        if Z == 0:
            out.append(None)
        else:
            z2 = z * z % _P
            out.append((X * z2 % _P, Y * z2 * z % _P))
    return out


class Secp256k1:
    """secp256k1, y^2 = x^3 + 7.  Interface values are affine or None."""
    name = "secp256k1"
    order = R
    zero = None
    gen = (GX, GY)

    @staticmethod
    def on_curve(P):
        return P is None or (P[1] * P[1] - P[0] ** 3 - 7) % Q == 0

    @staticmethod
    def neg(P):
        return None if P is None else (P[0], (-P[1]) % Q)

    @staticmethod
    def add(A, B):
        if A is None:
            return B
        if B is None:
            return A
        return _to_affine(_jmadd((A[0], A[1], 1), B))

    @classmethod
    def sub(cls, A, B):
        return cls.add(A, cls.neg(B))

    @classmethod
    def mul(cls, s, P):
        return cls.msm([(s, P)])

    @staticmethod
    def msm_jac(pairs, c=None):
        """Pippenger over signed-reduced scalars; returns a Jacobian triple."""
        terms = []
        for s, P in pairs:
            s = reduce_scalar(s, R)
            if s == 0 or P is None:
                continue
            if s < 0:
                s, P = -s, (P[0], (-P[1]) % Q)
            terms.append((s, P))
        n = len(terms)
        if n == 0:
            return (1, 1, 0)
        if c is None:
            c = 2 if n < 4 else 4 if n < 32 else 6 if n < 256 else 8 if n < 4096 else 11
        nbits = max(s.bit_length() for s, _ in terms)
        acc = (1, 1, 0)
        mask = (1 << c) - 1
        for w in range((nbits + c - 1) // c - 1, -1, -1):
            for _ in range(c):
                acc = _jdbl(acc)
            buckets = [None] * (mask + 1)
            sh = w * c
            for s, P in terms:
                d = (s >> sh) & mask
                if d:
                    b = buckets[d]
                    buckets[d] = (P[0], P[1], 1) if b is None else _jmadd(b, P)
            run = (1, 1, 0)
            tot = (1, 1, 0)
            for d in range(mask, 0, -1):
                if buckets[d] is not None:
                    run = _jadd(run, buckets[d])
                tot = _jadd(tot, run)
            acc = _jadd(acc, tot)
        return acc

    @classmethod
    def msm(cls, pairs):
        """`commit`'s result as a group element (src/Commitment.hs:416-417)."""
        return _to_affine(cls.msm_jac(list(pairs)))

    @classmethod
    def msm_many(cls, list_of_pairs):
        return _to_affine_many([cls.msm_jac(list(p)) for p in list_of_pairs])

    @classmethod
    def pair_ip_many(cls, b, a, pairs):
        """`collapsePoints b a gL gR` (src/Bulletproof.hs:213-214) for a list of pairs."""
        return _to_affine_many([cls.msm_jac([(b, l), (a, r)]) for l, r in pairs])

    @staticmethod
    def coords(P):
        """Affine coordinates for the transcript (app/Main.hs:78-80)."""
        if P is None:
            raise ValueError("reference `coords` is partial: no case for O (app/Main.hs:79)")
        return P

    @staticmethod
    def lift_x(x, root_policy="exp"):
        """`pointX` (elliptic-curve-0.3.0, un-vendored): A x <$> sr(x^3+7).  q = 3 mod 4, so
        Tonelli-Shanks collapses to rhs^((q+1)/4).  RootPolicy: 'exp' (that value as is),
        'even', 'smaller'."""
        rhs = (x * x * x + 7) % Q
        y = pow(rhs, (Q + 1) // 4, Q)
        if y * y % Q != rhs:
            return None
        if root_policy == "even" and y & 1:
            y = Q - y
        elif root_policy == "smaller" and y > Q - y:
            y = Q - y
        return (x % Q, y)


def straus_reference(pairs):
    """Literal `innerProduct` (src/Commitment.hs:325-335) with `normalizeBasis`
    (:364-367: strip signs, bases affine) and the projective mixed add `nrmlAdd`
    (:156-169, incomplete: P+P gives z=0).  Returns affine or None."""
    sbs = []
    for s, P in pairs:
        n = reduce_scalar(s, R)
        if n < 0:
            n, P = -n, Secp256k1.neg(P)
        sbs.append((n, P))
    v = (0, 1, 0)  # projective identity

    def dbl(Pp):  # dbl-2007-bl, a = 0 (projective)
        X1, Y1, Z1 = Pp
        if Z1 == 0:
            return Pp
        XX = X1 * X1 % _P
        w = 3 * XX % _P
        s = 2 * Y1 * Z1 % _P
        ss = s * s % _P
        sss = s * ss % _P
        Rr = Y1 * s % _P
        RR = Rr * Rr % _P
        Bv = ((X1 + Rr) * (X1 + Rr) - XX - RR) % _P
        h = (w * w - 2 * Bv) % _P
        return (h * s % _P, (w * (Bv - h) - 2 * RR) % _P, sss)

    def nrml_add(A, Pp):
        if A is None:
            return Pp
        x2, y2 = A
        x1, y1, z1 = Pp
        if z1 == 0:
            return (x2, y2, 1)
        u = (y2 * z1 - y1) % _P
        uu = u * u % _P
        vv_ = (x2 * z1 - x1) % _P
        vv = vv_ * vv_ % _P
        vvv = vv_ * vv % _P
        r = vv * x1 % _P
        a = (uu * z1 - vvv - 2 * r) % _P
        return (vv_ * a % _P, (u * (r - a) - vvv * y1) % _P, vvv * z1 % _P)

    for row in range(255, -1, -1):
        v = dbl(v)
        for n, B in sbs:
            if (n >> row) & 1:
                v = nrml_add(B, v)
    X, Y, Z = v
    if Z == 0:
        return None
    zi = pow(Z, -1, _P)
    return (X * zi % _P, Y * zi % _P)


class SecpRef(Secp256k1):
    """secp256k1 with the reference's own algorithms for `commit` (256-row Straus, ref_msm) and
    `collapsePoints` (129-row pair product, ref_pair_ip), from oracle/c/ref_ec.c.  Used to time
    "the reference's CPU path as written"; results equal Secp256k1's (same group elements) unless
    the reference's incomplete mixed add hits P + P."""
    name = "secp256k1"
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            import ctypes
            import os
            path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libbppp_oracle.so")
            if not os.path.exists(path):
                raise RuntimeError("oracle C library not built: run `make -C oracle`")
            L = ctypes.CDLL(path)
            L.ref_msm.argtypes = [ctypes.c_size_t, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t,
                                  ctypes.c_char_p]
            L.ref_pair_ip.argtypes = [ctypes.c_size_t, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                      ctypes.c_char_p, ctypes.c_char_p]
            cls._lib = L
        return cls._lib

    @staticmethod
    def _pt(P):
        return bytes(64) if P is None else P[0].to_bytes(32, "little") + P[1].to_bytes(32, "little")

    @staticmethod
    def _unpt(b):
        x, y = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little")
        return None if x == 0 and y == 0 else (x, y)

    @classmethod
    def msm(cls, pairs):
        import ctypes
        pairs = list(pairs)
        if not pairs:
            return None
        nz = [(reduce_scalar(s, R), p) for s, p in pairs if s % R]
        n_pad = len(pairs) - len(nz)              # zero-scalar bases: normalised and walked, never added
        mags = b"".join(abs(s).to_bytes(32, "little") for s, _ in nz)
        neg = bytes(1 if s < 0 else 0 for s, _ in nz)
        out = ctypes.create_string_buffer(64)
        cls.lib().ref_msm(len(nz), mags, neg, b"".join(cls._pt(p) for _, p in nz), n_pad, out)
        return cls._unpt(out.raw)

    @classmethod
    def msm_many(cls, list_of_pairs):
        return [cls.msm(p) for p in list_of_pairs]

    @classmethod
    def pair_ip_many(cls, b, a, pairs):
        """[b*gL + a*gR for (gL, gR) in pairs] with signed integers a, b (|.| < 2^129)."""
        import ctypes
        n = len(pairs)
        out = ctypes.create_string_buffer(64 * max(n, 1))
        cls.lib().ref_pair_ip(n, abs(b).to_bytes(32, "little"), int(b < 0), abs(a).to_bytes(32, "little"), int(a < 0),
                              b"".join(cls._pt(l) + cls._pt(r) for l, r in pairs), out)
        return [cls._unpt(out.raw[64 * i:64 * i + 64]) for i in range(n)]


class SecpPip(SecpRef):
    """The same group with an honest CPU multi-scalar multiplication (Pippenger, signed windows;
    pip_msm in oracle/c/ref_ec.c) instead of the reference's 256-row Straus -- baseline (ii) of
    SURVEY 8(d).  Generator folds keep the reference's 129-row pair product."""

    @classmethod
    def msm(cls, pairs):
        import ctypes
        nz = [(reduce_scalar(s, R), p) for s, p in pairs if s % R and p is not None]
        if not nz:
            return None
        L = cls.lib()
        L.pip_msm.argtypes = [ctypes.c_size_t, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
        out = ctypes.create_string_buffer(64)
        L.pip_msm(len(nz), b"".join(abs(s).to_bytes(32, "little") for s, _ in nz), bytes(1 if s < 0 else 0 for s, _ in nz),
                  b"".join(cls._pt(p) for _, p in nz), out)
        return cls._unpt(out.raw)


class Toy:
    """F_r as a vector space over itself (WrapV, src/Utils.hs:117-133)."""
    name = "toy"
    order = R
    zero = 0
    gen = 1

    @staticmethod
    def neg(P):
        return (-P) % R

    @staticmethod
    def add(A, B):
        return (A + B) % R

    @staticmethod
    def sub(A, B):
        return (A - B) % R

    @staticmethod
    def mul(s, P):
        return s * P % R

    @staticmethod
    def msm(pairs):
        return sum(s * P for s, P in pairs) % R

    @classmethod
    def msm_many(cls, lp):
        return [cls.msm(p) for p in lp]

    @staticmethod
    def pair_ip_many(b, a, pairs):
        return [(b * l + a * r) % R for l, r in pairs]

    @staticmethod
    def coords(P):
        return (P, 0)

    @staticmethod
    def lift_x(x, root_policy="exp"):
        return x % R
