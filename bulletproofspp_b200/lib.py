"""ctypes binding of include/bppp_b200.h -- a mirror of the reference's operator interface
(`commit`/`innerProduct`, `collapsePoints`, the `BPOpening` methods of NormLinear) with the same
argument meaning; errors raise BpppError (the reference: Maybe/panic, app/Main.hs:155-169)."""
import ctypes as C
import os

ARG_NL, ARG_IP = 0, 1
_LIB = None

# every symbol include/bppp_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "bppp_init", "bppp_free", "bppp_last_error", "bppp_abi_version", "bppp_launch_count", "bppp_sync",
    "bppp_msm", "bppp_msm_batch", "bppp_pair_fold", "bppp_rational_reduce",
    "bppp_nl_create", "bppp_nl_round_commit", "bppp_nl_round_fold", "bppp_nl_lengths", "bppp_nl_final",
    "bppp_nl_destroy", "bppp_nl_verify", "bppp_dbg_field", "bppp_dbg_ec",
]


class BpppError(RuntimeError):
    pass


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libbppp_b200.so")


def load_library():
    """Load the in-tree CUDA library.  Fails loudly when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise BpppError("CUDA library not built: %s missing (run `python -c 'import __graft_entry__ as g; g.build()'`)" % path)
    lib = C.CDLL(path)
    vp, sz, u8p, ip = C.c_void_p, C.c_size_t, C.c_char_p, C.c_int
    lib.bppp_init.argtypes = [ip, C.POINTER(vp)]
    lib.bppp_free.argtypes = [vp]
    lib.bppp_free.restype = None
    lib.bppp_last_error.argtypes = [vp]
    lib.bppp_last_error.restype = C.c_char_p
    lib.bppp_abi_version.restype = ip
    lib.bppp_launch_count.argtypes = [vp]
    lib.bppp_launch_count.restype = C.c_uint64
    lib.bppp_sync.argtypes = [vp]
    lib.bppp_msm.argtypes = [vp, sz, u8p, u8p, u8p]
    lib.bppp_msm_batch.argtypes = [vp, sz, sz, u8p, u8p, ip, u8p]
    lib.bppp_pair_fold.argtypes = [vp, sz, u8p, ip, u8p, ip, u8p, u8p]
    lib.bppp_rational_reduce.argtypes = [u8p, u8p, C.POINTER(ip), u8p, C.POINTER(ip)]
    lib.bppp_nl_create.argtypes = [vp, ip, sz, sz, sz, u8p, u8p, u8p, u8p, u8p, u8p, u8p, u8p, C.POINTER(vp)]
    lib.bppp_nl_round_commit.argtypes = [vp, u8p, u8p]
    lib.bppp_nl_round_fold.argtypes = [vp, u8p]
    lib.bppp_nl_lengths.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
    lib.bppp_nl_final.argtypes = [vp, u8p, u8p, u8p]
    lib.bppp_nl_destroy.argtypes = [vp]
    lib.bppp_nl_destroy.restype = None
    lib.bppp_nl_verify.argtypes = [vp, ip, sz, sz, sz, sz, u8p, u8p, u8p, u8p, u8p, u8p, u8p, u8p, u8p, sz, sz,
                                   u8p, u8p, sz, u8p, u8p, C.POINTER(ip)]
    lib.bppp_dbg_field.argtypes = [vp, ip, sz, u8p, u8p, u8p]
    lib.bppp_dbg_ec.argtypes = [vp, ip, sz, u8p, u8p, u8p]
    _LIB = lib
    return lib


# ----------------------------------------------------------------- byte helpers
def int_to_le(x):
    return int(x).to_bytes(32, "little")


def le_to_int(b):
    return int.from_bytes(b, "little")


def ints_to_bytes(xs):
    return b"".join(int(x).to_bytes(32, "little") for x in xs)


def bytes_to_ints(b):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def point_to_bytes(p):
    """affine (x, y) or None (identity = 64 zero bytes)"""
    if p is None:
        return bytes(64)
    return int(p[0]).to_bytes(32, "little") + int(p[1]).to_bytes(32, "little")


def points_to_bytes(ps):
    return b"".join(point_to_bytes(p) for p in ps)


def bytes_to_point(b):
    x, y = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little")
    return None if x == 0 and y == 0 else (x, y)


def bytes_to_points(b):
    return [bytes_to_point(b[i:i + 64]) for i in range(0, len(b), 64)]


def _buf(n):
    return C.create_string_buffer(max(n, 1))


class Context:
    """A device context (one per GPU / host thread)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.bppp_init(device, C.byref(h))
        if rc != 0:
            raise BpppError("bppp_init(device=%d) failed with %d: no CUDA device, and there is no CPU fallback" % (device, rc))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.bppp_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise BpppError("%s failed (%d): %s" % (what, rc, self.lib.bppp_last_error(self.h).decode()))

    def launch_count(self):
        return int(self.lib.bppp_launch_count(self.h))

    def sync(self):
        self._ck(self.lib.bppp_sync(self.h), "bppp_sync")

    # -- `commit` / `innerProduct` (src/Commitment.hs:416-417, 325-335)
    def msm(self, pairs):
        """[(scalar, point)] -> point"""
        pairs = list(pairs)
        out = _buf(64)
        self._ck(self.lib.bppp_msm(self.h, len(pairs), ints_to_bytes(s for s, _ in pairs),
                                   points_to_bytes(p for _, p in pairs), out), "bppp_msm")
        return bytes_to_point(out.raw[:64])

    def msm_batch(self, scalars, points, shared_points=True):
        """scalars: [batch][n] ints; points: [n] (shared) or [batch][n] -> [batch] points"""
        batch, n = len(scalars), len(scalars[0]) if scalars else 0
        sb = b"".join(ints_to_bytes(row) for row in scalars)
        pb = points_to_bytes(points) if shared_points else b"".join(points_to_bytes(r) for r in points)
        out = _buf(64 * batch)
        self._ck(self.lib.bppp_msm_batch(self.h, batch, n, sb, pb, 1 if shared_points else 0, out), "bppp_msm_batch")
        return bytes_to_points(out.raw[:64 * batch])

    def msm_batch_raw(self, batch, n, scalars_bytes, points_bytes, shared_points=True):
        out = _buf(64 * batch)
        self._ck(self.lib.bppp_msm_batch(self.h, batch, n, scalars_bytes, points_bytes, 1 if shared_points else 0, out),
                 "bppp_msm_batch")
        return out.raw[:64 * batch]

    # -- `collapsePoints b a gL gR` over a vector (src/Bulletproof.hs:213-214)
    def pair_fold(self, a, b, points):
        """signed ints a, b; out[i] = b*points[2i] + a*points[2i+1]"""
        n = len(points)
        out = _buf(64 * ((n + 1) // 2))
        self._ck(self.lib.bppp_pair_fold(self.h, n, int_to_le(abs(a)), 1 if a < 0 else 0, int_to_le(abs(b)),
                                         1 if b < 0 else 0, points_to_bytes(points), out), "bppp_pair_fold")
        return bytes_to_points(out.raw[:64 * ((n + 1) // 2)])

    # -- `rationalReduceScalar` (src/Commitment.hs:242-255)
    def rational_reduce(self, x):
        a, b = _buf(32), _buf(32)
        an, bn = C.c_int(), C.c_int()
        rc = self.lib.bppp_rational_reduce(int_to_le(x), a, C.byref(an), b, C.byref(bn))
        if rc:
            raise BpppError("bppp_rational_reduce failed (%d)" % rc)
        av, bv = le_to_int(a.raw[:32]), le_to_int(b.raw[:32])
        return (-av if an.value else av), (-bv if bn.value else bv)

    def dbg_field(self, op, a, b):
        n = len(a)
        out = _buf(32 * n)
        self._ck(self.lib.bppp_dbg_field(self.h, op, n, ints_to_bytes(a), ints_to_bytes(b), out), "bppp_dbg_field")
        return bytes_to_ints(out.raw[:32 * n])

    def dbg_ec(self, op, a, b):
        n = len(a)
        out = _buf(64 * n)
        self._ck(self.lib.bppp_dbg_ec(self.h, op, n, points_to_bytes(a), points_to_bytes(b), out), "bppp_dbg_ec")
        return bytes_to_points(out.raw[:64 * n])

    # -- verifyBPM's collapsed MSM check (src/Bulletproof.hs:370-378)
    def nl_verify(self, kind, g, G, H, q, s_pub, pub_w, c, es, XR, fw, fl, init):
        """Per-proof lists: q[b], s_pub[b], pub_w[b][N], c[b][M], es[b][k] (newest first),
        XR[b][k] = (X, R) newest first, fw[b][n_norm], fl[b][n_lin], init[b] = [(scalar, point)]."""
        B, N, M = len(q), len(G), len(H)
        k = len(es[0])
        n_norm, n_lin, n_init = len(fw[0]), len(fl[0]), len(init[0])
        ok = (C.c_int * B)()
        cat = lambda rows: b"".join(ints_to_bytes(r) for r in rows)
        self._ck(self.lib.bppp_nl_verify(
            self.h, kind, B, N, M, k, point_to_bytes(g), points_to_bytes(G), points_to_bytes(H), ints_to_bytes(q),
            ints_to_bytes(s_pub), cat(pub_w), cat(c), cat(es),
            b"".join(point_to_bytes(x) + point_to_bytes(r) for row in XR for x, r in row), n_norm, n_lin, cat(fw),
            cat(fl), n_init, cat([[s for s, _ in row] for row in init]),
            b"".join(points_to_bytes(p for _, p in row) for row in init), ok), "bppp_nl_verify")
        return [bool(v) for v in ok]


class NormLinearArgument:
    """Device-resident NormLinear argument state over a batch of proofs: the `BPOpening` methods
    of NL.NormLinear (src/Bulletproof/NormArgument.hs:153-178) behind `proveRoundM`
    (src/Bulletproof.hs:346-355).  The transcript stays with the caller."""

    def __init__(self, ctx, kind, g, G, H, q, s, w, l, c):
        """g, G[N], H[M]: shared generators; per proof lists q[b], s[b], w[b][N], l[b][M], c[b][M]."""
        self.ctx, self.B, self.N, self.M = ctx, len(q), len(G), len(H)
        h = C.c_void_p()
        cat = lambda rows: b"".join(ints_to_bytes(r) for r in rows)
        ctx._ck(ctx.lib.bppp_nl_create(ctx.h, kind, self.B, self.N, self.M, point_to_bytes(g), points_to_bytes(G),
                                       points_to_bytes(H), ints_to_bytes(q), ints_to_bytes(s), cat(w), cat(l), cat(c),
                                       C.byref(h)), "bppp_nl_create")
        self.h = h

    @classmethod
    def from_bytes(cls, ctx, kind, B, N, M, g, G, H, q, s, w, l, c):
        self = cls.__new__(cls)
        self.ctx, self.B, self.N, self.M = ctx, B, N, M
        h = C.c_void_p()
        ctx._ck(ctx.lib.bppp_nl_create(ctx.h, kind, B, N, M, g, G, H, q, s, w, l, c, C.byref(h)), "bppp_nl_create")
        self.h = h
        return self

    def round_commit_raw(self):
        X, R = _buf(64 * self.B), _buf(64 * self.B)
        self.ctx._ck(self.ctx.lib.bppp_nl_round_commit(self.h, X, R), "bppp_nl_round_commit")
        return X.raw[:64 * self.B], R.raw[:64 * self.B]

    def round_commit(self):
        X, R = self.round_commit_raw()
        return bytes_to_points(X), bytes_to_points(R)

    def round_fold(self, es):
        eb = es if isinstance(es, (bytes, bytearray)) else ints_to_bytes(es)
        self.ctx._ck(self.ctx.lib.bppp_nl_round_fold(self.h, eb), "bppp_nl_round_fold")

    def lengths(self):
        a, b = C.c_size_t(), C.c_size_t()
        self.ctx._ck(self.ctx.lib.bppp_nl_lengths(self.h, C.byref(a), C.byref(b)), "bppp_nl_lengths")
        return a.value, b.value

    def final(self):
        n, m = self.lengths()
        s, w, l = _buf(32 * self.B), _buf(32 * self.B * n), _buf(32 * self.B * m)
        self.ctx._ck(self.ctx.lib.bppp_nl_final(self.h, s, w, l), "bppp_nl_final")
        ws, ls = bytes_to_ints(w.raw[:32 * self.B * n]), bytes_to_ints(l.raw[:32 * self.B * m])
        return (bytes_to_ints(s.raw[:32 * self.B]), [ws[b * n:(b + 1) * n] for b in range(self.B)],
                [ls[b * m:(b + 1) * m] for b in range(self.B)])

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.bppp_nl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
