"""ctypes binding of include/bppp_b200.h -- a mirror of the reference's operator interface
(`commit`/`innerProduct`, `collapsePoints`, the `BPOpening` methods of NormLinear) with the same
argument meaning; errors raise BpppError (the reference: Maybe/panic, app/Main.hs:155-169)."""
import ctypes as C
import os

ARG_NL, ARG_IP = 0, 1
_LIB = None

# every symbol include/bppp_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "bppp_init", "bppp_free", "bppp_last_error", "bppp_tune_process", "bppp_abi_version", "bppp_launch_count", "bppp_sync",
    "bppp_msm", "bppp_msm_batch", "bppp_pair_fold", "bppp_rational_reduce",
    "bppp_nl_create", "bppp_nl_round_commit", "bppp_nl_round_fold", "bppp_nl_lengths", "bppp_nl_final",
    "bppp_nl_destroy", "bppp_nl_verify", "bppp_dbg_field", "bppp_dbg_ec",
    "bppp_fb_create", "bppp_fb_msm_batch", "bppp_fb_destroy",
    "bppp_rp_setup", "bppp_rp_free", "bppp_rp_last_error", "bppp_rp_info", "bppp_rp_points", "bppp_input_blind",
    "bppp_set_host_threads", "bppp_rp_prove_batch", "bppp_rp_verify_batch",
    "bppp_host_sha256", "bppp_host_oracle", "bppp_host_fr", "bppp_host_get_points", "bppp_host_transcript", "bppp_host_scheduler_selftest",
    "bppp_profile_enable", "bppp_profile_reset", "bppp_profile_report", "bppp_timer_start", "bppp_timer_stop",
    "bppp_measure_imad_peak", "bppp_gens_create", "bppp_gens_destroy", "bppp_gens_msm_batch",
    "bppp_set_device_host_threads", "bppp_nl_create_gens", "bppp_nl_verify_gens",
    "bppp_set_thread_host_threads", "bppp_ctx_device", "bppp_rp_contexts", "bppp_pinned_alloc", "bppp_pinned_free", "bppp_nl_set_shard", "bppp_nl_export", "bppp_rp_encoded_sizes", "bppp_rp_encode_batch", "bppp_rp_decode_batch",
    "bppp_nl_prove", "bppp_nl_challenges", "bppp_get_points", "bppp_dtr_create", "bppp_dtr_destroy", "bppp_dtr_reset",
    "bppp_dtr_oracle", "bppp_dev_random",
    "bppp_trrp_create", "bppp_trrp_destroy", "bppp_trrp_phase1", "bppp_trrp_phase2", "bppp_trrp_phase3", "bppp_trrp_commit_bl",
    "bppp_trrp_phase4", "bppp_nl_create_trrp", "bppp_trrp_verify_pub", "bppp_nl_verify_trrp",
    "bppp_dtr_absorb", "bppp_dtr_squeeze", "bppp_dtr_fits", "bppp_dtr_export", "bppp_rp_set_device_transcript", "bppp_nl_round_challenge",
    "bppp_nl_attach_transcript", "bppp_nl_prove_device", "bppp_rp_set_batch_verify", "bppp_nl_verify_trrp_rlc", "bppp_nl_verify_gens_rlc",
    "bppp_dtr_squeeze_inv", "bppp_trrp_want_inverses", "bppp_trrp_set_shared", "bppp_trrp_shared_coeffs",
    "bppp_gens_enable_lut", "bppp_rp_enable_lut",
    "bppp_comm_load", "bppp_comm_last_error", "bppp_comm_unique_id", "bppp_comm_create", "bppp_comm_destroy", "bppp_nl_prove_sharded",
    "bppp_trrp_set_transcript", "bppp_trrp_phase1_tr", "bppp_trrp_phase2_tr", "bppp_trrp_phase3_rnd", "bppp_trrp_commit_bl_tr",
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
    lib.bppp_nl_set_shard.argtypes = [vp, sz]
    lib.bppp_nl_export.argtypes = [vp, u8p, u8p, u8p, u8p]
    lib.bppp_nl_destroy.argtypes = [vp]
    lib.bppp_nl_destroy.restype = None
    lib.bppp_nl_verify.argtypes = [vp, ip, sz, sz, sz, sz, u8p, u8p, u8p, u8p, u8p, u8p, u8p, u8p, u8p, sz, sz,
                                   u8p, u8p, sz, u8p, u8p, C.POINTER(ip)]
    lib.bppp_dbg_field.argtypes = [vp, ip, sz, u8p, u8p, u8p]
    lib.bppp_dbg_ec.argtypes = [vp, ip, sz, u8p, u8p, u8p]
    lib.bppp_profile_enable.argtypes = [vp, ip]
    lib.bppp_profile_reset.argtypes = [vp]
    lib.bppp_profile_report.argtypes = [vp, C.c_char_p, sz]
    lib.bppp_timer_start.argtypes = [vp]
    lib.bppp_timer_stop.argtypes = [vp, C.POINTER(C.c_double)]
    lib.bppp_measure_imad_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.bppp_gens_create.argtypes = [vp, sz, sz, u8p, u8p, u8p, C.POINTER(vp)]
    lib.bppp_gens_destroy.argtypes = [vp]
    lib.bppp_gens_destroy.restype = None
    lib.bppp_gens_msm_batch.argtypes = [vp, sz, sz, u8p, u8p]
    lib.bppp_set_device_host_threads.argtypes = [ip]
    lib.bppp_set_device_host_threads.restype = None
    lib.bppp_nl_create_gens.argtypes = [vp, ip, sz, u8p, u8p, u8p, u8p, u8p, C.POINTER(vp)]
    lib.bppp_nl_verify_gens.argtypes = [vp, ip, sz, sz, u8p, u8p, u8p, u8p, u8p, u8p, sz, sz, u8p, u8p, sz, u8p, u8p,
                                        C.POINTER(ip)]
    lib.bppp_rp_contexts.argtypes = [vp, C.POINTER(vp), sz, C.POINTER(sz)]
    lib.bppp_ctx_device.argtypes = [vp]
    lib.bppp_rp_encoded_sizes.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
    lib.bppp_rp_encode_batch.argtypes = [vp, sz, u8p, u8p, u8p, u8p, u8p]
    lib.bppp_rp_decode_batch.argtypes = [vp, sz, u8p, u8p, u8p, u8p, u8p, C.POINTER(ip)]
    lib.bppp_pinned_alloc.argtypes = [sz, C.POINTER(vp)]
    lib.bppp_pinned_free.argtypes = [vp]
    lib.bppp_pinned_free.restype = None
    lib.bppp_fb_create.argtypes = [vp, sz, u8p, C.POINTER(vp)]
    lib.bppp_fb_msm_batch.argtypes = [vp, sz, u8p, u8p]
    lib.bppp_fb_destroy.argtypes = [vp]
    lib.bppp_fb_destroy.restype = None
    lib.bppp_rp_setup.argtypes = [vp, ip, ip, ip, C.c_char_p, ip, ip, sz, C.POINTER(RangeSpec), sz,
                                  C.POINTER(PublicSpec), C.POINTER(vp)]
    lib.bppp_rp_free.argtypes = [vp]
    lib.bppp_rp_free.restype = None
    lib.bppp_rp_last_error.argtypes = [vp]
    lib.bppp_rp_last_error.restype = C.c_char_p
    lib.bppp_rp_info.argtypes = [vp] + [C.POINTER(sz)] * 7
    lib.bppp_rp_points.argtypes = [vp, sz, u8p]
    lib.bppp_input_blind.argtypes = [C.c_char_p, C.c_uint64, u8p]
    lib.bppp_set_host_threads.argtypes = [ip]
    lib.bppp_set_host_threads.restype = None
    lib.bppp_rp_prove_batch.argtypes = [vp, sz, u8p, u8p, u8p, C.POINTER(C.c_char_p), u8p, u8p, u8p]
    lib.bppp_rp_verify_batch.argtypes = [vp, sz, sz, sz, sz, u8p, u8p, u8p, C.POINTER(ip)]
    lib.bppp_host_sha256.argtypes = [u8p, sz, u8p]
    lib.bppp_host_oracle.argtypes = [u8p, sz, ip, ip, u8p]
    lib.bppp_host_fr.argtypes = [ip, u8p, u8p, u8p]
    lib.bppp_host_get_points.argtypes = [C.c_char_p, sz, ip, u8p]
    lib.bppp_nl_prove.argtypes = [vp, sz, ip, sz, u8p, sz, u8p, u8p]
    lib.bppp_nl_challenges.argtypes = [sz, ip, sz, u8p, sz, u8p, u8p]
    lib.bppp_get_points.argtypes = [vp, C.c_char_p, sz, ip, u8p]
    lib.bppp_dtr_create.argtypes = [vp, sz, sz, ip, C.POINTER(vp)]
    lib.bppp_dtr_destroy.argtypes = [vp]
    lib.bppp_dtr_destroy.restype = None
    lib.bppp_dtr_reset.argtypes = [vp]
    lib.bppp_dtr_oracle.argtypes = [vp, u8p, sz, ip, u8p]
    lib.bppp_dev_random.argtypes = [vp, sz, C.POINTER(C.c_char_p), C.c_uint64, sz, u8p]
    lib.bppp_dtr_absorb.argtypes = [vp, u8p, sz, sz]
    lib.bppp_dtr_squeeze.argtypes = [vp, sz, u8p, u8p, u8p]
    lib.bppp_dtr_fits.argtypes = [vp, sz, sz, ip]
    lib.bppp_dtr_export.argtypes = [vp, sz, u8p, sz, C.POINTER(sz)]
    lib.bppp_rp_set_device_transcript.argtypes = [vp, ip]
    lib.bppp_nl_round_challenge.argtypes = [vp, u8p, u8p, u8p]
    lib.bppp_nl_attach_transcript.argtypes = [vp, vp]
    lib.bppp_rp_set_batch_verify.argtypes = [vp, ip]
    lib.bppp_gens_enable_lut.argtypes = [vp, C.c_double, C.POINTER(ip)]
    lib.bppp_rp_enable_lut.argtypes = [vp, C.c_double, C.POINTER(ip)]
    lib.bppp_comm_load.argtypes = [C.c_char_p]
    lib.bppp_comm_last_error.restype = C.c_char_p
    lib.bppp_comm_unique_id.argtypes = [u8p]
    lib.bppp_comm_create.argtypes = [vp, ip, ip, u8p, C.POINTER(vp)]
    lib.bppp_comm_destroy.argtypes = [vp]
    lib.bppp_comm_destroy.restype = None
    lib.bppp_nl_prove_sharded.argtypes = [vp, vp, sz, sz, sz, u8p, u8p, u8p, u8p, u8p]
    lib.bppp_nl_verify_gens_rlc.argtypes = [vp, ip, sz, sz, u8p, u8p, u8p, u8p, u8p, u8p, sz, sz, u8p, u8p, sz, u8p, u8p, u8p,
                                            C.POINTER(ip)]
    lib.bppp_nl_prove_device.argtypes = [vp, sz, u8p, u8p, u8p, u8p, u8p]
    lib.bppp_tune_process.argtypes = [ip]
    # this harness drives dedicated batch-proving processes (tests, bench.py): opt in to the process-wide
    # tuning (malloc arenas, blocking-sync device flags, pool pre-growth); BPPP_NO_TUNE=1 leaves the process alone
    if not os.environ.get("BPPP_NO_TUNE"):
        lib.bppp_tune_process(3)
    _LIB = lib
    return lib


class RangeSpec(C.Structure):
    _fields_ = [("min", C.c_uint8 * 16), ("max", C.c_uint8 * 16), ("base", C.c_uint32), ("is_shared", C.c_int32),
                ("is_output", C.c_int32), ("is_assumed", C.c_int32)]


class PublicSpec(C.Structure):
    _fields_ = [("amount", C.c_uint8 * 16), ("type", C.c_uint8 * 16), ("is_output", C.c_int32)]


def _i128(x):
    return (C.c_uint8 * 16)(*int(x).to_bytes(16, "little", signed=True))


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

    @classmethod
    def borrowed(cls, handle, device=0):
        """wrap a context owned by someone else (a lane of a RangeProofSetup)"""
        self = cls.__new__(cls)
        self.lib, self.h, self.device, self._borrowed = load_library(), handle, device, True
        return self

    def close(self):
        if getattr(self, "h", None) and not getattr(self, "_borrowed", False):
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

    # -- measurement support
    def profile_enable(self, on=True):
        self._ck(self.lib.bppp_profile_enable(self.h, 1 if on else 0), "bppp_profile_enable")

    def profile_reset(self):
        self._ck(self.lib.bppp_profile_reset(self.h), "bppp_profile_reset")

    def profile_report(self):
        import json
        buf = _buf(16384)
        self._ck(self.lib.bppp_profile_report(self.h, buf, 16384), "bppp_profile_report")
        return json.loads(buf.value.decode())

    def timer_start(self):
        self._ck(self.lib.bppp_timer_start(self.h), "bppp_timer_start")

    def timer_stop(self):
        ms = C.c_double()
        self._ck(self.lib.bppp_timer_stop(self.h, C.byref(ms)), "bppp_timer_stop")
        return ms.value

    def measure_imad_peak(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.lib.bppp_measure_imad_peak(self.h, C.byref(a), C.byref(b)), "bppp_measure_imad_peak")
        return a.value, b.value

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

    def gens_msm_batch(self, g, G, H, scalars):
        """MSMs over the resident generator list [g | G | H] (fixed-base tables): scalars [batch][n]"""
        gh = C.c_void_p()
        self._ck(self.lib.bppp_gens_create(self.h, len(G), len(H), point_to_bytes(g), points_to_bytes(G),
                                           points_to_bytes(H), C.byref(gh)), "bppp_gens_create")
        try:
            batch, n = len(scalars), len(scalars[0])
            out = _buf(64 * batch)
            self._ck(self.lib.bppp_gens_msm_batch(gh, batch, n, b"".join(ints_to_bytes(r) for r in scalars), out),
                     "bppp_gens_msm_batch")
            return bytes_to_points(out.raw[:64 * batch])
        finally:
            self.lib.bppp_gens_destroy(gh)

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

    def set_shard(self, first_element):
        self.ctx._ck(self.ctx.lib.bppp_nl_set_shard(self.h, first_element), "bppp_nl_set_shard")

    def export(self):
        """stored-form state: (nn[b], nl[b], points [b][n+m], c [b][m]) -- see bppp_nl_export"""
        n, m = self.lengths()
        nn, nl = _buf(32 * self.B), _buf(32 * self.B)
        pts, c = _buf(64 * self.B * (n + m)), _buf(32 * self.B * m)
        self.ctx._ck(self.ctx.lib.bppp_nl_export(self.h, nn, nl, pts, c), "bppp_nl_export")
        P = bytes_to_points(pts.raw[:64 * self.B * (n + m)])
        cs = bytes_to_ints(c.raw[:32 * self.B * m])
        return (bytes_to_ints(nn.raw[:32 * self.B]), bytes_to_ints(nl.raw[:32 * self.B]),
                [P[b * (n + m):(b + 1) * (n + m)] for b in range(self.B)], [cs[b * m:(b + 1) * m] for b in range(self.B)])

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


def _integer_log(b, n):
    r = 0
    while n >= b:
        n //= b
        r += 1
    return r


class RangeProofSetup:
    """`SetupRP` for TypedReciprocal / Binary range proofs (src/RangeProof.hs:25-55) built from a
    schema like the reference's CLI does (app/Parse.hs:97-172, app/Main.hs:255-335); proving and
    verifying run batched on the device behind `prove_batch` / `verify_batch` (= proveM / verifyM)."""

    def __init__(self, ctx, schema, show_format=0, root_policy=0):
        if not isinstance(schema, dict):
            import json
            with open(schema) as f:
                schema = json.load(f)
        self.ctx, self.schema = ctx, schema
        arg = {"ip": ARG_IP, "innerproduct": ARG_IP, "nl": ARG_NL, "normlinear": ARG_NL}[schema.get("argument", "IP").lower()]
        self.binary = bool(schema.get("binary", False))
        typed, con = schema.get("typed", False), schema.get("conserved", False)
        if typed and self.binary:
            raise BpppError("Can't make typed binary proof")
        self.arg = arg
        self.random_seed = schema.get("randomSeed", "default random seed")
        rs = []
        for r in schema["ranges"]:
            mn, mx = r.get("min", 0), r.get("max", 2 ** 64)
            if self.binary:
                base = 2
            else:
                l = _integer_log(2, mx - mn)
                base = r.get("base", l // max(_integer_log(2, l), 1))              # approxLogW (app/Parse.hs:193-197)
            spec = RangeSpec(_i128(mn), _i128(mx), base, int(r.get("isShared", False)), int(r.get("isOutput", False)),
                             int(r.get("isAssumed", False)))
            rs += [spec] * r.get("count", 1)
        ps = [PublicSpec(_i128(p["amount"]), _i128(p.get("type", 0)), int(p.get("isOutput", False)))
              for p in schema.get("public", [])]
        ra = (RangeSpec * max(len(rs), 1))(*rs)
        pa = (PublicSpec * max(len(ps), 1))(*ps)
        h = C.c_void_p()
        flag = con if self.binary else (typed or con)
        rc = ctx.lib.bppp_rp_setup(ctx.h, int(self.binary), arg, int(flag), schema.get("basisSeed", "test points").encode(),
                                   show_format, root_policy, len(rs), ra, len(ps), pa, C.byref(h))
        if rc:
            raise BpppError("bppp_rp_setup failed (%d): %s" % (rc, ctx.lib.bppp_last_error(ctx.h).decode()))
        self.h = h
        v = [C.c_size_t() for _ in range(7)]
        ctx.lib.bppp_rp_info(h, *[C.byref(x) for x in v])
        (self.n_inputs, self.num_rp_coms, self.nrm_len, self.lin_len, self.rounds, self.fin_norm, self.fin_lin) = [x.value for x in v]

    def _ck(self, rc, what):
        if rc:
            raise BpppError("%s failed (%d): %s" % (what, rc, self.ctx.lib.bppp_rp_last_error(self.h).decode()))

    def set_device_transcript(self, on=True):
        """run the Fiat-Shamir transcript of prove_batch / verify_batch on the device; bit-identical"""
        self._ck(self.ctx.lib.bppp_rp_set_device_transcript(self.h, int(bool(on))), "bppp_rp_set_device_transcript")

    def enable_lut(self, budget_gb):
        """full-multiples table for the setup's generators (csrc/lut.cuh); returns the window width (0 = none)"""
        c = C.c_int()
        self._ck(self.ctx.lib.bppp_rp_enable_lut(self.h, float(budget_gb), C.byref(c)), "bppp_rp_enable_lut")
        return c.value

    def set_batch_verify(self, on=True):
        """verify each lane's sub-batch by one random linear combination; per-proof checks locate failures"""
        self._ck(self.ctx.lib.bppp_rp_set_batch_verify(self.h, int(bool(on))), "bppp_rp_set_batch_verify")

    def contexts(self):
        """the contexts of all concurrent lanes (lane 0 first)"""
        n = C.c_size_t()
        self.ctx.lib.bppp_rp_contexts(self.h, None, 0, C.byref(n))
        arr = (C.c_void_p * n.value)()
        self._ck(self.ctx.lib.bppp_rp_contexts(self.h, arr, n.value, C.byref(n)), "bppp_rp_contexts")
        return [self.ctx] + [Context.borrowed(C.c_void_p(arr[i]), self.ctx.device) for i in range(1, n.value)]

    def points(self, count):
        out = _buf(64 * count)
        self._ck(self.ctx.lib.bppp_rp_points(self.h, count, out), "bppp_rp_points")
        return bytes_to_points(out.raw[:64 * count])

    def prove_batch_raw(self, batch, values, types, blinds, seeds):
        """bytes in / bytes out (see include/bppp_b200.h)."""
        nc = self.num_rp_coms + self.n_inputs
        coms, resp = _buf(64 * batch * nc), _buf(128 * batch * self.rounds)
        fin = _buf(32 * batch * (self.fin_norm + self.fin_lin))
        sa = (C.c_char_p * batch)(*[s.encode() if isinstance(s, str) else s for s in seeds])
        self._ck(self.ctx.lib.bppp_rp_prove_batch(self.h, batch, values, types, blinds, sa, coms, resp, fin), "bppp_rp_prove_batch")
        return coms.raw[:64 * batch * nc], resp.raw[:128 * batch * self.rounds], fin.raw[:32 * batch * (self.fin_norm + self.fin_lin)]

    def prove_batch(self, witnesses, seeds=None):
        """witnesses: [batch][n_inputs] dicts {"amount", "type"?, "blind"?} like witness.json.
        Returns [dict(coms, responses, finals)] in the oracle's proof shape."""
        B, n = len(witnesses), self.n_inputs
        seeds = seeds or [self.random_seed] * B
        R_ = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
        vals = b"".join(int_to_le(w["amount"] % R_) for ws in witnesses for w in ws)
        tys = b"".join(int_to_le(w.get("type", 0) % R_) for ws in witnesses for w in ws)
        blinds = None
        if any("blind" in w for ws in witnesses for w in ws):
            out = _buf(32)
            bl = []
            for ws, sd in zip(witnesses, seeds):
                for j, w in enumerate(ws):
                    if "blind" in w:
                        bl.append(int_to_le(w["blind"] % R_))
                    else:
                        self.ctx.lib.bppp_input_blind(sd.encode(), j + 1, out)
                        bl.append(out.raw[:32])
            blinds = b"".join(bl)
        coms, resp, fin = self.prove_batch_raw(B, vals, tys, blinds, seeds)
        nc, k, nf = self.num_rp_coms + n, self.rounds, self.fin_norm + self.fin_lin
        out = []
        for b in range(B):
            cs = bytes_to_points(coms[64 * b * nc:64 * (b + 1) * nc])
            rp = bytes_to_points(resp[128 * b * k:128 * (b + 1) * k])
            out.append(dict(coms=cs, responses=[(rp[2 * i], rp[2 * i + 1]) for i in range(k)],
                            finals=bytes_to_ints(fin[32 * b * nf:32 * (b + 1) * nf])))
        return out

    def encode_batch_raw(self, batch, coms, resp, fin):
        """-> (proof.bin images, commits.bin images), one fixed-size record per proof (encodeProof')"""
        pb, cb = C.c_size_t(), C.c_size_t()
        self.ctx.lib.bppp_rp_encoded_sizes(self.h, C.byref(pb), C.byref(cb))
        po, co = _buf(pb.value * batch), _buf(cb.value * batch)
        self._ck(self.ctx.lib.bppp_rp_encode_batch(self.h, batch, coms, resp, fin, po, co), "bppp_rp_encode_batch")
        return po.raw[:pb.value * batch], co.raw[:cb.value * batch], pb.value, cb.value

    def decode_batch_raw(self, batch, proof_bin, commits_bin):
        """decodeProof': -> (coms, responses, finals, ok[])"""
        nc, k, nf = self.num_rp_coms + self.n_inputs, self.rounds, self.fin_norm + self.fin_lin
        coms, resp, fin = _buf(64 * batch * nc), _buf(128 * batch * k), _buf(32 * batch * nf)
        ok = (C.c_int * batch)()
        self._ck(self.ctx.lib.bppp_rp_decode_batch(self.h, batch, proof_bin, commits_bin, coms, resp, fin, ok), "bppp_rp_decode_batch")
        return coms.raw[:64 * batch * nc], resp.raw[:128 * batch * k], fin.raw[:32 * batch * nf], [bool(v) for v in ok]

    def verify_batch_raw(self, batch, coms, resp, fin, rounds=None, n_norm=None, n_lin=None):
        ok = (C.c_int * batch)()
        self._ck(self.ctx.lib.bppp_rp_verify_batch(self.h, batch, self.rounds if rounds is None else rounds,
                                                   self.fin_norm if n_norm is None else n_norm,
                                                   self.fin_lin if n_lin is None else n_lin, coms, resp, fin, ok),
                 "bppp_rp_verify_batch")
        return [bool(v) for v in ok]

    def verify_batch(self, proofs):
        # the proof's shape comes from the setup, never from the proof (decodeProof', src/RangeProof.hs:70-71):
        # a proof with another number of rounds / final scalars / commitments is rejected without a device call
        nc, nf = self.num_rp_coms + self.n_inputs, self.fin_norm + self.fin_lin
        shaped = [len(p["responses"]) == self.rounds and len(p["finals"]) == nf and len(p["coms"]) == nc for p in proofs]
        good = [p for p, s in zip(proofs, shaped) if s]
        res = []
        if good:
            coms = b"".join(points_to_bytes(p["coms"]) for p in good)
            resp = b"".join(point_to_bytes(x) + point_to_bytes(r) for p in good for x, r in p["responses"])
            fin = b"".join(ints_to_bytes(p["finals"]) for p in good)
            res = self.verify_batch_raw(len(good), coms, resp, fin)
        it = iter(res)
        return [next(it) if s else False for s in shaped]

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.bppp_rp_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
