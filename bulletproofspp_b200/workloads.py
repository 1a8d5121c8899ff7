"""Synthetic workloads of SURVEY.md 8(d): the reference's shipped example configurations
(examples/*/schema.json + witness.json restated as Python dicts, so nothing needs /root/reference at run
time), their batched variants, and the synthetic norm-argument sweep.  Input construction only -- used by
bench.py, tools/ and tests/; no arithmetic of the hot path lives here."""
import ctypes as _C


_U64 = 2 ** 64


def _rec(arg, ranges, **kw):
    d = {"basisSeed": "test points", "ranges": ranges}
    if arg:
        d["argument"] = arg
    d.update(kw)
    return d


def _nby64(count, base):
    return _rec("NL", [{"count": count, "base": base, "min": 0, "max": _U64, "isOutput": True, "isShared": True}])


EXAMPLES = {
    # examples/64bit: reciprocal base 16, inline digits, IP argument
    "64bit": (_rec("IP", [{"base": 16, "min": 0, "max": _U64, "isOutput": True}]), [{"amount": 1000000000}]),
    # examples/32bit
    "32bit": (_rec(None, [{"base": 9, "min": 0, "max": 2 ** 32, "isOutput": True}]), [{"amount": 10000}]),
    # examples/rec_test: typed, shared bases 3 and 16, one assumed range, IP by default
    "rec_test": (_rec(None, [
        {"base": 3, "min": 0, "max": _U64, "isShared": True, "isOutput": True},
        {"base": 16, "min": -20, "max": 73786976294838206463, "isShared": True, "isOutput": False},
        {"base": 5, "min": 1, "max": 625, "isShared": False, "isAssumed": True, "isOutput": False}],
        typed=True, public=[{"amount": 1, "type": 15, "isOutput": False}]),
        [{"amount": 124, "type": 15}, {"amount": 1, "type": 15}, {"amount": 122, "type": 15}]),
    # examples/bin_test: the only shipped binary example
    "bin_test": ({"binary": True, "conserved": True, "basisSeed": "test points", "argument": "NL",
                  "ranges": [{"min": 3, "max": _U64, "isOutput": True},
                             {"count": 2, "min": 2, "max": _U64, "isOutput": False, "isAssumed": True}],
                  "public": [{"amount": 2, "isOutput": False}]},
                 [{"amount": 124}, {"amount": 1}, {"amount": 121}]),
    "64by64": (_nby64(64, 256), [{"amount": 10000}] * 64),
    "96by64": (_nby64(96, 256), [{"amount": 10000}] * 96),
    "128by64": (_nby64(128, 256), [{"amount": 10000}] * 128),
    # synthetic: a single 64-bit BINARY norm-argument proof (what BASELINE.json configs[0] describes;
    # Binary.hs:165-167 needs conserved + a balancing public input)
    "bin64": ({"binary": True, "conserved": True, "argument": "NL", "basisSeed": "test points",
               "ranges": [{"max": _U64, "isOutput": True}], "public": [{"amount": 10 ** 9, "isOutput": False}]},
              [{"amount": 10 ** 9}]),
    # synthetic: typed NL reciprocal proof exercising types, inline digits, has-bit and assumed ranges
    "typed_nl": (_rec("NL", [
        {"base": 3, "min": 0, "max": _U64, "isShared": True, "isOutput": True},
        {"base": 16, "min": -20, "max": 73786976294838206463, "isShared": True, "isOutput": False},
        {"base": 9, "min": 0, "max": 2 ** 32, "isOutput": False},
        {"base": 5, "min": 1, "max": 625, "isShared": False, "isAssumed": True, "isOutput": False}],
        typed=True, public=[{"amount": 1, "type": 15, "isOutput": False}]),
        [{"amount": 124 + 1000, "type": 15}, {"amount": 1, "type": 15}, {"amount": 1000, "type": 15},
         {"amount": 122, "type": 15}]),
}
# 32by64: base 64 shared with a has-bit (examples/32by64)
EXAMPLES["32by64"] = (_nby64(32, 64), [{"amount": 10000}] * 32)


def batched(name, batch):
    """SURVEY 8(d) batched variant: proof b uses randomSeed "default random seed#b" and values + b."""
    schema, wit = EXAMPLES[name]
    seeds = ["default random seed#%d" % b for b in range(batch)]
    wits = [[dict(w, amount=w["amount"] + b) for w in wit] for b in range(batch)]
    return schema, wits, seeds


# ------------------------------------------------------------------------------------------------
# Synthetic norm-argument sweep (BASELINE.json config 5): one NormLinear argument of
# N = 2^e norm elements and M = 6 linear elements.  Generators: the first 1 + N + M points of
# getPoints "test points" (app/Main.hs:68-72; derived on the device, bppp_get_points); scalars
# w_i, l_j, c_j, q = SHA256("sweep" || e || tag || index) mod r (the `random` encoding, derived on the
# device, bppp_dev_random); s makes the relation s = |w|^2_q + <c, l> hold so that the verifier accepts.
# ------------------------------------------------------------------------------------------------
R_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141


def sweep_rounds(e):
    """NormArgument.hs:165-178 for N = 2^e >= 8, M = 6: rounds until the final witness is (4, 1)"""
    return e - 2


def sweep_generators(ctx, count, seed=b"test points"):
    out = _C.create_string_buffer(64 * count)
    ctx._ck(ctx.lib.bppp_get_points(ctx.h, seed, count, 0, out), "bppp_get_points")
    return out.raw[:64 * count]


def sweep_scalars(ctx, e, tag, n):
    """n canonical scalars SHA256("sweep" || e || tag || index) mod r as 32-byte little-endian strings"""
    out = _C.create_string_buffer(32 * n)
    seeds = (_C.c_char_p * 1)(("sweep%d%s" % (e, tag)).encode())
    ctx._ck(ctx.lib.bppp_dev_random(ctx.h, 1, seeds, 0, n, out), "bppp_dev_random")
    return out.raw[:32 * n]


def sweep_inputs(ctx, e, M=6):
    """-> dict(N, M, rounds, q, s, w, l, c) with byte strings; s from the relation (big-integer loop on the host)"""
    N = 1 << e
    q, w, l, c = (sweep_scalars(ctx, e, t, n) for t, n in (("q", 1), ("w", N), ("l", M), ("c", M)))
    qi = int.from_bytes(q, "little")
    q2 = qi * qi % R_ORDER
    acc, wt = 0, q2
    mv = memoryview(w)
    for i in range(N):                                    # |w|^2_q = sum (q^2)^(i+1) w_i^2
        x = int.from_bytes(mv[32 * i:32 * i + 32], "little")
        acc += wt * (x * x % R_ORDER)
        wt = wt * q2 % R_ORDER
    acc %= R_ORDER
    for j in range(M):
        acc += int.from_bytes(c[32 * j:32 * j + 32], "little") * int.from_bytes(l[32 * j:32 * j + 32], "little")
    return dict(N=N, M=M, rounds=sweep_rounds(e), q=q, s=(acc % R_ORDER).to_bytes(32, "little"), w=w, l=l, c=c)
