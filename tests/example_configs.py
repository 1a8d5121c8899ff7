"""The reference's shipped example configurations (examples/*/schema.json + witness.json), restated
as Python dicts so the tests do not need /root/reference at run time, plus the synthetic configs
of SURVEY.md section 8(d)."""

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
