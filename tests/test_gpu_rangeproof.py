"""GPU parity of the full range-proof provers / verifiers (host C++ phases + device group
operations, through the C ABI) against the golden vectors the oracle produced and, for the small
configs, against the oracle run live.  Bit-exact: commitments, responses and final scalars."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu

from example_configs import EXAMPLES, batched

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NL_CONFIGS = ["bin_test", "bin64", "typed_nl", "32by64", "64by64", "96by64", "128by64"]
IP_CONFIGS = ["64bit", "32bit", "rec_test"]          # reciprocal proofs over the IP argument (app/Parse.hs:100 default)


def load_golden(name):
    with open(os.path.join(GOLD, name.replace("#", "_b") + ".json")) as f:
        g = json.load(f)
    pt = lambda p: None if p is None else (int(p[0], 16), int(p[1], 16))
    g["coms"] = [pt(p) for p in g["coms"]]
    g["responses"] = [(pt(x), pt(r)) for x, r in g["responses"]]
    g["finals"] = [int(v, 16) for v in g["finals"]]
    return g


@pytest.mark.parametrize("name", NL_CONFIGS + IP_CONFIGS)
def test_prove_matches_golden_and_verifies(ctx, name):
    import bulletproofspp_b200 as bp
    schema, wit = EXAMPLES[name]
    g = load_golden(name)
    setup = bp.RangeProofSetup(ctx, schema)
    assert (setup.nrm_len, setup.lin_len, setup.rounds) == (g["nrm_len"], g["lin_len"], g["rounds"])
    proof = setup.prove_batch([wit])[0]
    assert proof["coms"] == g["coms"]
    assert proof["responses"] == g["responses"]
    assert proof["finals"] == g["finals"]
    assert setup.verify_batch([proof]) == [True]
    # tampering is rejected
    bad = dict(proof, finals=[(proof["finals"][0] + 1) % (2 ** 200)] + proof["finals"][1:])
    bad2 = dict(proof, coms=[proof["coms"][1], proof["coms"][0]] + proof["coms"][2:])
    assert setup.verify_batch([bad, proof, bad2]) == [False, True, False]
    setup.close()


def test_batched_proofs_have_independent_transcripts(ctx):
    import bulletproofspp_b200 as bp
    schema, wits, seeds = batched("128by64", 3)
    setup = bp.RangeProofSetup(ctx, schema)
    proofs = setup.prove_batch(wits, seeds)
    g1 = load_golden("128by64#1")
    assert proofs[1]["coms"] == g1["coms"] and proofs[1]["responses"] == g1["responses"] and proofs[1]["finals"] == g1["finals"]
    assert proofs[0]["responses"] != proofs[1]["responses"]
    assert setup.verify_batch(proofs) == [True, True, True]
    # a proof does not verify under another proof's commitments
    mixed = dict(proofs[0], coms=proofs[1]["coms"])
    assert setup.verify_batch([mixed]) == [False]
    setup.close()


def test_live_oracle_agreement_with_explicit_blinds(ctx):
    """small config, witness with caller-supplied blinds and a custom random seed, oracle run live"""
    import bulletproofspp_b200 as bp
    from oracle.curve import Secp256k1 as G
    from oracle.rangeproof import load_schema, load_witness, prove, verify
    from oracle.transcript import ZKPT
    schema = dict(EXAMPLES["bin64"][0], randomSeed="another seed")
    wit = [{"amount": 10 ** 9, "blind": 123456789}]
    so = load_schema(schema, G)
    po = prove(so, ZKPT(G, so.random_seed), load_witness(so, wit))
    setup = bp.RangeProofSetup(ctx, schema)
    p = setup.prove_batch([wit])[0]
    assert p["coms"] == po["coms"] and p["responses"] == po["responses"]
    assert p["finals"] == po["opening"].vec.get_witness()
    assert setup.verify_batch([p]) == [True]
    assert verify(so, ZKPT(G, None), dict(coms=p["coms"], responses=p["responses"], opening=po["opening"]))
    setup.close()


def test_invalid_witness_is_rejected(ctx):
    import bulletproofspp_b200 as bp
    schema, wit = EXAMPLES["bin_test"]
    setup = bp.RangeProofSetup(ctx, schema)
    with pytest.raises(bp.BpppError):
        setup.prove_batch([[{"amount": 125}, {"amount": 1}, {"amount": 121}]])      # unbalanced (Binary.hs:165-167)
    setup.close()
    schema, wit = EXAMPLES["32by64"]
    setup = bp.RangeProofSetup(ctx, schema)
    with pytest.raises(bp.BpppError):
        setup.prove_batch([[{"amount": 2 ** 64}] + wit[1:]])                          # out of range
    setup.close()


def test_ip_and_nl_give_different_proofs_for_the_same_statement(ctx):
    import bulletproofspp_b200 as bp
    schema, wit = EXAMPLES["64bit"]
    ip = bp.RangeProofSetup(ctx, schema)
    nl = bp.RangeProofSetup(ctx, dict(schema, argument="NL"))
    p_ip, p_nl = ip.prove_batch([wit])[0], nl.prove_batch([wit])[0]
    assert p_ip["coms"][2:] == p_nl["coms"][2:]            # digit / multiplicity / input commitments do not depend on q
    assert p_ip["responses"] != p_nl["responses"]
    assert ip.verify_batch([p_ip]) == [True] and nl.verify_batch([p_nl]) == [True]
    ip.close()
    nl.close()


@pytest.mark.parametrize("name", ["64bit", "bin_test", "128by64"])
def test_wire_format_roundtrip_and_oracle_bytes(ctx, name):
    """encodeProof' / decodeProof' (src/RangeProof.hs:60-85, src/Encoding.hs): the bytes equal the
    oracle's encoding of the golden proof; decode -> verify accepts; a flipped sign bit is rejected."""
    import bulletproofspp_b200 as bp
    from bulletproofspp_b200 import lib as L
    from oracle.encoding import encode_commitments, put_field
    schema, wit = EXAMPLES[name]
    g = load_golden(name)
    setup = bp.RangeProofSetup(ctx, schema)
    B = 2
    n = setup.n_inputs
    vals = b"".join(L.int_to_le(w["amount"]) for w in wit) * B
    tys = b"".join(L.int_to_le(w.get("type", 0)) for w in wit) * B
    coms, resp, fin = setup.prove_batch_raw(B, vals, tys, None, [setup.random_seed] * B)
    proof_bin, commits_bin, pb, cb = setup.encode_batch_raw(B, coms, resp, fin)
    k = setup.num_rp_coms
    exp_proof = b"".join(put_field(v) for v in g["finals"]) + encode_commitments(g["coms"][:k] + [p for xr in g["responses"] for p in xr])
    exp_commits = encode_commitments(g["coms"][k:])
    assert proof_bin[:pb] == exp_proof and proof_bin[pb:] == exp_proof
    assert commits_bin[:cb] == exp_commits
    if name == "64bit":
        assert pb == 418                                     # the paper's published size on secp256k1
    c2, r2, f2, ok = setup.decode_batch_raw(B, proof_bin, commits_bin)
    assert ok == [True, True] and (c2, r2, f2) == (coms, resp, fin)
    assert setup.verify_batch_raw(B, c2, r2, f2) == [True, True]
    nsc = setup.fin_norm + setup.fin_lin
    bad = bytearray(proof_bin)
    bad[32 * nsc] ^= 1                                       # flip the sign bit of the first commitment
    c3, r3, f3, ok3 = setup.decode_batch_raw(B, bytes(bad), commits_bin)
    assert ok3 == [True, True] and setup.verify_batch_raw(B, c3, r3, f3) == [False, True]
    setup.close()


@pytest.mark.parametrize("name", ["typed_nl", "32by64", "128by64"])
def test_device_scalar_phases_match_host_phases(ctx, name, monkeypatch):
    """The TypedReciprocal scalar phases run on the device by default (bppp_trrp_*); with
    BPPP_RP_HOST_PHASES=1 the same proofs come from the host implementation of those phases.
    Both must be bit-identical (and equal to the golden vectors, checked above).  A batch of 70
    proofs spreads over several lanes."""
    import bulletproofspp_b200 as bp
    schema, wits, seeds = batched(name, {"32by64": 70, "128by64": 3}.get(name, 1))   # typed proofs must stay balanced
    dev = bp.RangeProofSetup(ctx, schema)
    p_dev = dev.prove_batch(wits, seeds)
    monkeypatch.setenv("BPPP_RP_HOST_PHASES", "1")
    host = bp.RangeProofSetup(ctx, schema)
    monkeypatch.delenv("BPPP_RP_HOST_PHASES")
    p_host = host.prove_batch(wits, seeds)
    assert p_dev == p_host
    assert all(dev.verify_batch(p_host)) and all(host.verify_batch(p_dev))
    dev.close()
    host.close()


@pytest.mark.parametrize("rounds_on_device", ["1", "0"])
@pytest.mark.parametrize("name,count", [("typed_nl", 1), ("32by64", 70), ("128by64", 3)])
def test_device_transcript_is_bit_identical(ctx, name, count, rounds_on_device, monkeypatch):
    """with bppp_rp_set_device_transcript the commitments are rendered (`show`) and hashed
    (shaOracle) on the device and the norm blinders drawn there; the proofs must equal the host-transcript
    ones bit for bit (and the golden vectors), verification must accept them and reject tampering."""
    import bulletproofspp_b200 as bp
    # BPPP_DEVICE_ROUNDS=0: per-round challenges from the device, round constants on the host
    # (bppp_nl_round_challenge); default: the whole round loop on the device (bppp_nl_prove_device)
    monkeypatch.setenv("BPPP_DEVICE_ROUNDS", rounds_on_device)
    schema, wits, seeds = batched(name, count)
    host = bp.RangeProofSetup(ctx, schema)
    dev = bp.RangeProofSetup(ctx, schema)
    dev.set_device_transcript(True)
    p_host = host.prove_batch(wits, seeds)
    p_dev = dev.prove_batch(wits, seeds)
    assert p_dev == p_host
    if name == "128by64":
        g1 = load_golden("128by64#1")
        assert p_dev[1]["coms"] == g1["coms"] and p_dev[1]["responses"] == g1["responses"] and p_dev[1]["finals"] == g1["finals"]
    assert all(dev.verify_batch(p_host)) and all(host.verify_batch(p_dev))
    p = p_dev[0]
    bad = dict(p, finals=[(p["finals"][0] + 1) % (2 ** 200)] + p["finals"][1:])
    bad2 = dict(p, coms=[p["coms"][1], p["coms"][0]] + p["coms"][2:])
    bad3 = dict(p, responses=[p["responses"][1], p["responses"][0]] + p["responses"][2:])
    assert dev.verify_batch([bad, p, bad2, bad3]) == [False, True, False, False]
    # a second batch through the same setup (the transcripts are reset, the random counters restart)
    assert dev.prove_batch(wits, seeds) == p_host
    # back to the host transcript on the same setup
    dev.set_device_transcript(False)
    assert dev.prove_batch(wits[:1], seeds[:1]) == p_host[:1]
    host.close()
    dev.close()


def test_device_transcript_stages_in_one_squeeze(ctx, gens):
    """bppp_dtr_absorb + bppp_dtr_squeeze: the verifier's form -- all commitments absorbed first, then the
    challenges of every stage in one launch; each equals the oracle's transcript at that stage."""
    import ctypes as C
    from bulletproofspp_b200 import lib as L
    from oracle.curve import Secp256k1 as G
    from oracle.transcript import ZKPT
    B = 5
    base = gens(60)
    calls = [7, 1, 1, 2, 2, 2]
    rows = [[[base[(11 * b + 5 * ci + j) % 60] for j in range(n)] for b in range(B)] for ci, n in enumerate(calls)]
    t = C.c_void_p()
    ctx._ck(ctx.lib.bppp_dtr_create(ctx.h, B, sum(calls), 0, C.byref(t)), "bppp_dtr_create")
    for ci, n in enumerate(calls):
        # rows padded to a stride of n + 3 points
        raw = b"".join(b"".join(L.point_to_bytes(p) for p in rows[ci][b]) + bytes(64 * 3) for b in range(B))
        ctx._ck(ctx.lib.bppp_dtr_absorb(t, raw, n + 3, n), "bppp_dtr_absorb")
    plan = [(1, 1), (2, 1), (3, 1), (1, 2), (3, 2), (1, 3), (1, 4), (1, 5), (1, 6), (2, 0)]
    idx = bytes(i for i, _ in plan)
    st = bytes(s_ for _, s_ in plan)
    out = C.create_string_buffer(32 * B * len(plan))
    ctx._ck(ctx.lib.bppp_dtr_squeeze(t, len(plan), idx, st, out), "bppp_dtr_squeeze")
    got = L.bytes_to_ints(out.raw[:32 * B * len(plan)])
    for b in range(B):
        zk = ZKPT(G, None)
        want = {}
        for ci in range(len(calls)):
            ch = zk.oracle(rows[ci][b], 3)
            for i in (1, 2, 3):
                want[(i, ci + 1)] = ch[i - 1]
        for j, (i, s_) in enumerate(plan):
            assert got[b * len(plan) + j] == want[(i, s_ or len(calls))], (b, j)
    ctx.lib.bppp_dtr_destroy(t)


@pytest.mark.parametrize("name,count", [("32by64", 70), ("128by64", 5), ("bin_test", 4)])
def test_batch_verification_accepts_and_locates_bad_proofs(ctx, name, count):
    """one random linear combination per lane sub-batch (one fixed-base MSM + one Pippenger over all
    per-proof points).  All-good batches are accepted by the combination alone; with tampered proofs in the batch the
    combination fails and the per-proof checks return exactly the verdicts of plain verification."""
    import bulletproofspp_b200 as bp
    if name == "bin_test":
        schema, wit = EXAMPLES[name]
        wits, seeds = [wit] * count, ["seed %d" % i for i in range(count)]
    else:
        schema, wits, seeds = batched(name, count)
    setup = bp.RangeProofSetup(ctx, schema)
    proofs = setup.prove_batch(wits, seeds)
    plain = setup.verify_batch(proofs)
    assert all(plain)
    setup.set_batch_verify(True)
    l0 = ctx.launch_count()
    assert setup.verify_batch(proofs) == plain
    p = proofs[1]
    bad1 = dict(p, finals=[(p["finals"][0] + 1) % (2 ** 200)] + p["finals"][1:])
    bad2 = dict(proofs[2], coms=[proofs[2]["coms"][1], proofs[2]["coms"][0]] + proofs[2]["coms"][2:])
    mixed = [proofs[0], bad1, bad2] + proofs[3:]
    want = [True, False, False] + [True] * (count - 3)
    assert setup.verify_batch(mixed) == want
    # a proof under another proof's commitments, last in the batch
    swapped = proofs[:-1] + [dict(proofs[-1], coms=proofs[0]["coms"])]
    if name != "bin_test":
        assert setup.verify_batch(swapped) == [True] * (count - 1) + [False]
    setup.set_batch_verify(False)
    assert setup.verify_batch(mixed) == want
    setup.close()


@pytest.mark.parametrize("name,count,gb", [("128by64", 12, 2.5), ("32by64", 40, 1.4), ("typed_nl", 9, 6.0)])
def test_full_multiples_table_is_bit_identical(ctx, name, count, gb):
    """csrc/lut.cuh: with a full-multiples table for the generator list every fixed-base MSM of the batch prover and
    verifier is table lookups + mixed additions (k_msm_lut) instead of the nine-bit bucket kernel.  Same proof bits,
    same verdicts; the budgets here select window widths 11 and 12 (and the widest that fits for the small typed setup)."""
    import bulletproofspp_b200 as bp
    if name == "typed_nl":                                # typed proofs must stay balanced: same witness, different seeds
        schema, wit = EXAMPLES[name]
        wits, seeds = [wit] * count, ["typed seed %d" % i for i in range(count)]
    else:
        schema, wits, seeds = batched(name, count)
    plain = bp.RangeProofSetup(ctx, schema)
    lut = bp.RangeProofSetup(ctx, schema)
    c = lut.enable_lut(gb)
    p0 = 1 + lut.nrm_len + lut.lin_len
    want = max(cc for cc in range(10, 17) if p0 * ((256 + cc - 1) // cc) * 2 ** (cc - 1) * 64 <= gb * 1e9)
    assert c == want and (name, c) in (("128by64", 11), ("32by64", 12), ("typed_nl", c)), (c, want)
    for dev_tr in (False, True):
        plain.set_device_transcript(dev_tr)
        lut.set_device_transcript(dev_tr)
        p0 = plain.prove_batch(wits, seeds)
        p1 = lut.prove_batch(wits, seeds)
        assert p0 == p1
        assert all(lut.verify_batch(p0)) and all(plain.verify_batch(p1))
        # a lone proof takes the table too (32-term chunks, warp-level sum of the chunk partials)
        assert lut.prove_batch(wits[:1], seeds[:1]) == p0[:1] and lut.verify_batch(p0[:1]) == [True]
    bad = dict(p1[1], finals=[(p1[1]["finals"][0] + 1) % (2 ** 200)] + p1[1]["finals"][1:])
    assert lut.verify_batch([p1[0], bad] + p1[2:]) == [True, False] + [True] * (count - 2)
    lut.set_batch_verify(True)
    assert lut.verify_batch([p1[0], bad] + p1[2:]) == [True, False] + [True] * (count - 2)
    plain.close()
    lut.close()


def test_hybrid_round_mode_is_bit_identical(ctx, monkeypatch):
    """BPPP_HYBRID_MAX switches a tensor-mode argument to generator folding for its last rounds (the
    folded generators are materialised by one small fixed-base MSM per block).  Same proof bits."""
    import bulletproofspp_b200 as bp
    schema, wits, seeds = batched("128by64", 2)
    g1 = load_golden("128by64#1")
    setup = bp.RangeProofSetup(ctx, schema)
    plain = setup.prove_batch(wits, seeds)
    monkeypatch.setenv("BPPP_HYBRID_MAX", "96")
    proofs = setup.prove_batch(wits, seeds)
    monkeypatch.delenv("BPPP_HYBRID_MAX")
    assert proofs == plain
    p = proofs[1]
    assert p["coms"] == g1["coms"] and p["responses"] == g1["responses"] and p["finals"] == g1["finals"]
    assert setup.verify_batch(proofs) == [True, True]
    setup.close()


def test_wrong_shape_and_off_curve_points_are_rejected(ctx):
    """the proof's shape comes from the setup, and every supplied point must be on the curve.
    An extra round + an extra final scalar (a free, unbound term in the scalar check) must not verify."""
    import ctypes as C
    import bulletproofspp_b200 as bp
    from bulletproofspp_b200 import lib as L
    schema, wit = EXAMPLES["typed_nl"]
    setup = bp.RangeProofSetup(ctx, schema)
    proof = setup.prove_batch([wit])[0]
    assert setup.verify_batch([proof]) == [True]
    # one more round (a copy of the newest response) and one more final scalar
    longer = dict(proof, responses=[proof["responses"][0]] + proof["responses"], finals=proof["finals"] + [1])
    assert setup.verify_batch([longer, proof]) == [False, True]
    # the raw C entry point refuses shapes that are not the setup's
    coms = L.points_to_bytes(longer["coms"])
    resp = b"".join(L.point_to_bytes(x) + L.point_to_bytes(r) for x, r in longer["responses"])
    fin = L.ints_to_bytes(longer["finals"])
    ok = (C.c_int * 1)()
    rc = ctx.lib.bppp_rp_verify_batch(setup.h, 1, setup.rounds + 1, setup.fin_norm + 1, setup.fin_lin, coms, resp, fin, ok)
    assert rc != 0 and not ok[0]
    # an off-curve commitment / response point: same x, y + 1
    x, y = proof["coms"][0]
    off = dict(proof, coms=[(x, y + 1)] + proof["coms"][1:])
    xr, yr = proof["responses"][0][0]
    off2 = dict(proof, responses=[((xr, yr + 1), proof["responses"][0][1])] + proof["responses"][1:])
    assert setup.verify_batch([off, proof, off2]) == [False, True, False]
    setup.close()


def test_argument_seam_validates_shapes_and_points(ctx, gens):
    """bppp_nl_verify refuses final-witness lengths that k rounds cannot leave; bppp_msm / bppp_gens_create
    refuse bases that are not on the curve (the reference's pointX / fromA cannot produce them)."""
    import ctypes as C
    import bulletproofspp_b200 as bp
    from bulletproofspp_b200 import lib as L
    pts = gens(8)
    sc = L.ints_to_bytes([3, 5, 7])
    good = L.points_to_bytes(pts[:3])
    out = C.create_string_buffer(64)
    assert ctx.lib.bppp_msm(ctx.h, 3, sc, good, out) == 0
    bad = L.points_to_bytes([pts[0], (pts[1][0], pts[1][1] ^ 1), pts[2]])
    assert ctx.lib.bppp_msm(ctx.h, 3, sc, bad, out) == 3            # BPPP_ERR_RANGE
    h = C.c_void_p()
    assert ctx.lib.bppp_gens_create(ctx.h, 2, 0, L.point_to_bytes(pts[0]), bad[64:], None, C.byref(h)) == 3
    # N = 4, M = 0, k = 1 rounds leave 2 norm scalars: 3 is refused
    N, k = 4, 1
    z32 = bytes(32)
    okv = (C.c_int * 1)()
    args = lambda n_norm: (ctx.h, bp.ARG_NL, 1, N, 0, k, L.point_to_bytes(pts[0]), L.points_to_bytes(pts[1:1 + N]), None,
                           L.int_to_le(2), z32, z32 * N, None, L.int_to_le(9),
                           L.point_to_bytes(pts[5]) + L.point_to_bytes(pts[6]), n_norm, 0, z32 * n_norm, None, 0, None, None, okv)
    assert ctx.lib.bppp_nl_verify(*args(3)) == 1
    assert ctx.lib.bppp_nl_verify(*args(2)) == 0
