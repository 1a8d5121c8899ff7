"""Proof / commitment wire format (oracle; test infrastructure only).

Restates src/Encoding.hs:75-134 and src/RangeProof.hs:60-85:
  * a scalar / coordinate = four big-endian Word64, LEAST significant word first (`Binary (Prime p)`);
  * commitments = packed sign bits (point i -> bit i mod 8, LSB first, of byte i div 8; set = the
    larger y, `getXAndSign`), then the x coordinates;
  * proof.bin = final witness scalars (`getWitness`), then the commitments rpComs ++ bpComs
    (bpComs = unPairs responses, NEWEST round first); commits.bin = the n input commitments.
"""
from .field import Q
from .curve import Secp256k1


def put_field(x):
    return b"".join(((x >> (64 * i)) & (2 ** 64 - 1)).to_bytes(8, "big") for i in range(4))


def get_field(b, mod):
    return sum(int.from_bytes(b[8 * i:8 * i + 8], "big") << (64 * i) for i in range(4)) % mod


def encode_commitments(pts):
    signs = bytearray((len(pts) + 7) // 8)
    for i, (x, y) in enumerate(pts):
        if y > (Q - y) % Q:
            signs[i // 8] |= 1 << (i % 8)
    return bytes(signs) + b"".join(put_field(x) for x, _ in pts)


def decode_commitments(b, n):
    ns = (n + 7) // 8
    signs, out = b[:ns], []
    for i in range(n):
        x = get_field(b[ns + 32 * i:ns + 32 * i + 32], Q)
        p = Secp256k1.lift_x(x)
        if p is None:
            return None
        y = p[1]
        big = (signs[i // 8] >> (i % 8)) & 1
        if (y > (Q - y) % Q) != bool(big):         # fromXWithSign (Encoding.hs:97-104)
            y = (Q - y) % Q
        out.append((x, y))
    return out


def encode_proof(setup, proof):
    """`encodeProof'` -> (commits.bin bytes, proof.bin bytes)."""
    k = setup.num_rp_coms
    rp_coms, n_coms = proof["coms"][:k], proof["coms"][k:]
    bp_coms = [p for xr in proof["responses"] for p in xr]
    scs = proof["opening"].vec.get_witness() if "opening" in proof else proof["finals"]
    return encode_commitments(n_coms), b"".join(put_field(s) for s in scs) + encode_commitments(rp_coms + bp_coms)
