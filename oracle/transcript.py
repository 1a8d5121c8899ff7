"""Fiat-Shamir transcript, RNG and generator derivation (oracle; test infrastructure only).

Follows app/Main.hs:64-87 (`hash`, `getPoints`, `shaOracle`, `hashToScalar(s)`),
src/ZKP.hs:68-101 (`ZKPT`: transcript list + random counter) and
src/Encoding.hs:75-79 (digest -> field element).

Un-pinnable policies (see oracle/__init__.py, "PARITY UNPINNED"):
  * TranscriptFormat: how `show` renders a `Prime p` inside the hash pre-image --
    "P <dec>" (derived Show of galois-field-1.0.1's `newtype Prime = P Natural`,
    the default here) or bare "<dec>" (the in-repo FastPrime, FastPrime.hs:129-130).
  * RootPolicy of `pointX` (see curve.Secp256k1.lift_x).
"""
import hashlib

from .field import Q, R

PREFIXED_P = "PrefixedP"
BARE_DECIMAL = "BareDecimal"


def digest_to_int(d):
    """`Binary (Prime p)`.get: four big-endian Word64, first word least significant
    (src/Encoding.hs:75-79)."""
    w = [int.from_bytes(d[8 * i:8 * i + 8], "big") for i in range(4)]
    return w[0] + (w[1] << 64) + (w[2] << 128) + (w[3] << 192)


def hash_to(data, mod):
    """`hash = decode . fromStrict . SHA.hash` then `toP` (app/Main.hs:64-65)."""
    return digest_to_int(hashlib.sha256(data).digest()) % mod


def show_field(x, fmt=PREFIXED_P):
    return (b"P " if fmt == PREFIXED_P else b"") + str(x).encode()


def get_points(group, seed, count, root_policy="exp"):
    """First `count` elements of `getPoints seed` (app/Main.hs:68-72): x = hash(seed ++ show n)
    in Fq for n = 0,1,..., kept when pointX succeeds."""
    if isinstance(seed, str):
        seed = seed.encode()
    out, n = [], 0
    mod = Q if group.name == "secp256k1" else R
    while len(out) < count:
        p = group.lift_x(hash_to(seed + str(n).encode(), mod), root_policy)
        n += 1
        if p is not None:
            out.append(p)
    return out


class ZKPT:
    """The reference's transcript monad state (src/ZKP.hs:68-101): `cs` is the list of every
    commitment so far, NEWEST FIRST (`cs' = xs ++ cs`); `n` the `random` counter."""

    def __init__(self, group, random_seed=None, fmt=PREFIXED_P):
        self.group = group
        self.fmt = fmt
        self.seed = None if random_seed is None else (
            random_seed.encode() if isinstance(random_seed, str) else random_seed)
        self.cs = []
        self._enc = []          # cached "show x <> show y" per commitment
        self.n = 0
        self.hashed_bytes = 0

    def _coords(self, p):
        x, y = self.group.coords(p)
        return show_field(x, self.fmt) + show_field(y, self.fmt)

    def random(self):
        """`random` (src/ZKP.hs:90-93) with h = hashToScalar rn . show (app/Main.hs:177)."""
        if self.seed is None:
            raise RuntimeError("No Random in Verifier (app/Main.hs:193)")
        v = hash_to(self.seed + str(self.n).encode(), R)
        self.n += 1
        return v

    def oracle(self, xs, count=1):
        """`oracle xs` (src/ZKP.hs:96-101) -> first `count` scalars of `shaOracle cs'`
        (app/Main.hs:75-80): hash(show i ++ show (length cs') ++ concat coords), i = 1.."""
        self.cs = list(xs) + self.cs
        self._enc = [self._coords(p) for p in xs] + self._enc
        body = str(len(self.cs)).encode() + b"".join(self._enc)
        out = []
        for i in range(1, count + 1):
            data = str(i).encode() + body
            self.hashed_bytes += len(data)
            out.append(hash_to(data, R))
        return out


def input_blinds(random_seed, count):
    """`hashToScalars ("Blinding " <> rn)` (app/Main.hs:86-87,275-276), positions 1.."""
    if isinstance(random_seed, str):
        random_seed = random_seed.encode()
    return [hash_to(b"Blinding " + random_seed + str(i).encode(), R) for i in range(1, count + 1)]
