"""secp256k1 field constants and helpers (oracle; test infrastructure only).

Naming follows the reference (`elliptic-curve`): Fq = base field (coordinates),
Fr = scalar field (group order).  Constants:
src/Data/Curve/Weierstrass/FastSECP256K1.hs:34-56,117-127.
"""

Q = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F  # base field
R = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141  # group order
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
B_COEFF = 7

# CM constants (FastSECP256K1.hs:39,53): cube roots of unity in Fq / Fr
BETA = 0x7AE96A2B657C07106E64479EAC3434E99CF0497512F58995C1396C28719501EE
LAMBDA = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72


def inv(x, m=R):
    return pow(x % m, -1, m)


def batch_inverse(vs, m=R):
    """Montgomery batch inversion mapping 0 -> 0 (src/Data/Field/BatchInverse.hs:14-24)."""
    acc = 1
    pref = []
    for x in vs:
        pref.append(acc)
        if x % m:
            acc = acc * x % m
    y = inv(acc, m)
    out = [0] * len(vs)
    for i in range(len(vs) - 1, -1, -1):
        x = vs[i] % m
        if x:
            out[i] = y * pref[i] % m
            y = y * x % m
    return out


def powers(a, n, m=R):
    """`take n (powers a)` = 1, a, a^2 ... (src/Utils.hs:104-105)."""
    out, x = [], 1
    for _ in range(n):
        out.append(x)
        x = x * a % m
    return out


def powers1(a, n, m=R):
    """`take n (powers' a)` = a, a^2 ... (src/Utils.hs:107-108)."""
    out, x = [], a % m
    for _ in range(n):
        out.append(x)
        x = x * a % m
    return out


def powers2(b, a, n, m=R):
    """`take n (powers'' b a)` = b, b*a, b*a^2 ... (src/Utils.hs:110-111)."""
    out, x = [], b % m
    for _ in range(n):
        out.append(x)
        x = x * a % m
    return out


def reduce_scalar(x, m=R):
    """Centred lift (src/Commitment.hs:276-279)."""
    x %= m
    return -(m - x) if x > (m - x) else x


def rational_reduce_scalar(x, m=R):
    """First (a, b) of the truncated extended Euclid with a = b*x (mod m) and
    a^2 <= 2m (src/Commitment.hs:242-255, `quot` truncates toward zero)."""
    a0, a1 = (m, 0), (reduce_scalar(x, m), 1)
    while a1[0] * a1[0] > 2 * m:
        n, d = a0[0], a1[0]
        qt = abs(n) // abs(d)
        if (n < 0) != (d < 0):
            qt = -qt
        a0, a1 = a1, (a0[0] - qt * a1[0], a0[1] - qt * a1[1])
    return a1
