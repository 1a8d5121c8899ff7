// secp256k1 group law on top of fp.cuh (y^2 = x^3 + 7, a = 0).
//
// Replaces the point arithmetic the reference reaches through `nrmlAdd` / `dbl'`
// (src/Commitment.hs:94-169) and elliptic-curve-0.3.0.  Unlike the reference's mixed add
// (src/Commitment.hs:156-169, which returns z = 0 for P + P) every addition here is complete:
// the exceptional cases branch to a doubling / the identity.
//
// Affine points are (x, y) with the identity encoded as (0, 0) -- not on the curve -- which is
// also the C-ABI encoding (64 zero bytes).  Jacobian points are (X, Y, Z), identity Z = 0.
#pragma once
#include "fp.cuh"

namespace bppp {

struct Affine {
    u256 x, y;
};
struct Jac {
    u256 X, Y, Z;
};

BP_HD bool aff_is_inf(const Affine& p) { return u256_is_zero(p.x) && u256_is_zero(p.y); }
BP_HD Affine aff_inf() {
    Affine p;
    p.x = u256_zero();
    p.y = u256_zero();
    return p;
}
BP_HD Affine aff_neg(const Affine& p) {
    Affine r;
    r.x = p.x;
    r.y = fq::neg(p.y);
    return r;
}
BP_HD Affine aff_cneg(const Affine& p, bool neg) { return neg ? aff_neg(p) : p; }
BP_HD Jac jac_inf() {
    Jac r;
    r.X = u256_one();
    r.Y = u256_one();
    r.Z = u256_zero();
    return r;
}
BP_HD bool jac_is_inf(const Jac& p) { return u256_is_zero(p.Z); }
BP_HD Jac jac_from_aff(const Affine& p) {
    Jac r;
    if (aff_is_inf(p)) return jac_inf();
    r.X = p.x;
    r.Y = p.y;
    r.Z = u256_one();
    return r;
}
BP_HD Jac jac_neg(const Jac& p) {
    Jac r = p;
    r.Y = fq::neg(p.Y);
    return r;
}

// dbl-2009-l (a = 0): 2M + 5S.  F = FqCall (multiplications as calls: throughput kernels) or FqInl
// (inlined: latency-bound chains, see fp.cuh)
template <class F>
BP_HD Jac jac_dbl_t(const Jac& p) {
    if (jac_is_inf(p)) return p;
    u256 A = F::sqr(p.X);
    u256 B = F::sqr(p.Y);
    u256 C = F::sqr(B);
    u256 t = fq::add(p.X, B);
    u256 D = fq::dbl(fq::sub(fq::sub(F::sqr(t), A), C));
    u256 E = fq::add(fq::dbl(A), A);
    u256 Fs = F::sqr(E);
    Jac r;
    r.X = fq::sub(Fs, fq::dbl(D));
    u256 C8 = fq::dbl(fq::dbl(fq::dbl(C)));
    r.Y = fq::sub(F::mul(E, fq::sub(D, r.X)), C8);
    r.Z = fq::dbl(F::mul(p.Y, p.Z));
    return r;
}
BP_HD Jac jac_dbl(const Jac& p) { return jac_dbl_t<FqCall>(p); }

// Jacobian + affine, complete: 8M + 3S
BP_HD Jac jac_madd(const Jac& p, const Affine& q) {
    if (aff_is_inf(q)) return p;
    if (jac_is_inf(p)) return jac_from_aff(q);
    u256 Z1Z1 = fq::sqr(p.Z);
    u256 U2 = fq::mul(q.x, Z1Z1);
    u256 S2 = fq::mul(fq::mul(q.y, p.Z), Z1Z1);
    u256 H = fq::sub(U2, p.X);
    u256 r = fq::sub(S2, p.Y);
    if (u256_is_zero(H)) {
        if (u256_is_zero(r)) return jac_dbl(p);
        return jac_inf();
    }
    u256 HH = fq::sqr(H);
    u256 HHH = fq::mul(H, HH);
    u256 V = fq::mul(p.X, HH);
    Jac o;
    o.X = fq::sub(fq::sub(fq::sqr(r), HHH), fq::dbl(V));
    o.Y = fq::sub(fq::mul(r, fq::sub(V, o.X)), fq::mul(p.Y, HHH));
    o.Z = fq::mul(p.Z, H);
    return o;
}

// Jacobian + Jacobian, complete: 12M + 4S
template <class F>
BP_HD Jac jac_add_t(const Jac& p, const Jac& q) {
    if (jac_is_inf(p)) return q;
    if (jac_is_inf(q)) return p;
    u256 Z1Z1 = F::sqr(p.Z);
    u256 Z2Z2 = F::sqr(q.Z);
    u256 U1 = F::mul(p.X, Z2Z2);
    u256 U2 = F::mul(q.X, Z1Z1);
    u256 S1 = F::mul(F::mul(p.Y, q.Z), Z2Z2);
    u256 S2 = F::mul(F::mul(q.Y, p.Z), Z1Z1);
    u256 H = fq::sub(U2, U1);
    u256 r = fq::sub(S2, S1);
    if (u256_is_zero(H)) {
        if (u256_is_zero(r)) return jac_dbl_t<FqCall>(p);
        return jac_inf();
    }
    u256 HH = F::sqr(H);
    u256 HHH = F::mul(H, HH);
    u256 V = F::mul(U1, HH);
    Jac o;
    o.X = fq::sub(fq::sub(F::sqr(r), HHH), fq::dbl(V));
    o.Y = fq::sub(F::mul(r, fq::sub(V, o.X)), F::mul(S1, HHH));
    o.Z = F::mul(F::mul(p.Z, q.Z), H);
    return o;
}
BP_HD Jac jac_add(const Jac& p, const Jac& q) { return jac_add_t<FqCall>(p, q); }

// ---------------------------------------------------------------- XYZZ coordinates
// (X, Y, ZZ, ZZZ) with x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity: ZZ = 0.  The mixed addition is
// 8M + 2S (one squaring less than Jacobian + affine), which is what the bucket accumulation of the
// fixed-base MSM spends nearly all its time in.  (madd-2008-s, add-2008-s, dbl-2008-s-1, mdbl-2008-s-1)
struct Xyzz {
    u256 X, Y, ZZ, ZZZ;
};
BP_HD Xyzz xyzz_inf() {
    Xyzz r;
    r.X = u256_one(); r.Y = u256_one(); r.ZZ = u256_zero(); r.ZZZ = u256_zero();
    return r;
}
BP_HD bool xyzz_is_inf(const Xyzz& p) { return u256_is_zero(p.ZZ); }
BP_HD Xyzz xyzz_neg(const Xyzz& p) {
    Xyzz r = p;
    r.Y = fq::neg(p.Y);
    return r;
}
BP_HD Xyzz xyzz_dbl_aff(const Affine& q) {           // 2 * (affine point), never the identity on an odd-order curve
    u256 U = fq::dbl(q.y), V = fq::sqr(U), W = fq::mul(U, V), S = fq::mul(q.x, V);
    u256 x2 = fq::sqr(q.x), M = fq::add(fq::dbl(x2), x2);
    Xyzz r;
    r.X = fq::sub(fq::sqr(M), fq::dbl(S));
    r.Y = fq::sub(fq::mul(M, fq::sub(S, r.X)), fq::mul(W, q.y));
    r.ZZ = V;
    r.ZZZ = W;
    return r;
}
template <class F>
BP_HD Xyzz xyzz_dbl_t(const Xyzz& p) {
    if (xyzz_is_inf(p)) return p;
    u256 U = fq::dbl(p.Y), V = F::sqr(U), W = F::mul(U, V), S = F::mul(p.X, V);
    u256 x2 = F::sqr(p.X), M = fq::add(fq::dbl(x2), x2);
    Xyzz r;
    r.X = fq::sub(F::sqr(M), fq::dbl(S));
    r.Y = fq::sub(F::mul(M, fq::sub(S, r.X)), F::mul(W, p.Y));
    r.ZZ = F::mul(V, p.ZZ);
    r.ZZZ = F::mul(W, p.ZZZ);
    return r;
}
BP_HD Xyzz xyzz_dbl(const Xyzz& p) { return xyzz_dbl_t<FqCall>(p); }
// XYZZ + affine, complete: 8M + 2S
template <class F>
BP_HD Xyzz xyzz_madd_t(const Xyzz& p, const Affine& q) {
    if (aff_is_inf(q)) return p;
    if (xyzz_is_inf(p)) {
        Xyzz r;
        r.X = q.x; r.Y = q.y; r.ZZ = u256_one(); r.ZZZ = u256_one();
        return r;
    }
    u256 U2 = F::mul(q.x, p.ZZ);
    u256 S2 = F::mul(q.y, p.ZZZ);
    u256 P = fq::sub(U2, p.X);
    u256 R = fq::sub(S2, p.Y);
    if (u256_is_zero(P)) {
        if (u256_is_zero(R)) return xyzz_dbl_aff(q);
        return xyzz_inf();
    }
    u256 PP = F::sqr(P);
    u256 PPP = F::mul(P, PP);
    u256 Q = F::mul(p.X, PP);
    Xyzz o;
    o.X = fq::sub(fq::sub(F::sqr(R), PPP), fq::dbl(Q));
    o.Y = fq::sub(F::mul(R, fq::sub(Q, o.X)), F::mul(p.Y, PPP));
    o.ZZ = F::mul(p.ZZ, PP);
    o.ZZZ = F::mul(p.ZZZ, PPP);
    return o;
}
BP_HD Xyzz xyzz_madd(const Xyzz& p, const Affine& q) { return xyzz_madd_t<FqCall>(p, q); }
// XYZZ + XYZZ, complete: 12M + 2S
template <class F>
BP_HD Xyzz xyzz_add_t(const Xyzz& p, const Xyzz& q) {
    if (xyzz_is_inf(p)) return q;
    if (xyzz_is_inf(q)) return p;
    u256 U1 = F::mul(p.X, q.ZZ), U2 = F::mul(q.X, p.ZZ);
    u256 S1 = F::mul(p.Y, q.ZZZ), S2 = F::mul(q.Y, p.ZZZ);
    u256 P = fq::sub(U2, U1);
    u256 R = fq::sub(S2, S1);
    if (u256_is_zero(P)) {
        if (u256_is_zero(R)) return xyzz_dbl_t<FqCall>(p);
        return xyzz_inf();
    }
    u256 PP = F::sqr(P);
    u256 PPP = F::mul(P, PP);
    u256 Q = F::mul(U1, PP);
    Xyzz o;
    o.X = fq::sub(fq::sub(F::sqr(R), PPP), fq::dbl(Q));
    o.Y = fq::sub(F::mul(R, fq::sub(Q, o.X)), F::mul(S1, PPP));
    o.ZZ = F::mul(F::mul(p.ZZ, q.ZZ), PP);
    o.ZZZ = F::mul(F::mul(p.ZZZ, q.ZZZ), PPP);
    return o;
}
BP_HD Xyzz xyzz_add(const Xyzz& p, const Xyzz& q) { return xyzz_add_t<FqCall>(p, q); }
// the same point in Jacobian form with Z = ZZ*ZZZ  (X/ZZ = X ZZ ZZZ^2 / Z^2,  Y/ZZZ = Y ZZ^3 ZZZ^2 / Z^3): 5M + 2S
BP_HD Jac xyzz_to_jac(const Xyzz& p) {
    if (xyzz_is_inf(p)) return jac_inf();
    u256 w2 = fq::sqr(p.ZZZ), a = fq::mul(p.ZZ, w2);           // ZZ * ZZZ^2
    Jac r;
    r.X = fq::mul(p.X, a);
    r.Y = fq::mul(p.Y, fq::mul(fq::sqr(p.ZZ), a));             // Y * ZZ^3 * ZZZ^2
    r.Z = fq::mul(p.ZZ, p.ZZZ);
    return r;
}

// single-point conversion (one field inversion); batch conversion lives in kernels.cu
BP_HD Affine jac_to_aff(const Jac& p) {
    if (jac_is_inf(p)) return aff_inf();
    u256 zi = fq::inv(p.Z);
    u256 zi2 = fq::sqr(zi);
    Affine r;
    r.x = fq::mul(p.X, zi2);
    r.y = fq::mul(p.Y, fq::mul(zi2, zi));
    return r;
}

// ---------------------------------------------------------------- scalar recoding helpers
// Joint sparse form (Solinas) of two non-negative integers k0, k1 < 2^255.
// digits[j] = (u0 + 1) | ((u1 + 1) << 2), u in {-1,0,1}; returns the number of digits
// (<= bitlen + 1).  Joint Hamming density 1/2.
BP_HD int jsf_recode(uint8_t* digits, u256 k0, u256 k1, int max_digits) {
    int n = 0;
    uint32_t d0 = 0, d1 = 0;
    while (n < max_digits) {
        bool z0 = u256_is_zero(k0) && d0 == 0;
        bool z1 = u256_is_zero(k1) && d1 == 0;
        if (z0 && z1) break;
        uint32_t l0 = (k0.v[0] & 7u) + d0;   // low 3 bits of k0 + carry (mod 8 is what matters)
        uint32_t l1 = (k1.v[0] & 7u) + d1;
        int u0 = 0, u1 = 0;
        if (l0 & 1u) {
            u0 = 2 - (int)(l0 & 3u);                       // mods 4: 1 -> 1, 3 -> -1
            if (((l0 & 7u) == 3u || (l0 & 7u) == 5u) && (l1 & 3u) == 2u) u0 = -u0;
        }
        if (l1 & 1u) {
            u1 = 2 - (int)(l1 & 3u);
            if (((l1 & 7u) == 3u || (l1 & 7u) == 5u) && (l0 & 3u) == 2u) u1 = -u1;
        }
        if ((int)(2 * d0) == 1 + u0) d0 = 1 - d0;
        if ((int)(2 * d1) == 1 + u1) d1 = 1 - d1;
        // k >>= 1
#pragma unroll
        for (int i = 0; i < 7; i++) {
            k0.v[i] = (k0.v[i] >> 1) | (k0.v[i + 1] << 31);
            k1.v[i] = (k1.v[i] >> 1) | (k1.v[i + 1] << 31);
        }
        k0.v[7] >>= 1;
        k1.v[7] >>= 1;
        digits[n++] = (uint8_t)((u0 + 1) | ((u1 + 1) << 2));
    }
    return n;
}

// signed window digit j (width c <= 16) of a canonical scalar s < 2^256 given the carry from the
// digit below; returns digit in [-2^(c-1), 2^(c-1)] and updates carry.
BP_HD int signed_digit(const u256& s, int j, int c, int& carry) {
    int bit = j * c;
    uint32_t w = 0;
    if (bit < 256) {
        int limb = bit >> 5, sh = bit & 31;
        uint64_t two = s.v[limb];
        if (limb + 1 < 8) two |= (uint64_t)s.v[limb + 1] << 32;
        w = (uint32_t)(two >> sh) & ((1u << c) - 1u);
    }
    int d = (int)w + carry;
    if (d > (1 << (c - 1))) {
        d -= (1 << c);
        carry = 1;
    } else {
        carry = 0;
    }
    return d;
}

}  // namespace bppp
