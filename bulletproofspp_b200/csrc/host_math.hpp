// Host-side scalar helpers for the round sequencing that stays on the CPU (the north star keeps
// the Fiat-Shamir transcript and the per-round constants on the host).
//
//   * Fr arithmetic: the portable path of fp.cuh (Montgomery, 8 x 32-bit limbs)
//   * batch inversion, 0 -> 0                       (src/Data/Field/BatchInverse.hs:14-24)
//   * rationalReduceScalar: truncated extended Euclid (src/Commitment.hs:242-255, 269-288)
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>
#include "fp.cuh"

// the big-integer helpers and rationalReduceScalar also run on the device (the device-resident round loop)
#ifdef __CUDACC__
#define BP_HDN __host__ __device__ inline
#else
#define BP_HDN inline
#endif

namespace bppp {
namespace host {

typedef unsigned __int128 u128;

inline u256 from_bytes(const uint8_t* b) {          // 32-byte little-endian
    u256 r;
    memcpy(r.v, b, 32);
    return r;
}
inline void to_bytes(uint8_t* b, const u256& a) { memcpy(b, a.v, 32); }
inline u256 from_u64(uint64_t x) {
    u256 r = u256_zero();
    r.v[0] = (uint32_t)x;
    r.v[1] = (uint32_t)(x >> 32);
    return r;
}
inline bool fr_is_canonical(const u256& a) { return !u256_geq(a, fr::modulus()); }
inline bool fq_is_canonical(const u256& a) { return !u256_geq(a, fq::modulus()); }

inline u256 fr_pow_u256(u256 base, const u256& e) {
    u256 acc = fr::one();
    for (int i = 255; i >= 0; i--) {
        acc = fr::sqr(acc);
        if (u256_bit(e, i)) acc = fr::mul(acc, base);
    }
    return acc;
}
// Montgomery-domain inverse (a^(r-2)); inv(0) = 0
inline u256 fr_inv(const u256& a) {
    u256 e = fr::modulus();
    e.v[0] -= 2;                                     // r - 2 (low limb 0xD0364141 - 2, no borrow)
    return fr_pow_u256(a, e);
}
// in-place batch inversion of Montgomery values, 0 -> 0
inline void fr_batch_inv(u256* v, size_t n) {
    std::vector<u256> pre(n);
    u256 acc = fr::one();
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        if (!u256_is_zero(v[i])) acc = fr::mul(acc, v[i]);
    }
    u256 y = fr_inv(acc);
    for (size_t i = n; i-- > 0;) {
        if (u256_is_zero(v[i])) continue;
        u256 x = v[i];
        v[i] = fr::mul(y, pre[i]);
        y = fr::mul(y, x);
    }
}

// ------------------------------------------------------------------ small signed big integers
struct SBig {               // magnitude in 5 x 64-bit limbs (< 2^320), sign
    uint64_t m[5];
    bool neg;
};
BP_HDN SBig sb_from_u256(const u256& a, bool neg) {
    SBig r;
    for (int i = 0; i < 4; i++) r.m[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
    r.m[4] = 0;
    r.neg = neg;
    return r;
}
BP_HDN u256 sb_to_u256(const SBig& a) {
    u256 r;
    for (int i = 0; i < 4; i++) { r.v[2 * i] = (uint32_t)a.m[i]; r.v[2 * i + 1] = (uint32_t)(a.m[i] >> 32); }
    return r;
}
BP_HDN bool sb_is_zero(const SBig& a) { return !(a.m[0] | a.m[1] | a.m[2] | a.m[3] | a.m[4]); }
BP_HDN int sb_bitlen(const SBig& a) {
    for (int i = 4; i >= 0; i--)
#ifdef __CUDA_ARCH__
        if (a.m[i]) return 64 * i + 64 - __clzll((long long)a.m[i]);
#else
        if (a.m[i]) return 64 * i + 64 - __builtin_clzll(a.m[i]);
#endif
    return 0;
}
BP_HDN int sb_cmp_mag(const SBig& a, const SBig& b) {
    for (int i = 4; i >= 0; i--)
        if (a.m[i] != b.m[i]) return a.m[i] < b.m[i] ? -1 : 1;
    return 0;
}
BP_HDN void sb_sub_mag(SBig& r, const SBig& a, const SBig& b) {   // |a| >= |b|
    u128 br = 0;
    for (int i = 0; i < 5; i++) {
        u128 d = (u128)a.m[i] - b.m[i] - br;
        r.m[i] = (uint64_t)d;
        br = (d >> 127) & 1;
    }
}
BP_HDN void sb_add_mag(SBig& r, const SBig& a, const SBig& b) {
    u128 c = 0;
    for (int i = 0; i < 5; i++) {
        c += (u128)a.m[i] + b.m[i];
        r.m[i] = (uint64_t)c;
        c >>= 64;
    }
}
BP_HDN SBig sb_shl(const SBig& a, int s) {
    SBig r;
    r.neg = a.neg;
    int w = s / 64, b = s % 64;
    for (int i = 4; i >= 0; i--) {
        uint64_t v = 0;
        if (i - w >= 0) {
            v = a.m[i - w] << b;
            if (b && i - w - 1 >= 0) v |= a.m[i - w - 1] >> (64 - b);
        }
        r.m[i] = v;
    }
    return r;
}
// a - b with signs
BP_HDN SBig sb_sub(const SBig& a, const SBig& b) {
    SBig r;
    if (a.neg != b.neg) {            // a - b = a + (-b): same sign as a, magnitudes add
        sb_add_mag(r, a, b);
        r.neg = a.neg;
    } else {
        int c = sb_cmp_mag(a, b);
        if (c >= 0) { sb_sub_mag(r, a, b); r.neg = a.neg; }
        else { sb_sub_mag(r, b, a); r.neg = !a.neg; }
    }
    if (sb_is_zero(r)) r.neg = false;
    return r;
}
// one Euclid step with Haskell `quot` semantics:  q = n quot d;  (n, ns) -= q * (d, ds)
// implemented as shift-subtract on magnitudes, applying the same steps to the cofactors.
BP_HDN void euclid_step(SBig& n, SBig& ns, const SBig& d, const SBig& ds) {
    // remainder keeps the sign of n; |n| -= k*2^s*|d| for the bits of |q|
    bool qneg = (n.neg != d.neg);
    int sh = sb_bitlen(n) - sb_bitlen(d);
    for (; sh >= 0; sh--) {
        SBig t = sb_shl(d, sh);
        if (sb_cmp_mag(n, t) >= 0) {
            bool nneg = n.neg;
            sb_sub_mag(n, n, t);
            n.neg = sb_is_zero(n) ? false : nneg;
            // ns -= (+-2^sh) * ds
            SBig u = sb_shl(ds, sh);
            u.neg = (ds.neg != qneg);
            ns = sb_sub(ns, u);
        }
    }
}
// The same step with the quotient taken in one go (Knuth D with a 64-bit divisor digit): the
// estimate from the leading 64 bits of d is at most 2 too large.  Falls back to the bitwise step
// when the quotient may not fit 63 bits (only for tiny inputs).
BP_HDN bool sb_mulsub_small(SBig& n, const SBig& d, uint64_t q) {      // |n| -= q*|d|; true if it went negative
    u128 carry = 0;
    uint64_t borrow = 0;
    for (int i = 0; i < 5; i++) {
        carry += (u128)d.m[i] * q;
        uint64_t sub = (uint64_t)carry;
        carry >>= 64;
        uint64_t t = n.m[i] - sub;
        uint64_t b1 = n.m[i] < sub;
        uint64_t t2 = t - borrow;
        uint64_t b2 = t < borrow;
        n.m[i] = t2;
        borrow = b1 | b2;
    }
    return borrow || carry;
}
BP_HDN void sb_addmul_small(SBig& r, const SBig& a, uint64_t q) {      // |r| += q*|a|
    u128 carry = 0;
    for (int i = 0; i < 5; i++) {
        carry += (u128)a.m[i] * q + r.m[i];
        r.m[i] = (uint64_t)carry;
        carry >>= 64;
    }
}
BP_HDN void euclid_step_fast(SBig& n, SBig& ns, const SBig& d, const SBig& ds) {
    const int bn = sb_bitlen(n), bd = sb_bitlen(d);
    if (bn < bd) return;                                   // quotient 0
    if (bn - bd > 61 || bd == 0) { euclid_step(n, ns, d, ds); return; }
    // leading 64 bits of d (normalised) and the matching window of n (at most 64 + 62 bits)
    const int sh = bd > 64 ? bd - 64 : 0;
    auto window = [&](const SBig& v) -> u128 {
        // (v >> sh) truncated to 128 bits
        const int w = sh / 64, b = sh % 64;
        uint64_t x0 = w < 5 ? v.m[w] : 0, x1 = w + 1 < 5 ? v.m[w + 1] : 0, x2 = w + 2 < 5 ? v.m[w + 2] : 0;
        uint64_t lo = b ? (x0 >> b) | (x1 << (64 - b)) : x0;
        uint64_t hi = b ? (x1 >> b) | (x2 << (64 - b)) : x1;
        return ((u128)hi << 64) | lo;
    };
    const uint64_t dt = (uint64_t)window(d);
    const u128 nt = window(n);
    // >= the true quotient, by at most 2 when dt is normalised; exact for bd <= 64 (sh = 0).
    // nt >> 64 < dt in both cases, so the 128/64 division cannot overflow.
    uint64_t q;
#if defined(__x86_64__) && !defined(__CUDACC__)
    {
        uint64_t rem_;
        asm("divq %4" : "=a"(q), "=d"(rem_) : "0"((uint64_t)nt), "1"((uint64_t)(nt >> 64)), "r"(dt) : "cc");
    }
#else
    q = (uint64_t)(nt / dt);
#endif
    if (q == 0) return;
    const bool nneg = n.neg, qneg = (n.neg != d.neg);
    SBig saved = n;
    while (sb_mulsub_small(n, d, q)) {                     // over-estimate: retry with q - 1
        n = saved;
        q--;
        if (q == 0) return;
    }
    n.neg = sb_is_zero(n) ? false : nneg;
    // ns -= (+-q) * ds
    const bool uneg = (ds.neg != qneg);                    // sign of the term u = +-q*ds that is subtracted
    if (ns.neg != uneg || sb_is_zero(ns)) {                // ns - u: signs differ -> magnitudes add, sign of ns (or of -u)
        const bool rneg = sb_is_zero(ns) ? !uneg : ns.neg;
        sb_addmul_small(ns, ds, q);
        ns.neg = sb_is_zero(ns) ? false : rneg;
    } else {                                               // same sign: |ns| - q|ds| may change sign
        SBig u;
        for (int i = 0; i < 5; i++) u.m[i] = 0;
        u.neg = uneg;
        sb_addmul_small(u, ds, q);
        ns = sb_sub(ns, u);
    }
}
struct Ratio {              // a = b * x (mod r), |a|,|b| about sqrt(r)
    u256 a, b;              // magnitudes
    bool a_neg, b_neg;
};
// x canonical (non-Montgomery) in [0, r)
BP_HDN Ratio rational_reduce(const u256& x) {
    const u256 r = fr::modulus();
    // centred lift (src/Commitment.hs:276-279): n > r - n  ->  -(r - n)
    u256 rm;
    u256_sub(rm, r, x);
    bool neg = !u256_geq(rm, x);                  // x > r - x
    SBig a0 = sb_from_u256(r, false), s0 = sb_from_u256(u256_zero(), false);
    SBig a1 = sb_from_u256(neg ? rm : x, neg), s1 = sb_from_u256(u256_one(), false);
    if (sb_is_zero(a1)) a1.neg = false;
    // stop when a1^2 <= 2r  <=>  bit test via comparison of squares (257-bit): use 128-bit halves
    auto too_big = [&](const SBig& v) {
        // v^2 > 2r ?   v < 2^257.  Quick outs on bit length, exact compare near the boundary.
        int bl = sb_bitlen(v);
        if (bl > 129) return true;
        if (bl < 129) return false;
        // bl == 129: v = 2^128 + lo, v^2 = 2^256 + 2^129*lo + lo^2; compare with 2r
        // compute v^2 exactly in 5 limbs (needs 258 bits)
        uint64_t w[3] = {v.m[0], v.m[1], v.m[2]};
        uint64_t sq[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 3; i++) {
            u128 c = 0;
            for (int j = 0; j < 3; j++) {
                c += (u128)w[i] * w[j] + sq[i + j];
                sq[i + j] = (uint64_t)c;
                c >>= 64;
            }
            sq[i + 3] += (uint64_t)c;
        }
        // 2r
        uint64_t rr[5];
        SBig R5 = sb_from_u256(r, false);
        for (int i = 4; i >= 0; i--) rr[i] = (R5.m[i] << 1) | (i ? R5.m[i - 1] >> 63 : 0);
        if (sq[5]) return true;
        for (int i = 4; i >= 0; i--)
            if (sq[i] != rr[i]) return sq[i] > rr[i];
        return false;
    };
    while (too_big(a1)) {
        euclid_step_fast(a0, s0, a1, s1);
        SBig t = a0; a0 = a1; a1 = t;
        t = s0; s0 = s1; s1 = t;
    }
    Ratio o;
    o.a = sb_to_u256(a1); o.a_neg = a1.neg;
    o.b = sb_to_u256(s1); o.b_neg = s1.neg;
    return o;
}
// signed small integer -> Fr (Montgomery)
BP_HDN u256 fr_from_signed(const u256& mag, bool neg) {
    u256 m = fr::to_mont(mag);
    return neg ? fr::neg(m) : m;
}

}  // namespace host
}  // namespace bppp
