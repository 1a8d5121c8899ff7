// 256-bit field arithmetic for secp256k1 in 8 x 32-bit limbs (sm_100a).
//
//   Fq  base field   q = 2^256 - 2^32 - 977   plain residues, always canonical (< q);
//                    reduction folds the high half by (2^32 + 977)  -- the special-prime
//                    idea of the reference's FastPrime (src/Data/Field/Galois/FastPrime/Internal.hs:939-957)
//   Fr  scalar field r = group order          Montgomery residues (R = 2^256), canonical (< r)
//
// Replaces the reference's `Prime p` (galois-field, Natural + mod) used by every
// field operation under src/Commitment.hs and src/Bulletproof*.hs.
//
// On the device the 256x256 product is a carry-chained mad.lo.cc/madc.hi.cc schedule over two
// interleaved accumulator rows (ptxas fuses each lo/hi pair into one IMAD.WIDE.U32[.X]); the
// portable path (host unit tests, and the reference the device self-test compares against)
// uses 64-bit C arithmetic.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BP_HD __host__ __device__ __forceinline__
#define BP_D __device__ __forceinline__
#else
#define BP_HD inline
#define BP_D inline
#endif

namespace bppp {

struct u256 {
    uint32_t v[8];
};

BP_HD u256 u256_zero() {
    u256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
BP_HD u256 u256_one() {
    u256 r = u256_zero();
    r.v[0] = 1;
    return r;
}
BP_HD bool u256_is_zero(const u256& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i];
    return o == 0;
}
BP_HD bool u256_eq(const u256& a, const u256& b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
    return o == 0;
}
// a >= b
BP_HD bool u256_geq(const u256& a, const u256& b) {
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.v[i] - b.v[i] - br;
        br = (d >> 63) & 1;
    }
    return br == 0;
}
BP_HD int u256_bit(const u256& a, int i) { return (a.v[i >> 5] >> (i & 31)) & 1; }

// r = a + b, returns carry
BP_HD uint32_t u256_add(u256& r, const u256& a, const u256& b) {
#if defined(__CUDA_ARCH__)
    uint32_t c;
    asm("add.cc.u32 %0,%9,%17; addc.cc.u32 %1,%10,%18; addc.cc.u32 %2,%11,%19; addc.cc.u32 %3,%12,%20;"
        "addc.cc.u32 %4,%13,%21; addc.cc.u32 %5,%14,%22; addc.cc.u32 %6,%15,%23; addc.cc.u32 %7,%16,%24;"
        "addc.u32 %8,0,0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    return c;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
#endif
}
// r = a - b, returns borrow (1 if a < b)
BP_HD uint32_t u256_sub(u256& r, const u256& a, const u256& b) {
#if defined(__CUDA_ARCH__)
    uint32_t c;
    asm("sub.cc.u32 %0,%9,%17; subc.cc.u32 %1,%10,%18; subc.cc.u32 %2,%11,%19; subc.cc.u32 %3,%12,%20;"
        "subc.cc.u32 %4,%13,%21; subc.cc.u32 %5,%14,%22; subc.cc.u32 %6,%15,%23; subc.cc.u32 %7,%16,%24;"
        "subc.u32 %8,0,0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    return c & 1;
#else
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.v[i] - b.v[i] - br;
        r.v[i] = (uint32_t)d;
        br = (d >> 63) & 1;
    }
    return (uint32_t)br;
#endif
}
// r = mask ? a : b   (mask all-ones or zero)
BP_HD u256 u256_sel(uint32_t mask, const u256& a, const u256& b) {
    u256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (a.v[i] & mask) | (b.v[i] & ~mask);
    return r;
}

// ------------------------------------------------------------------ 256 x 256 -> 512
BP_HD void mul_wide_portable(uint32_t t[16], const u256& a, const u256& b) {
    for (int i = 0; i < 16; i++) t[i] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a.v[j] * b.v[i] + t[i + j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        t[i + 8] = (uint32_t)c;
    }
}

#if defined(__CUDA_ARCH__)
// x[0..7] += {a0,a1,a2,a3} * b as four (lo,hi) pairs with one carry chain; the carry out of
// x[7] is added into x8.
BP_D void madc_row(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3, uint32_t& x4, uint32_t& x5,
                   uint32_t& x6, uint32_t& x7, uint32_t& x8, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                   uint32_t b) {
    asm("mad.lo.cc.u32 %0,%9,%13,%0; madc.hi.cc.u32 %1,%9,%13,%1;"
        "madc.lo.cc.u32 %2,%10,%13,%2; madc.hi.cc.u32 %3,%10,%13,%3;"
        "madc.lo.cc.u32 %4,%11,%13,%4; madc.hi.cc.u32 %5,%11,%13,%5;"
        "madc.lo.cc.u32 %6,%12,%13,%6; madc.hi.cc.u32 %7,%12,%13,%7;"
        "addc.u32 %8,%8,0;"
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7), "+r"(x8)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// first row: x[0..7] = {a0..a3} * b
BP_D void mul_row(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3, uint32_t& x4, uint32_t& x5, uint32_t& x6,
                  uint32_t& x7, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    asm("mul.lo.u32 %0,%8,%12; mul.hi.u32 %1,%8,%12; mul.lo.u32 %2,%9,%12; mul.hi.u32 %3,%9,%12;"
        "mul.lo.u32 %4,%10,%12; mul.hi.u32 %5,%10,%12; mul.lo.u32 %6,%11,%12; mul.hi.u32 %7,%11,%12;"
        : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3), "=r"(x4), "=r"(x5), "=r"(x6), "=r"(x7)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

// Row-wise schoolbook product.  E holds the (lo,hi) pairs that start on an even limb position,
// O[m] holds limb position m+1 (pairs that start on an odd position).  For row i the products
// a_j*b_i with i+j even go to E, the others to O; the chain whose top pair already exists
// spills its carry into the next (fresh) limb.
BP_D void mul_wide_dev(uint32_t t[16], const u256& a, const u256& b) {
    uint32_t E[18], O[18];
#pragma unroll
    for (int i = 8; i < 18; i++) { E[i] = 0; O[i] = 0; }
    mul_row(E[0], E[1], E[2], E[3], E[4], E[5], E[6], E[7], a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
    mul_row(O[0], O[1], O[2], O[3], O[4], O[5], O[6], O[7], a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
        if (i & 1) {
            // even j -> odd positions i+j -> O[i+j-1]; odd j -> even positions -> E[i+j]
            madc_row(O[i - 1], O[i], O[i + 1], O[i + 2], O[i + 3], O[i + 4], O[i + 5], O[i + 6], O[i + 7], a.v[0],
                     a.v[2], a.v[4], a.v[6], b.v[i]);
            madc_row(E[i + 1], E[i + 2], E[i + 3], E[i + 4], E[i + 5], E[i + 6], E[i + 7], E[i + 8], E[i + 9],
                     a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
        } else {
            madc_row(E[i], E[i + 1], E[i + 2], E[i + 3], E[i + 4], E[i + 5], E[i + 6], E[i + 7], E[i + 8], a.v[0],
                     a.v[2], a.v[4], a.v[6], b.v[i]);
            madc_row(O[i], O[i + 1], O[i + 2], O[i + 3], O[i + 4], O[i + 5], O[i + 6], O[i + 7], O[i + 8], a.v[1],
                     a.v[3], a.v[5], a.v[7], b.v[i]);
        }
    }
    // t = E + (O << 32)
    t[0] = E[0];
    uint32_t c1, dummy;
    asm("add.cc.u32 %0,%9,%17; addc.cc.u32 %1,%10,%18; addc.cc.u32 %2,%11,%19; addc.cc.u32 %3,%12,%20;"
        "addc.cc.u32 %4,%13,%21; addc.cc.u32 %5,%14,%22; addc.cc.u32 %6,%15,%23; addc.cc.u32 %7,%16,%24;"
        "addc.u32 %8,0,0;"
        : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(c1)
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(O[0]),
          "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
    asm("add.cc.u32 %7,%22,0xffffffff; addc.cc.u32 %0,%8,%15; addc.cc.u32 %1,%9,%16; addc.cc.u32 %2,%10,%17;"
        "addc.cc.u32 %3,%11,%18; addc.cc.u32 %4,%12,%19; addc.cc.u32 %5,%13,%20; addc.u32 %6,%14,%21;"
        : "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15]), "=r"(dummy)
        : "r"(E[9]), "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]), "r"(O[8]), "r"(O[9]),
          "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]), "r"(c1));
}

// shorter carry chains for the squaring schedule: x[0..2n-1] += {a...} * b, carry into x[2n]
BP_D void madc_row3(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3, uint32_t& x4, uint32_t& x5,
                    uint32_t& x6, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
    asm("mad.lo.cc.u32 %0,%7,%10,%0; madc.hi.cc.u32 %1,%7,%10,%1;"
        "madc.lo.cc.u32 %2,%8,%10,%2; madc.hi.cc.u32 %3,%8,%10,%3;"
        "madc.lo.cc.u32 %4,%9,%10,%4; madc.hi.cc.u32 %5,%9,%10,%5;"
        "addc.u32 %6,%6,0;"
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6)
        : "r"(a0), "r"(a1), "r"(a2), "r"(b));
}
BP_D void madc_row2(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3, uint32_t& x4, uint32_t a0,
                    uint32_t a1, uint32_t b) {
    asm("mad.lo.cc.u32 %0,%5,%7,%0; madc.hi.cc.u32 %1,%5,%7,%1;"
        "madc.lo.cc.u32 %2,%6,%7,%2; madc.hi.cc.u32 %3,%6,%7,%3;"
        "addc.u32 %4,%4,0;"
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4)
        : "r"(a0), "r"(a1), "r"(b));
}
BP_D void madc_row1(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t a0, uint32_t b) {
    asm("mad.lo.cc.u32 %0,%3,%4,%0; madc.hi.cc.u32 %1,%3,%4,%1; addc.u32 %2,%2,0;"
        : "+r"(x0), "+r"(x1), "+r"(x2)
        : "r"(a0), "r"(b));
}

// 256-bit square: the 28 cross products a_i*a_j (i<j) once, doubled by a one-bit shift, plus the
// 8 diagonal squares.  Row i sends the products with odd i+j to O (index 2i) and the ones with
// even i+j to E (index 2i+2); every chain's carry lands in a limb that is still 0 or 1.
BP_D void sqr_wide_dev(uint32_t t[16], const u256& a) {
    uint32_t E[16], O[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { E[i] = 0; O[i] = 0; }
    const uint32_t* v = a.v;
    mul_row(O[0], O[1], O[2], O[3], O[4], O[5], O[6], O[7], v[1], v[3], v[5], v[7], v[0]);
    madc_row3(E[2], E[3], E[4], E[5], E[6], E[7], E[8], v[2], v[4], v[6], v[0]);
    madc_row3(O[2], O[3], O[4], O[5], O[6], O[7], O[8], v[2], v[4], v[6], v[1]);
    madc_row3(E[4], E[5], E[6], E[7], E[8], E[9], E[10], v[3], v[5], v[7], v[1]);
    madc_row3(O[4], O[5], O[6], O[7], O[8], O[9], O[10], v[3], v[5], v[7], v[2]);
    madc_row2(E[6], E[7], E[8], E[9], E[10], v[4], v[6], v[2]);
    madc_row2(O[6], O[7], O[8], O[9], O[10], v[4], v[6], v[3]);
    madc_row2(E[8], E[9], E[10], E[11], E[12], v[5], v[7], v[3]);
    madc_row2(O[8], O[9], O[10], O[11], O[12], v[5], v[7], v[4]);
    madc_row1(E[10], E[11], E[12], v[6], v[4]);
    madc_row1(O[10], O[11], O[12], v[6], v[5]);
    madc_row1(E[12], E[13], E[14], v[7], v[5]);
    madc_row1(O[12], O[13], O[14], v[7], v[6]);
    // s = E + (O << 32)   (fits 511 bits: it is half of a^2 - diagonal)
    uint32_t s[16];
    s[0] = 0;                                   // E[0] is never written
    uint32_t c1;
    asm("add.cc.u32 %0,%9,%17; addc.cc.u32 %1,%10,%18; addc.cc.u32 %2,%11,%19; addc.cc.u32 %3,%12,%20;"
        "addc.cc.u32 %4,%13,%21; addc.cc.u32 %5,%14,%22; addc.cc.u32 %6,%15,%23; addc.cc.u32 %7,%16,%24;"
        "addc.u32 %8,0,0;"
        : "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(s[8]), "=r"(c1)
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(O[0]),
          "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
    uint32_t dummy;
    asm("add.cc.u32 %7,%22,0xffffffff; addc.cc.u32 %0,%8,%15; addc.cc.u32 %1,%9,%16; addc.cc.u32 %2,%10,%17;"
        "addc.cc.u32 %3,%11,%18; addc.cc.u32 %4,%12,%19; addc.cc.u32 %5,%13,%20; addc.u32 %6,%14,%21;"
        : "=r"(s[9]), "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15]), "=r"(dummy)
        : "r"(E[9]), "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]), "r"(O[8]), "r"(O[9]),
          "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]), "r"(c1));
    // t = 2*s + diag
    uint32_t d[16];
    d[0] = 0;
#pragma unroll
    for (int k = 1; k < 16; k++) d[k] = __funnelshift_l(s[k - 1], s[k], 1);
    uint32_t D[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { D[2 * i] = v[i] * v[i]; D[2 * i + 1] = __umulhi(v[i], v[i]); }
    asm("add.cc.u32 %0,%8,%16; addc.cc.u32 %1,%9,%17; addc.cc.u32 %2,%10,%18; addc.cc.u32 %3,%11,%19;"
        "addc.cc.u32 %4,%12,%20; addc.cc.u32 %5,%13,%21; addc.cc.u32 %6,%14,%22; addc.cc.u32 %7,%15,%23;"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
        : "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]), "r"(D[0]),
          "r"(D[1]), "r"(D[2]), "r"(D[3]), "r"(D[4]), "r"(D[5]), "r"(D[6]), "r"(D[7]));
    asm("addc.cc.u32 %0,%8,%16; addc.cc.u32 %1,%9,%17; addc.cc.u32 %2,%10,%18; addc.cc.u32 %3,%11,%19;"
        "addc.cc.u32 %4,%12,%20; addc.cc.u32 %5,%13,%21; addc.cc.u32 %6,%14,%22; addc.u32 %7,%15,%23;"
        : "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
        : "r"(d[8]), "r"(d[9]), "r"(d[10]), "r"(d[11]), "r"(d[12]), "r"(d[13]), "r"(d[14]), "r"(d[15]), "r"(D[8]),
          "r"(D[9]), "r"(D[10]), "r"(D[11]), "r"(D[12]), "r"(D[13]), "r"(D[14]), "r"(D[15]));
}
#endif

BP_HD void mul_wide(uint32_t t[16], const u256& a, const u256& b) {
#if defined(__CUDA_ARCH__)
    mul_wide_dev(t, a, b);
#else
    mul_wide_portable(t, a, b);
#endif
}

// ============================================================================ Fq
namespace fq {
// q = 2^256 - C,  C = 2^32 + 977
#define BP_FQ_C0 977u
BP_HD u256 modulus() {
    u256 p;
    p.v[0] = 0xFFFFFC2Fu; p.v[1] = 0xFFFFFFFEu;
#pragma unroll
    for (int i = 2; i < 8; i++) p.v[i] = 0xFFFFFFFFu;
    return p;
}
// r + C with carry out
BP_HD uint32_t add_c(u256& r, const u256& a) {
    u256 c = u256_zero();
    c.v[0] = BP_FQ_C0;
    c.v[1] = 1;
    return u256_add(r, a, c);
}
// canonical value of (carry:a) known to be < 2q
BP_HD u256 cond_sub(const u256& a, uint32_t carry) {
    u256 t;
    uint32_t c2 = add_c(t, a);          // a >= q  <=>  a + C >= 2^256
    uint32_t m = 0u - ((carry | c2) & 1u);
    return u256_sel(m, t, a);
}
BP_HD u256 add(const u256& a, const u256& b) {
    u256 s;
    uint32_t c = u256_add(s, a, b);
    return cond_sub(s, c);
}
BP_HD u256 sub(const u256& a, const u256& b) {
    u256 d, t;
    uint32_t br = u256_sub(d, a, b);
    u256 p = modulus();
    u256_add(t, d, p);
    return u256_sel(0u - br, t, d);
}
BP_HD u256 neg(const u256& a) {
    u256 d;
    u256_sub(d, modulus(), a);
    return u256_sel(u256_is_zero(a) ? 0xFFFFFFFFu : 0u, a, d);
}
BP_HD u256 dbl(const u256& a) { return add(a, a); }

// reduce a 512-bit product
BP_HD u256 reduce512(const uint32_t t[16]) {
    // fold 1: lo + hi*977 + (hi << 32)  -> 9 limbs + small overflow
    uint32_t r[10];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)t[8 + i] * BP_FQ_C0 + t[i];
        if (i > 0) c += t[8 + i - 1];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    c += t[15];
    r[8] = (uint32_t)c;
    r[9] = (uint32_t)(c >> 32);
    // fold 2: (r[8], r[9]) * C   (r9:r8 < 2^34)
    uint64_t h = ((uint64_t)r[9] << 32) | r[8];
    uint64_t lo = h * BP_FQ_C0;               // < 2^44
    u256 out;
    c = (uint64_t)r[0] + (uint32_t)lo;
    out.v[0] = (uint32_t)c; c >>= 32;
    c += (uint64_t)r[1] + (uint32_t)(lo >> 32) + (uint32_t)h;
    out.v[1] = (uint32_t)c; c >>= 32;
    c += (uint64_t)r[2] + (uint32_t)(h >> 32);
    out.v[2] = (uint32_t)c; c >>= 32;
#pragma unroll
    for (int i = 3; i < 8; i++) {
        c += r[i];
        out.v[i] = (uint32_t)c;
        c >>= 32;
    }
    // value = out + c*2^256 with c in {0,1}; if c then out is tiny (< 2^45) so out + C < q
    return cond_sub(out, (uint32_t)c);
}
#if defined(__CUDA_ARCH__)
// The same reduction with explicit carry chains (device).  The eight products hi_i * 977 are independent 64-bit
// multiplications (no addend, no carry between them: written as a multiply-accumulate chain, ptxas needs a zeroed
// register pair per step -- 16 moves on the multiplier pipe, and the chain is serial).  The even ones laid end to end
// are one 256-bit number, the odd ones another at limb 1, so r = lo + EVEN + ((ODD + hi) << 32) is three add chains
// on the integer pipe.
BP_D u256 reduce512_dev(const uint32_t t[16]) {
    const uint32_t K = BP_FQ_C0;
    uint32_t pl[8], ph[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint64_t p = (uint64_t)t[8 + i] * K;
        pl[i] = (uint32_t)p;
        ph[i] = (uint32_t)(p >> 32);
    }
    // r = lo + EVEN   (limbs 0 .. 7, carry into limb 8)
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7, r8, r9;
    asm("add.cc.u32 %0,%9,%17; addc.cc.u32 %1,%10,%18; addc.cc.u32 %2,%11,%19; addc.cc.u32 %3,%12,%20;"
        "addc.cc.u32 %4,%13,%21; addc.cc.u32 %5,%14,%22; addc.cc.u32 %6,%15,%23; addc.cc.u32 %7,%16,%24;"
        "addc.u32 %8,0,0;"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3), "=&r"(r4), "=&r"(r5), "=&r"(r6), "=&r"(r7), "=&r"(r8)
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(pl[0]), "r"(ph[0]), "r"(pl[2]),
          "r"(ph[2]), "r"(pl[4]), "r"(ph[4]), "r"(pl[6]), "r"(ph[6]));
    // s = ODD + hi   (8 limbs and a carry; they sit at limbs 1 .. 9 of r)
    uint32_t s0, s1, s2, s3, s4, s5, s6, s7, s8;
    asm("add.cc.u32 %0,%9,%17; addc.cc.u32 %1,%10,%18; addc.cc.u32 %2,%11,%19; addc.cc.u32 %3,%12,%20;"
        "addc.cc.u32 %4,%13,%21; addc.cc.u32 %5,%14,%22; addc.cc.u32 %6,%15,%23; addc.cc.u32 %7,%16,%24;"
        "addc.u32 %8,0,0;"
        : "=&r"(s0), "=&r"(s1), "=&r"(s2), "=&r"(s3), "=&r"(s4), "=&r"(s5), "=&r"(s6), "=&r"(s7), "=&r"(s8)
        : "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]), "r"(pl[1]), "r"(ph[1]), "r"(pl[3]),
          "r"(ph[3]), "r"(pl[5]), "r"(ph[5]), "r"(pl[7]), "r"(ph[7]));
    // r += s << 32
    asm("add.cc.u32 %0,%0,%9; addc.cc.u32 %1,%1,%10; addc.cc.u32 %2,%2,%11; addc.cc.u32 %3,%3,%12;"
        "addc.cc.u32 %4,%4,%13; addc.cc.u32 %5,%5,%14; addc.cc.u32 %6,%6,%15; addc.cc.u32 %7,%7,%16;"
        "addc.u32 %8,%17,0;"
        : "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(r8), "=r"(r9)
        : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(s4), "r"(s5), "r"(s6), "r"(s7), "r"(s8));
    // fold 2: (r9:r8) < 2^34 times C = 2^32 + 977:  + (r9:r8) * 977 at limb 0, + (r9:r8) at limb 1
    const uint64_t v = (uint64_t)r8 * K;
    const uint32_t vlo = (uint32_t)v, vhi = (uint32_t)(v >> 32) + r9 * K;
    u256 out;
    uint32_t c1, c2;
    asm("add.cc.u32 %0,%9,%17; addc.cc.u32 %1,%10,%18; addc.cc.u32 %2,%11,0; addc.cc.u32 %3,%12,0;"
        "addc.cc.u32 %4,%13,0; addc.cc.u32 %5,%14,0; addc.cc.u32 %6,%15,0; addc.cc.u32 %7,%16,0; addc.u32 %8,0,0;"
        : "=&r"(out.v[0]), "=&r"(out.v[1]), "=&r"(out.v[2]), "=&r"(out.v[3]), "=&r"(out.v[4]), "=&r"(out.v[5]), "=&r"(out.v[6]),
          "=&r"(out.v[7]), "=&r"(c1)
        : "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(r4), "r"(r5), "r"(r6), "r"(r7), "r"(vlo), "r"(vhi));
    asm("add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%9; addc.cc.u32 %2,%2,0; addc.cc.u32 %3,%3,0;"
        "addc.cc.u32 %4,%4,0; addc.cc.u32 %5,%5,0; addc.cc.u32 %6,%6,0; addc.u32 %7,0,0;"
        : "+r"(out.v[1]), "+r"(out.v[2]), "+r"(out.v[3]), "+r"(out.v[4]), "+r"(out.v[5]), "+r"(out.v[6]), "+r"(out.v[7]), "=r"(c2)
        : "r"(r8), "r"(r9));
    // value = out + (c1 + c2) * 2^256 with c1 + c2 in {0, 1}; if set, out is tiny (< 2^45) so out + C < q
    return cond_sub(out, c1 | c2);
}
#define BP_REDUCE512 reduce512_dev
#else
#define BP_REDUCE512 reduce512
#endif
#if defined(__CUDA_ARCH__) && !defined(BPPP_FQ_INLINE)
// On the device the product and the square are real functions (ptxas keeps the operands in
// registers across the call).  Fully inlined, one mixed add is ~35 KB of SASS and k_msm_gens
// ~800 KB: the hot loop ran out of the 32 KB L1.5 instruction cache and a third of its stall
// samples were "no instruction".  As calls the loop body fits the instruction caches.
static __device__ __noinline__ u256 mul_call(u256 a, u256 b) {
    uint32_t t[16];
    mul_wide_dev(t, a, b);
    return BP_REDUCE512(t);
}
static __device__ __noinline__ u256 sqr_call(u256 a) {
    uint32_t t[16];
    sqr_wide_dev(t, a);
    return BP_REDUCE512(t);
}
BP_D u256 mul(const u256& a, const u256& b) { return mul_call(a, b); }
BP_D u256 sqr(const u256& a) { return sqr_call(a); }
#else
BP_HD u256 mul(const u256& a, const u256& b) {
    uint32_t t[16];
    mul_wide(t, a, b);
    return BP_REDUCE512(t);
}
BP_HD u256 sqr(const u256& a) {
#if defined(__CUDA_ARCH__)
    uint32_t t[16];
    sqr_wide_dev(t, a);
    return BP_REDUCE512(t);
#else
    return mul(a, a);
#endif
}
#endif
// Always-inlined product / square for the few LATENCY-bound kernels (one thread walking a chain of
// dependent group operations: the Horner of a single large MSM, bucket-reduction running sums): inlined,
// the independent multiplications of one group operation interleave their carry chains in the pipeline,
// where a sequence of calls runs them back to back.  Throughput kernels keep the calls (code size).
#if defined(__CUDA_ARCH__)
BP_D u256 mul_inl(const u256& a, const u256& b) {
    uint32_t t[16];
    mul_wide_dev(t, a, b);
    return BP_REDUCE512(t);
}
BP_D u256 sqr_inl(const u256& a) {
    uint32_t t[16];
    sqr_wide_dev(t, a);
    return BP_REDUCE512(t);
}
#else
BP_HD u256 mul_inl(const u256& a, const u256& b) { return mul(a, b); }
BP_HD u256 sqr_inl(const u256& a) { return sqr(a); }
#endif
BP_HD u256 mul_small(const u256& a, uint32_t k) {   // k < 2^16
    uint64_t c = 0;
    u256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] * k;
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    // c * 2^256 == c * C
    uint64_t lo = c * BP_FQ_C0;
    u256 f = u256_zero();
    f.v[0] = (uint32_t)lo;
    uint64_t m = (lo >> 32) + c;
    f.v[1] = (uint32_t)m;
    f.v[2] = (uint32_t)(m >> 32);
    return add(cond_sub(r, 0), f);
}
BP_HD u256 sqr_n(u256 a, int n) {
    for (int i = 0; i < n; i++) a = sqr(a);
    return a;
}
// a^(q-2): 255 squarings + 15 multiplications
BP_HD u256 inv(const u256& a) {
    u256 x2 = mul(sqr(a), a);
    u256 x3 = mul(sqr(x2), a);
    u256 x6 = mul(sqr_n(x3, 3), x3);
    u256 x9 = mul(sqr_n(x6, 3), x3);
    u256 x11 = mul(sqr_n(x9, 2), x2);
    u256 x22 = mul(sqr_n(x11, 11), x11);
    u256 x44 = mul(sqr_n(x22, 22), x22);
    u256 x88 = mul(sqr_n(x44, 44), x44);
    u256 x176 = mul(sqr_n(x88, 88), x88);
    u256 x220 = mul(sqr_n(x176, 44), x44);
    u256 x223 = mul(sqr_n(x220, 3), x3);
    u256 t = mul(sqr_n(x223, 23), x22);
    t = mul(sqr_n(t, 5), a);
    t = mul(sqr_n(t, 3), x2);
    t = mul(sqr_n(t, 2), a);
    return t;
}
}  // namespace fq
// field-multiplication policies for the group law templates of ec.cuh
struct FqCall {
    static BP_HD u256 mul(const u256& a, const u256& b) { return fq::mul(a, b); }
    static BP_HD u256 sqr(const u256& a) { return fq::sqr(a); }
};
struct FqInl {
    static BP_HD u256 mul(const u256& a, const u256& b) { return fq::mul_inl(a, b); }
    static BP_HD u256 sqr(const u256& a) { return fq::sqr_inl(a); }
};

// ============================================================================ Fr
namespace fr {
BP_HD u256 modulus() {
    u256 p;
    p.v[0] = 0xD0364141u; p.v[1] = 0xBFD25E8Cu; p.v[2] = 0xAF48A03Bu; p.v[3] = 0xBAAEDCE6u;
    p.v[4] = 0xFFFFFFFEu; p.v[5] = 0xFFFFFFFFu; p.v[6] = 0xFFFFFFFFu; p.v[7] = 0xFFFFFFFFu;
    return p;
}
#define BP_FR_N0INV 0x5588B13Fu   // -r^{-1} mod 2^32
// R mod r (Montgomery one) and R^2 mod r
BP_HD u256 one() {
    u256 p;
    p.v[0] = 0x2FC9BEBFu; p.v[1] = 0x402DA173u; p.v[2] = 0x50B75FC4u; p.v[3] = 0x45512319u;
    p.v[4] = 1u; p.v[5] = 0; p.v[6] = 0; p.v[7] = 0;
    return p;
}
BP_HD u256 r2() {
    u256 p;
    p.v[0] = 0x67D7D140u; p.v[1] = 0x896CF214u; p.v[2] = 0x0E7CF878u; p.v[3] = 0x741496C2u;
    p.v[4] = 0x5BCD07C6u; p.v[5] = 0xE697F5E4u; p.v[6] = 0x81C69BC5u; p.v[7] = 0x9D671CD5u;
    return p;
}
BP_HD u256 cond_sub(const u256& a, uint32_t carry) {
    u256 t;
    uint32_t br = u256_sub(t, a, modulus());
    uint32_t m = 0u - ((carry | (br ^ 1u)) & 1u);
    return u256_sel(m, t, a);
}
BP_HD u256 add(const u256& a, const u256& b) {
    u256 s;
    uint32_t c = u256_add(s, a, b);
    return cond_sub(s, c);
}
BP_HD u256 sub(const u256& a, const u256& b) {
    u256 d, t;
    uint32_t br = u256_sub(d, a, b);
    u256_add(t, d, modulus());
    return u256_sel(0u - br, t, d);
}
BP_HD u256 neg(const u256& a) {
    u256 d;
    u256_sub(d, modulus(), a);
    return u256_sel(u256_is_zero(a) ? 0xFFFFFFFFu : 0u, a, d);
}
// Montgomery reduction of a 512-bit value t < r * 2^256
BP_HD u256 redc(uint32_t t[16]) {
    const u256 n = modulus();
    uint32_t top = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t m = t[i] * BP_FR_N0INV;
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)m * n.v[j] + t[i + j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
#pragma unroll
        for (int k = i + 8; k < 16; k++) {
            c += t[k];
            t[k] = (uint32_t)c;
            c >>= 32;
        }
        top += (uint32_t)c;
    }
    u256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[8 + i];
    return cond_sub(r, top);
}
#if defined(__CUDA_ARCH__)
// Montgomery reduction with the same carry-chained mad.lo.cc / madc.hi.cc rows as the multiplier:
// row i adds m_i * r at limb offset i as an even chain (limbs of r with even index, positions
// i..i+7) and an odd chain (positions i+1..i+8); the carry out of each chain is counted in cnt[]
// (one level above the chain) and folded in once at the end.
BP_D u256 redc_dev(uint32_t t[16]) {
    const uint32_t N0 = 0xD0364141u, N1 = 0xBFD25E8Cu, N2 = 0xAF48A03Bu, N3 = 0xBAAEDCE6u, N4 = 0xFFFFFFFEu,
                   N5 = 0xFFFFFFFFu, N6 = 0xFFFFFFFFu, N7 = 0xFFFFFFFFu;
    uint32_t cnt[18];
    uint32_t tt[18];
#pragma unroll
    for (int i = 0; i < 16; i++) tt[i] = t[i];
    tt[16] = 0; tt[17] = 0;
#pragma unroll
    for (int i = 0; i < 18; i++) cnt[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t m = tt[i] * BP_FR_N0INV;
        madc_row(tt[i], tt[i + 1], tt[i + 2], tt[i + 3], tt[i + 4], tt[i + 5], tt[i + 6], tt[i + 7], cnt[i + 8], N0, N2, N4, N6, m);
        madc_row(tt[i + 1], tt[i + 2], tt[i + 3], tt[i + 4], tt[i + 5], tt[i + 6], tt[i + 7], tt[i + 8], cnt[i + 9], N1, N3, N5, N7, m);
    }
    u256 r, c;
#pragma unroll
    for (int i = 0; i < 8; i++) { r.v[i] = tt[8 + i]; c.v[i] = cnt[8 + i]; }
    u256 s;
    uint32_t top = u256_add(s, r, c) + cnt[16];
    return cond_sub(s, top ? 1u : 0u);
}
#endif
#if defined(__CUDA_ARCH__) && !defined(BPPP_FQ_INLINE)
static __device__ __noinline__ u256 mul_call(u256 a, u256 b) {      // a call, like fq::mul (code size)
    uint32_t t[16];
    mul_wide_dev(t, a, b);
    return redc_dev(t);
}
BP_D u256 mul(const u256& a, const u256& b) { return mul_call(a, b); }
// the squaring schedule of the Fq multiplier (36 products instead of 64): Fermat inversions are 252 squarings
static __device__ __noinline__ u256 sqr_call(u256 a) {
    uint32_t t[16];
    sqr_wide_dev(t, a);
    return redc_dev(t);
}
BP_D u256 sqr(const u256& a) { return sqr_call(a); }
#else
BP_HD u256 mul(const u256& a, const u256& b) {
    uint32_t t[16];
    mul_wide(t, a, b);
#if defined(__CUDA_ARCH__)
    return redc_dev(t);
#else
    return redc(t);
#endif
}
BP_HD u256 sqr(const u256& a) { return mul(a, a); }
#endif
BP_HD u256 to_mont(const u256& a) { return mul(a, r2()); }
BP_HD u256 from_mont(const u256& a) {
    uint32_t t[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { t[i] = a.v[i]; t[8 + i] = 0; }
    return redc(t);
}
BP_HD u256 dbl(const u256& a) { return add(a, a); }
// a^(r-2) in Montgomery form, 4-bit windows; inv(0) = 0 (batchInverse maps 0 to 0, BatchInverse.hs:14-24)
BP_HD u256 inv(const u256& a) {
    const uint32_t E[8] = {0xD036413Fu, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    u256 tb[16];
    tb[0] = one();
    tb[1] = a;
    for (int i = 2; i < 16; i++) tb[i] = mul(tb[i - 1], a);
    u256 acc = tb[15];                                   // top nibble of r-2 is 0xF
    for (int k = 62; k >= 0; k--) {
        acc = sqr(sqr(sqr(sqr(acc))));
        uint32_t nib = (E[k >> 3] >> ((k & 7) * 4)) & 15u;
        if (nib) acc = mul(acc, tb[nib]);
    }
    return acc;
}
}  // namespace fr

}  // namespace bppp
