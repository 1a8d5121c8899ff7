// Host scalar-field arithmetic (secp256k1 group order r) in 4 x 64-bit Montgomery limbs.
// Used by the host-side range-proof phases and round sequencing; the device twin is fp.cuh.
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>

namespace bppp {
namespace h64 {
typedef unsigned __int128 u128;

static const uint64_t N[4] = {0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL};
static const uint64_t N0INV = 0x4b0dff665588b13fULL;
static const uint64_t ONE_M[4] = {0x402DA1732FC9BEBFULL, 0x4551231950B75FC4ULL, 0x0000000000000001ULL, 0};
static const uint64_t R2_M[4] = {0x896CF21467D7D140ULL, 0x741496C20E7CF878ULL, 0xE697F5E45BCD07C6ULL, 0x9D671CD581C69BC5ULL};

struct Fr {
    uint64_t v[4];
    bool operator==(const Fr& o) const { return !((v[0] ^ o.v[0]) | (v[1] ^ o.v[1]) | (v[2] ^ o.v[2]) | (v[3] ^ o.v[3])); }
    bool operator!=(const Fr& o) const { return !(*this == o); }
    bool is_zero() const { return !(v[0] | v[1] | v[2] | v[3]); }
};

inline Fr zero() { Fr r = {{0, 0, 0, 0}}; return r; }
inline Fr one() { Fr r; memcpy(r.v, ONE_M, 32); return r; }

inline bool geq_n(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--)
        if (a[i] != N[i]) return a[i] > N[i];
    return true;
}
inline void sub_n(uint64_t a[4]) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - N[i] - br;
        a[i] = (uint64_t)d;
        br = (d >> 127) & 1;
    }
}
inline Fr add(const Fr& a, const Fr& b) {
    Fr r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a.v[i] + b.v[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
    }
    if (c || geq_n(r.v)) sub_n(r.v);
    return r;
}
inline Fr sub(const Fr& a, const Fr& b) {
    Fr r;
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.v[i] - b.v[i] - br;
        r.v[i] = (uint64_t)d;
        br = (d >> 127) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)r.v[i] + N[i];
            r.v[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}
inline Fr neg(const Fr& a) { return a.is_zero() ? a : sub(zero(), a); }
// CIOS Montgomery product (portable form; the reference point for the MULX/ADX one below)
inline Fr mul_portable(const Fr& a, const Fr& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a.v[j] * b.v[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * N0INV;
        c = (u128)m * N[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * N[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fr r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_n(r.v)) sub_n(r.v);
    return r;
}
#if defined(__x86_64__) && !defined(BPPP_HOST_PORTABLE_FR)
#define BPPP_HOST_FR_ADX 1
// One row  (t0..t5) += x[0..3] * y  with MULX and the two carry chains of ADCX / ADOX
// (BMI2 + ADX; bppp_init refuses to start on a CPU without them).  t5 collects both carry-outs.
#define BPPP_FR_ROW(T0, T1, T2, T3, T4, T5, X0, X1, X2, X3, Y)                                        \
    asm("xorl %%eax, %%eax\n\t"                                                                       \
        "mulx %[x0], %%r8, %%r9\n\t adcx %%r8, %[t0]\n\t adox %%r9, %[t1]\n\t"                        \
        "mulx %[x1], %%r8, %%r9\n\t adcx %%r8, %[t1]\n\t adox %%r9, %[t2]\n\t"                        \
        "mulx %[x2], %%r8, %%r9\n\t adcx %%r8, %[t2]\n\t adox %%r9, %[t3]\n\t"                        \
        "mulx %[x3], %%r8, %%r9\n\t adcx %%r8, %[t3]\n\t adox %%r9, %[t4]\n\t"                        \
        "adcx %%rax, %[t4]\n\t adox %%rax, %[t5]\n\t adcx %%rax, %[t5]\n\t"                            \
        : [t0] "+r"(T0), [t1] "+r"(T1), [t2] "+r"(T2), [t3] "+r"(T3), [t4] "+r"(T4), [t5] "+r"(T5)      \
        : [x0] "rm"(X0), [x1] "rm"(X1), [x2] "rm"(X2), [x3] "rm"(X3), "d"(Y)                            \
        : "rax", "r8", "r9", "cc")
inline Fr mul(const Fr& a, const Fr& b) {
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7 = 0, t8 = 0, t9 = 0;
    const uint64_t a0 = a.v[0], a1 = a.v[1], a2 = a.v[2], a3 = a.v[3];
    const uint64_t n0 = N[0], n1 = N[1], n2 = N[2], n3 = N[3];
    // each outer step adds a*b_i, then m*N (which zeroes the lowest limb) and moves up one limb;
    // the running value stays below 2N, so the limb above the window is 0 or 1
    BPPP_FR_ROW(t0, t1, t2, t3, t4, t5, a0, a1, a2, a3, b.v[0]);
    BPPP_FR_ROW(t0, t1, t2, t3, t4, t5, n0, n1, n2, n3, t0 * N0INV);
    BPPP_FR_ROW(t1, t2, t3, t4, t5, t6, a0, a1, a2, a3, b.v[1]);
    BPPP_FR_ROW(t1, t2, t3, t4, t5, t6, n0, n1, n2, n3, t1 * N0INV);
    BPPP_FR_ROW(t2, t3, t4, t5, t6, t7, a0, a1, a2, a3, b.v[2]);
    BPPP_FR_ROW(t2, t3, t4, t5, t6, t7, n0, n1, n2, n3, t2 * N0INV);
    BPPP_FR_ROW(t3, t4, t5, t6, t7, t8, a0, a1, a2, a3, b.v[3]);
    BPPP_FR_ROW(t3, t4, t5, t6, t7, t8, n0, n1, n2, n3, t3 * N0INV);
    (void)t9;
    Fr r = {{t4, t5, t6, t7}};
    if (t8 || geq_n(r.v)) sub_n(r.v);
    return r;
}
inline bool host_cpu_ok() { return __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("adx"); }
#else
inline Fr mul(const Fr& a, const Fr& b) { return mul_portable(a, b); }
inline bool host_cpu_ok() { return true; }
#endif
inline Fr sqr(const Fr& a) { return mul(a, a); }
inline Fr dbl(const Fr& a) { return add(a, a); }
inline Fr from_canon(const uint64_t c[4]) {
    Fr a = {{c[0], c[1], c[2], c[3]}}, r2;
    memcpy(r2.v, R2_M, 32);
    return mul(a, r2);
}
inline Fr from_bytes(const uint8_t* b) {       // 32-byte LE canonical
    uint64_t c[4];
    memcpy(c, b, 32);
    return from_canon(c);
}
inline void to_canon(uint64_t c[4], const Fr& a) {
    Fr o = {{1, 0, 0, 0}};
    Fr r = mul(a, o);
    memcpy(c, r.v, 32);
}
inline void to_bytes(uint8_t* b, const Fr& a) {
    uint64_t c[4];
    to_canon(c, a);
    memcpy(b, c, 32);
}
inline Fr from_u64(uint64_t x) {
    uint64_t c[4] = {x, 0, 0, 0};
    return from_canon(c);
}
inline Fr from_u128(u128 x) {
    uint64_t c[4] = {(uint64_t)x, (uint64_t)(x >> 64), 0, 0};
    return from_canon(c);
}
inline Fr from_i128(__int128 x) { return x < 0 ? neg(from_u128((u128)(-x))) : from_u128((u128)x); }
// reduce an arbitrary 256-bit integer (e.g. a SHA-256 digest) mod r
inline Fr from_wide(const uint64_t c[4]) {
    uint64_t t[4] = {c[0], c[1], c[2], c[3]};
    if (geq_n(t)) sub_n(t);                    // 2^256 < 2r
    return from_canon(t);
}
inline Fr pow_u64(Fr base, uint64_t e) {
    Fr acc = one();
    while (e) {
        if (e & 1) acc = mul(acc, base);
        e >>= 1;
        if (e) base = sqr(base);
    }
    return acc;
}
inline Fr inv(const Fr& a) {                   // a^(r-2); inv(0) = 0
    uint64_t e[4] = {N[0] - 2, N[1], N[2], N[3]};
    Fr acc = one();
    for (int i = 255; i >= 0; i--) {
        acc = sqr(acc);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a);
    }
    return acc;
}
// batchInverse, 0 -> 0 (src/Data/Field/BatchInverse.hs:14-24)
inline void batch_inv(Fr* v, size_t n) {
    std::vector<Fr> pre(n);
    Fr acc = one();
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        if (!v[i].is_zero()) acc = mul(acc, v[i]);
    }
    Fr y = inv(acc);
    for (size_t i = n; i-- > 0;) {
        if (v[i].is_zero()) continue;
        Fr x = v[i];
        v[i] = mul(y, pre[i]);
        y = mul(y, x);
    }
}
inline std::vector<Fr> batch_inv(std::vector<Fr> v) {
    batch_inv(v.data(), v.size());
    return v;
}
// a, a^2, ... (powers', src/Utils.hs:107-108)
inline std::vector<Fr> powers1(const Fr& a, size_t n) {
    std::vector<Fr> o(n);
    Fr x = a;
    for (size_t i = 0; i < n; i++) { o[i] = x; x = mul(x, a); }
    return o;
}

}  // namespace h64
}  // namespace bppp
