// Host-side Fiat-Shamir transcript, RNG and generator derivation, exactly as the reference's CLI
// wires them (the north star keeps the transcript on the host):
//   hash / getPoints / shaOracle / hashToScalar(s)   app/Main.hs:64-87
//   ZKPT (commitment list newest-first + random ctr)  src/ZKP.hs:68-101
//   digest -> field element                           src/Encoding.hs:75-79
// Two behaviours live in un-vendored packages and are policies here ("parity
// unpinned"): how `show` renders a `Prime p` ("P <dec>" by default, or bare "<dec>") and which
// square root `pointX` returns (rhs^((q+1)/4) by default).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "../ec.cuh"
#include "fr64.hpp"
#include "sha256.hpp"

namespace bppp {
namespace tr {
using h64::Fr;
typedef unsigned __int128 u128;

enum ShowFormat { PREFIXED_P = 0, BARE_DECIMAL = 1 };
enum RootPolicy { ROOT_EXP = 0, ROOT_EVEN = 1, ROOT_SMALLER = 2 };

// decimal rendering of a 256-bit little-endian integer (4 x u64)
inline void append_decimal(std::string& out, const uint64_t v[4]) {
    // peel 19 decimal digits at a time (2^64 > 10^19), then emit digits two at a time
    static const char PAIRS[] =
        "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
        "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
    uint64_t t[4] = {v[0], v[1], v[2], v[3]};
    uint64_t chunks[5];
    int nc = 0;
    const uint64_t TEN19 = 10000000000000000000ULL;
    int top = 3;
    while (top >= 0 && t[top] == 0) top--;
    while (top >= 0) {
        // (rem, t[i]) / 10^19 by multiplication with the precomputed reciprocal of the (normalised)
        // divisor -- Moeller & Granlund, "Improved division by invariant integers", algorithm 4
        const uint64_t RECIP = 0xd83c94fb6d2ac34aULL;             // floor((2^128 - 1) / 10^19) - 2^64
        uint64_t rem = 0;
        for (int i = top; i >= 0; i--) {
            const uint64_t u1 = rem, u0 = t[i];
            u128 q = (u128)RECIP * u1 + (((u128)u1 << 64) | u0);
            uint64_t q1 = (uint64_t)(q >> 64) + 1, q0 = (uint64_t)q;
            uint64_t r = u0 - q1 * TEN19;
            if (r > q0) { q1--; r += TEN19; }
            if (r >= TEN19) { q1++; r -= TEN19; }
            t[i] = q1;
            rem = r;
        }
        chunks[nc++] = rem;
        while (top >= 0 && t[top] == 0) top--;
    }
    if (nc == 0) { out.push_back('0'); return; }
    char buf[100];
    char* end = buf + sizeof buf;
    char* p = end;
    for (int c = 0; c < nc; c++) {
        uint64_t x = chunks[c];
        char* stop = p - 19;                       // full chunks are zero-padded to 19 digits
        while (x >= 100) { unsigned r = (unsigned)(x % 100); x /= 100; p -= 2; p[0] = PAIRS[2 * r]; p[1] = PAIRS[2 * r + 1]; }
        if (x >= 10) { p -= 2; p[0] = PAIRS[2 * x]; p[1] = PAIRS[2 * x + 1]; }
        else { *--p = (char)('0' + x); }
        if (c + 1 < nc) while (p > stop) *--p = '0';
    }
    out.append(p, end - p);
}
inline void append_show_field(std::string& out, const uint64_t v[4], int fmt) {
    if (fmt == PREFIXED_P) out.append("P ");
    append_decimal(out, v);
}
inline int format_uint(char* dst, uint64_t x) {        // decimal, no terminator; returns the length (<= 20)
    char buf[24];
    char* p = buf + sizeof buf;
    do { *--p = (char)('0' + x % 10); x /= 10; } while (x);
    int n = (int)(buf + sizeof buf - p);
    memcpy(dst, p, n);
    return n;
}
inline void append_uint(std::string& out, uint64_t x) {
    char buf[24];
    out.append(buf, format_uint(buf, x));
}
// digest -> integer: four big-endian Word64, first word least significant (Encoding.hs:75-79)
inline void digest_to_words(uint64_t w[4], const uint8_t d[32]) {
    for (int i = 0; i < 4; i++) {
        uint64_t x = 0;
        for (int k = 0; k < 8; k++) x = (x << 8) | d[8 * i + k];
        w[i] = x;
    }
}
inline Fr hash_to_fr(const std::string& s) {
    uint8_t d[32];
    sha::digest(d, s);
    uint64_t w[4];
    digest_to_words(w, d);
    return h64::from_wide(w);
}
// `coords (A x y) = show x <> show y` of an affine point given as 64 LE bytes (app/Main.hs:78-80)
inline std::string show_point(const uint8_t p[64], int fmt) {
    std::string s;
    s.reserve(170);
    uint64_t x[4], y[4];
    memcpy(x, p, 32);
    memcpy(y, p + 32, 32);
    append_show_field(s, x, fmt);
    append_show_field(s, y, fmt);
    return s;
}

// The reference's ZKPT state for one proof.
struct Zkpt {
    int fmt = PREFIXED_P;
    std::string seed;                 // randomSeed; empty + no_random => verifier
    bool no_random = false;
    uint64_t n_random = 0;
    size_t n_coms = 0;
    // concat of coords, NEWEST FIRST (cs' = xs ++ cs): kept at the END of `store` so that new
    // commitments are written in front of it without moving what is already there
    std::string store;
    size_t start = 0;
    const uint8_t* bdata() const { return (const uint8_t*)store.data() + start; }
    size_t bsize() const { return store.size() - start; }
    void prepend(const char* p, size_t len) {
        if (len > start) {                                        // grow at the front
            const size_t used = bsize(), cap = 2 * (used + len) + 8192;
            std::string bigger(cap, '\0');
            memcpy(&bigger[cap - used], bdata(), used);
            store.swap(bigger);
            start = cap - used;
        }
        start -= len;
        memcpy(&store[start], p, len);
    }
    uint64_t hashed_bytes = 0;

    // `random` (ZKP.hs:90-93) with h = hashToScalar rn . show (app/Main.hs:177)
    Fr random() {
        std::string s = seed;
        append_uint(s, n_random++);
        return hash_to_fr(s);
    }
    // the next n values of `random`, two one-block hashes at a time when seed <> counter fits a block
    void random_fill(Fr* out, size_t n) {
        size_t i = 0;
        if (seed.size() + 20 <= 55) {
            uint8_t m0[64], m1[64], d0[32], d1[32];
            memcpy(m0, seed.data(), seed.size());
            memcpy(m1, seed.data(), seed.size());
            for (; i + 1 < n; i += 2) {
                size_t l0 = seed.size() + format_uint((char*)m0 + seed.size(), n_random++);
                size_t l1 = seed.size() + format_uint((char*)m1 + seed.size(), n_random++);
                sha::digest_short_x2(d0, m0, l0, d1, m1, l1);
                uint64_t w[4];
                digest_to_words(w, d0);
                out[i] = h64::from_wide(w);
                digest_to_words(w, d1);
                out[i + 1] = h64::from_wide(w);
            }
        }
        for (; i < n; i++) out[i] = random();
    }
    // the same values as canonical 32-byte little-endian scalars (for blinders that go straight to the device)
    void random_fill_canonical(uint8_t* out, size_t n) {
        size_t i = 0;
        auto put = [](uint8_t* dst, const uint8_t d[32]) {
            uint64_t w[4];
            digest_to_words(w, d);
            if (h64::geq_n(w)) h64::sub_n(w);                 // 2^256 < 2r
            memcpy(dst, w, 32);
        };
        if (seed.size() + 20 <= 55) {
            uint8_t m0[64], m1[64], d0[32], d1[32];
            memcpy(m0, seed.data(), seed.size());
            memcpy(m1, seed.data(), seed.size());
            for (; i + 1 < n; i += 2) {
                size_t l0 = seed.size() + format_uint((char*)m0 + seed.size(), n_random++);
                size_t l1 = seed.size() + format_uint((char*)m1 + seed.size(), n_random++);
                sha::digest_short_x2(d0, m0, l0, d1, m1, l1);
                put(out + 32 * i, d0);
                put(out + 32 * (i + 1), d1);
            }
        }
        for (; i < n; i++) {
            std::string s = seed;
            append_uint(s, n_random++);
            uint8_t d[32];
            sha::digest(d, s);
            put(out + 32 * i, d);
        }
    }
    // `oracle xs` -> first `count` scalars of shaOracle cs' (ZKP.hs:96-101, app/Main.hs:75-80)
    void absorb(const uint8_t* pts, size_t npts) {
        std::string add;
        add.reserve(npts * 170);
        for (size_t i = 0; i < npts; i++) add += show_point(pts + 64 * i, fmt);
        prepend(add.data(), add.size());
        n_coms += npts;
    }
    // scalar i (1-based) of the current transcript: hash(show i <> show (length cs) <> coords)
    static std::string prefix_of(uint64_t i, uint64_t n_coms) {
        std::string pre;
        append_uint(pre, i);
        append_uint(pre, n_coms);
        return pre;
    }
    static Fr digest_to_fr(const uint8_t d[32]) {
        uint64_t w[4];
        digest_to_words(w, d);
        return h64::from_wide(w);
    }
    void squeeze(Fr* out, int count) {
        int i = 1;
        for (; i + 1 <= count; i += 2) {                  // two challenges at a time (two-stream SHA)
            std::string p0 = prefix_of(i, n_coms), p1 = prefix_of(i + 1, n_coms);
            uint8_t d0[32], d1[32];
            sha::digest2x2(d0, (const uint8_t*)p0.data(), p0.size(), bdata(), bsize(), d1,
                           (const uint8_t*)p1.data(), p1.size(), bdata(), bsize());
            hashed_bytes += p0.size() + p1.size() + 2 * bsize();
            out[i - 1] = digest_to_fr(d0);
            out[i] = digest_to_fr(d1);
        }
        if (i <= count) {
            std::string pre = prefix_of(i, n_coms);
            uint8_t d[32];
            sha::digest3(d, (const uint8_t*)pre.data(), pre.size(), bdata(), bsize(), nullptr, 0);
            hashed_bytes += pre.size() + bsize();
            out[i - 1] = digest_to_fr(d);
        }
    }
    void oracle(const uint8_t* pts, size_t npts, Fr* out, int count) {
        absorb(pts, npts);
        squeeze(out, count);
    }
    // the first challenge of two different transcripts, hashed together
    static void oracle_pair(Zkpt& a, const uint8_t* pa, Zkpt& b, const uint8_t* pb, size_t npts, Fr* ea, Fr* eb) {
        a.absorb(pa, npts);
        b.absorb(pb, npts);
        std::string p0 = prefix_of(1, a.n_coms), p1 = prefix_of(1, b.n_coms);
        uint8_t d0[32], d1[32];
        sha::digest2x2(d0, (const uint8_t*)p0.data(), p0.size(), a.bdata(), a.bsize(), d1,
                       (const uint8_t*)p1.data(), p1.size(), b.bdata(), b.bsize());
        a.hashed_bytes += p0.size() + a.bsize();
        b.hashed_bytes += p1.size() + b.bsize();
        *ea = digest_to_fr(d0);
        *eb = digest_to_fr(d1);
    }
    // Verifier side of proveBPM's challenges: all k round messages are known up front, and the
    // transcript after round r is a SUFFIX of the final one (newest first).  pts = the (X,R) pairs
    // in hashing order (oldest round first); out[r] = challenge of round r.
    void oracle_rounds(const uint8_t* const* round_pts, size_t k, Fr* out) {
        if (!k) return;
        // prepend the rounds in hashing order (oldest first): round r's transcript is the last len[r]
        // bytes of the store (lengths from the end stay valid when the store grows at the front)
        std::vector<size_t> len(k);
        {
            std::vector<std::string> shown(k);
            for (size_t r = 0; r < k; r++) shown[r] = show_point(round_pts[r], fmt) + show_point(round_pts[r] + 64, fmt);
            for (size_t r = 0; r < k; r++) { prepend(shown[r].data(), shown[r].size()); len[r] = bsize(); }
        }
        const uint8_t* end = (const uint8_t*)store.data() + store.size();
        size_t r = 0;
        for (; r + 1 < k; r += 2) {
            std::string p0 = prefix_of(1, n_coms + 2 * (r + 1)), p1 = prefix_of(1, n_coms + 2 * (r + 2));
            uint8_t d0[32], d1[32];
            sha::digest2x2(d0, (const uint8_t*)p0.data(), p0.size(), end - len[r], len[r], d1,
                           (const uint8_t*)p1.data(), p1.size(), end - len[r + 1], len[r + 1]);
            hashed_bytes += p0.size() + p1.size() + len[r] + len[r + 1];
            out[r] = digest_to_fr(d0);
            out[r + 1] = digest_to_fr(d1);
        }
        if (r < k) {
            std::string pre = prefix_of(1, n_coms + 2 * (r + 1));
            uint8_t d[32];
            sha::digest3(d, (const uint8_t*)pre.data(), pre.size(), end - len[r], len[r], nullptr, 0);
            hashed_bytes += pre.size() + len[r];
            out[r] = digest_to_fr(d);
        }
        n_coms += 2 * k;
    }
};

// `hashToScalars ("Blinding " <> rn)` position j (1-based)  (app/Main.hs:86-87, 275-276)
inline Fr input_blind(const std::string& random_seed, uint64_t j) {
    std::string s = "Blinding " + random_seed;
    append_uint(s, j);
    return hash_to_fr(s);
}

// getPoints seed (app/Main.hs:68-72): x = hash(seed <> show n) in Fq, kept when x^3 + 7 is a square
inline std::vector<Affine> get_points(const std::string& seed, size_t count, int root_policy) {
    std::vector<Affine> out;
    out.reserve(count);
    // exponent (q + 1) / 4
    u256 e = fq::modulus();
    {   // (q + 1) >> 2
        u256 one = u256_one(), t;
        u256_add(t, e, one);           // q + 1 < 2^256
        for (int i = 0; i < 8; i++) e.v[i] = (t.v[i] >> 2) | (i < 7 ? t.v[i + 1] << 30 : 0);
    }
    for (uint64_t n = 0; out.size() < count; n++) {
        std::string s = seed;
        append_uint(s, n);
        uint8_t d[32];
        sha::digest(d, s);
        uint64_t w[4];
        digest_to_words(w, d);
        u256 x;
        for (int i = 0; i < 4; i++) { x.v[2 * i] = (uint32_t)w[i]; x.v[2 * i + 1] = (uint32_t)(w[i] >> 32); }
        x = fq::cond_sub(x, 0);                           // toP: 2^256 < 2q
        u256 seven = u256_zero();
        seven.v[0] = 7;
        u256 rhs = fq::add(fq::mul(fq::sqr(x), x), seven);
        u256 y = u256_one();
        for (int i = 255; i >= 0; i--) {
            y = fq::sqr(y);
            if (u256_bit(e, i)) y = fq::mul(y, rhs);
        }
        if (!u256_eq(fq::sqr(y), rhs)) continue;
        u256 ny = fq::neg(y);
        if (root_policy == ROOT_EVEN && (y.v[0] & 1)) y = ny;
        else if (root_policy == ROOT_SMALLER && !u256_geq(ny, y)) y = ny;
        Affine p;
        p.x = x; p.y = y;
        out.push_back(p);
    }
    return out;
}

}  // namespace tr
}  // namespace bppp
