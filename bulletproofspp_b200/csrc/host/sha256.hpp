// SHA-256 (FIPS 180-4) for the host-side Fiat-Shamir transcript (app/Main.hs:64-65 uses
// cryptohash-sha256).  SHA-NI path when the CPU has it, portable path otherwise.
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
#endif

namespace bppp {
namespace sha {

static const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
    0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
    0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
    0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
    0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
    0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
    0xc67178f2};

inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

inline void compress_portable(uint32_t st[8], const uint8_t* p, size_t nblk) {
    for (; nblk; nblk--, p += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; i++)
            w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
        for (int i = 0; i < 64; i++) {
            uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = h + S1 + ch + K[i] + w[i];
            uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
    }
}

#if defined(__x86_64__)
__attribute__((target("sha,sse4.1,ssse3"))) inline void compress_shani(uint32_t st[8], const uint8_t* p, size_t nblk) {
    const __m128i MASK = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i TMP = _mm_loadu_si128((const __m128i*)&st[0]);
    __m128i STATE1 = _mm_loadu_si128((const __m128i*)&st[4]);
    TMP = _mm_shuffle_epi32(TMP, 0xB1);
    STATE1 = _mm_shuffle_epi32(STATE1, 0x1B);
    __m128i STATE0 = _mm_alignr_epi8(TMP, STATE1, 8);
    STATE1 = _mm_blend_epi16(STATE1, TMP, 0xF0);
    for (; nblk; nblk--, p += 64) {
        __m128i ABEF = STATE0, CDGH = STATE1;
        __m128i M[4];
        for (int i = 0; i < 4; i++) M[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(p + 16 * i)), MASK);
        for (int r = 0; r < 16; r++) {
            __m128i msg = _mm_add_epi32(M[r & 3], _mm_loadu_si128((const __m128i*)&K[4 * r]));
            STATE1 = _mm_sha256rnds2_epu32(STATE1, STATE0, msg);
            if (r >= 3 && r < 15) {          // schedule word group r+1 from groups r-3 .. r
                __m128i t = _mm_alignr_epi8(M[r & 3], M[(r + 3) & 3], 4);
                M[(r + 1) & 3] = _mm_sha256msg2_epu32(_mm_add_epi32(M[(r + 1) & 3], t), M[r & 3]);
            }
            msg = _mm_shuffle_epi32(msg, 0x0E);
            STATE0 = _mm_sha256rnds2_epu32(STATE0, STATE1, msg);
            if (r >= 1 && r < 13) M[(r + 3) & 3] = _mm_sha256msg1_epu32(M[(r + 3) & 3], M[r & 3]);
        }
        STATE0 = _mm_add_epi32(STATE0, ABEF);
        STATE1 = _mm_add_epi32(STATE1, CDGH);
    }
    TMP = _mm_shuffle_epi32(STATE0, 0x1B);
    STATE1 = _mm_shuffle_epi32(STATE1, 0xB1);
    STATE0 = _mm_blend_epi16(TMP, STATE1, 0xF0);
    STATE1 = _mm_alignr_epi8(STATE1, TMP, 8);
    _mm_storeu_si128((__m128i*)&st[0], STATE0);
    _mm_storeu_si128((__m128i*)&st[4], STATE1);
}
// two independent messages, block for block: SHA256RNDS2 is latency-bound, so interleaving a
// second stream nearly doubles the throughput of one core (used for pairs of transcripts)
__attribute__((target("sha,sse4.1,ssse3"))) inline void compress_shani_x2(uint32_t sa[8], const uint8_t* pa, uint32_t sb[8],
                                                                           const uint8_t* pb, size_t nblk) {
    const __m128i MASK = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i TA = _mm_shuffle_epi32(_mm_loadu_si128((const __m128i*)&sa[0]), 0xB1);
    __m128i A1 = _mm_shuffle_epi32(_mm_loadu_si128((const __m128i*)&sa[4]), 0x1B);
    __m128i A0 = _mm_alignr_epi8(TA, A1, 8);
    A1 = _mm_blend_epi16(A1, TA, 0xF0);
    __m128i TB = _mm_shuffle_epi32(_mm_loadu_si128((const __m128i*)&sb[0]), 0xB1);
    __m128i B1 = _mm_shuffle_epi32(_mm_loadu_si128((const __m128i*)&sb[4]), 0x1B);
    __m128i B0 = _mm_alignr_epi8(TB, B1, 8);
    B1 = _mm_blend_epi16(B1, TB, 0xF0);
    for (; nblk; nblk--, pa += 64, pb += 64) {
        const __m128i SA0 = A0, SA1 = A1, SB0 = B0, SB1 = B1;
        __m128i MA[4], MB[4];
        for (int i = 0; i < 4; i++) {
            MA[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(pa + 16 * i)), MASK);
            MB[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(pb + 16 * i)), MASK);
        }
#pragma GCC unroll 16
        for (int r = 0; r < 16; r++) {
            const __m128i k = _mm_loadu_si128((const __m128i*)&K[4 * r]);
            __m128i ma = _mm_add_epi32(MA[r & 3], k), mb = _mm_add_epi32(MB[r & 3], k);
            A1 = _mm_sha256rnds2_epu32(A1, A0, ma);
            B1 = _mm_sha256rnds2_epu32(B1, B0, mb);
            if (r >= 3 && r < 15) {
                __m128i ta = _mm_alignr_epi8(MA[r & 3], MA[(r + 3) & 3], 4);
                __m128i tb = _mm_alignr_epi8(MB[r & 3], MB[(r + 3) & 3], 4);
                MA[(r + 1) & 3] = _mm_sha256msg2_epu32(_mm_add_epi32(MA[(r + 1) & 3], ta), MA[r & 3]);
                MB[(r + 1) & 3] = _mm_sha256msg2_epu32(_mm_add_epi32(MB[(r + 1) & 3], tb), MB[r & 3]);
            }
            ma = _mm_shuffle_epi32(ma, 0x0E);
            mb = _mm_shuffle_epi32(mb, 0x0E);
            A0 = _mm_sha256rnds2_epu32(A0, A1, ma);
            B0 = _mm_sha256rnds2_epu32(B0, B1, mb);
            if (r >= 1 && r < 13) {
                MA[(r + 3) & 3] = _mm_sha256msg1_epu32(MA[(r + 3) & 3], MA[r & 3]);
                MB[(r + 3) & 3] = _mm_sha256msg1_epu32(MB[(r + 3) & 3], MB[r & 3]);
            }
        }
        A0 = _mm_add_epi32(A0, SA0); A1 = _mm_add_epi32(A1, SA1);
        B0 = _mm_add_epi32(B0, SB0); B1 = _mm_add_epi32(B1, SB1);
    }
    TA = _mm_shuffle_epi32(A0, 0x1B);
    A1 = _mm_shuffle_epi32(A1, 0xB1);
    _mm_storeu_si128((__m128i*)&sa[0], _mm_blend_epi16(TA, A1, 0xF0));
    _mm_storeu_si128((__m128i*)&sa[4], _mm_alignr_epi8(A1, TA, 8));
    TB = _mm_shuffle_epi32(B0, 0x1B);
    B1 = _mm_shuffle_epi32(B1, 0xB1);
    _mm_storeu_si128((__m128i*)&sb[0], _mm_blend_epi16(TB, B1, 0xF0));
    _mm_storeu_si128((__m128i*)&sb[4], _mm_alignr_epi8(B1, TB, 8));
}
inline bool has_shani() {
    static int cached = -1;
    if (cached < 0) {
        unsigned a, b, c, d;
        cached = 0;
        if (__get_cpuid_count(7, 0, &a, &b, &c, &d)) cached = (b >> 29) & 1;
        if (cached && __get_cpuid(1, &a, &b, &c, &d)) cached = ((c >> 19) & 1) && ((c >> 9) & 1);   // sse4.1, ssse3
    }
    return cached == 1;
}
#endif

inline void compress(uint32_t st[8], const uint8_t* p, size_t nblk) {
#if defined(__x86_64__)
    if (has_shani()) { compress_shani(st, p, nblk); return; }
#endif
    compress_portable(st, p, nblk);
}

// one-shot digest over up to three concatenated pieces (avoids building the message)
inline void digest3(uint8_t out[32], const uint8_t* a, size_t na, const uint8_t* b, size_t nb, const uint8_t* c, size_t nc) {
    uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint8_t buf[128];
    size_t fill = 0;
    uint64_t total = (uint64_t)na + nb + nc;
    const uint8_t* ps[3] = {a, b, c};
    size_t ns[3] = {na, nb, nc};
    for (int k = 0; k < 3; k++) {
        const uint8_t* p = ps[k];
        size_t n = ns[k];
        if (fill) {
            size_t take = 64 - fill < n ? 64 - fill : n;
            memcpy(buf + fill, p, take);
            fill += take; p += take; n -= take;
            if (fill == 64) { compress(st, buf, 1); fill = 0; }
        }
        if (n >= 64) {
            size_t blk = n / 64;
            compress(st, p, blk);
            p += blk * 64; n -= blk * 64;
        }
        if (n) { memcpy(buf + fill, p, n); fill += n; }
    }
    buf[fill++] = 0x80;
    size_t padto = fill <= 56 ? 64 : 128;
    memset(buf + fill, 0, padto - fill);
    uint64_t bits = total * 8;
    for (int i = 0; i < 8; i++) buf[padto - 1 - i] = (uint8_t)(bits >> (8 * i));
    compress(st, buf, padto / 64);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(st[i] >> 24); out[4 * i + 1] = (uint8_t)(st[i] >> 16);
        out[4 * i + 2] = (uint8_t)(st[i] >> 8); out[4 * i + 3] = (uint8_t)st[i];
    }
}
// Streaming form used for PAIRS of messages: both are cut into pieces (prefix, body, ...) and
// advanced together with the two-stream compressor; whatever does not line up falls back to the
// single-stream one.
struct Stream {
    uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint8_t buf[128];
    size_t fill = 0;
    uint64_t total = 0;
    const uint8_t* p = nullptr;          // current piece
    size_t n = 0;
    void piece(const uint8_t* q, size_t len) { p = q; n = len; total += len; }
    // top up the partial block from the current piece; true when a full block sits in buf
    void top_up() {
        if (fill) {
            size_t take = 64 - fill < n ? 64 - fill : n;
            memcpy(buf + fill, p, take);
            fill += take; p += take; n -= take;
            if (fill == 64) { compress(st, buf, 1); fill = 0; }
        }
    }
    void stash() { if (n) { memcpy(buf + fill, p, n); fill += n; p += n; n = 0; } }
    void finish(uint8_t out[32]) {
        buf[fill++] = 0x80;
        size_t padto = fill <= 56 ? 64 : 128;
        memset(buf + fill, 0, padto - fill);
        uint64_t bits = total * 8;
        for (int i = 0; i < 8; i++) buf[padto - 1 - i] = (uint8_t)(bits >> (8 * i));
        compress(st, buf, padto / 64);
        for (int i = 0; i < 8; i++) {
            out[4 * i] = (uint8_t)(st[i] >> 24); out[4 * i + 1] = (uint8_t)(st[i] >> 16);
            out[4 * i + 2] = (uint8_t)(st[i] >> 8); out[4 * i + 3] = (uint8_t)st[i];
        }
    }
};
// digests of (a0 | a1) and (b0 | b1)
inline void digest2x2(uint8_t outa[32], const uint8_t* a0, size_t na0, const uint8_t* a1, size_t na1, uint8_t outb[32],
                      const uint8_t* b0, size_t nb0, const uint8_t* b1, size_t nb1) {
    Stream A, B;
    const uint8_t* pa[2] = {a0, a1};
    const uint8_t* pb[2] = {b0, b1};
    size_t la[2] = {na0, na1}, lb[2] = {nb0, nb1};
    for (int k = 0; k < 2; k++) {
        A.piece(pa[k], la[k]);
        B.piece(pb[k], lb[k]);
        A.top_up();
        B.top_up();
        size_t blk = (A.n < B.n ? A.n : B.n) / 64;
#if defined(__x86_64__)
        if (blk && has_shani()) {
            compress_shani_x2(A.st, A.p, B.st, B.p, blk);
            A.p += blk * 64; A.n -= blk * 64;
            B.p += blk * 64; B.n -= blk * 64;
        }
#endif
        if (A.n >= 64) { size_t b = A.n / 64; compress(A.st, A.p, b); A.p += b * 64; A.n -= b * 64; }
        if (B.n >= 64) { size_t b = B.n / 64; compress(B.st, B.p, b); B.p += b * 64; B.n -= b * 64; }
        A.stash();
        B.stash();
    }
    A.finish(outa);
    B.finish(outb);
}
// two messages of at most 55 bytes each (one padded block): the RNG's hash(seed <> show counter)
inline void digest_short_x2(uint8_t outa[32], const uint8_t* a, size_t na, uint8_t outb[32], const uint8_t* b, size_t nb) {
    static const uint32_t IV[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t sa[8], sb[8];
    memcpy(sa, IV, 32);
    memcpy(sb, IV, 32);
    uint8_t ba[64], bb[64];
    memset(ba, 0, 64);
    memset(bb, 0, 64);
    memcpy(ba, a, na); ba[na] = 0x80; ba[62] = (uint8_t)((na * 8) >> 8); ba[63] = (uint8_t)(na * 8);
    memcpy(bb, b, nb); bb[nb] = 0x80; bb[62] = (uint8_t)((nb * 8) >> 8); bb[63] = (uint8_t)(nb * 8);
#if defined(__x86_64__)
    if (has_shani()) compress_shani_x2(sa, ba, sb, bb, 1);
    else
#endif
    { compress_portable(sa, ba, 1); compress_portable(sb, bb, 1); }
    for (int i = 0; i < 8; i++) {
        outa[4 * i] = (uint8_t)(sa[i] >> 24); outa[4 * i + 1] = (uint8_t)(sa[i] >> 16);
        outa[4 * i + 2] = (uint8_t)(sa[i] >> 8); outa[4 * i + 3] = (uint8_t)sa[i];
        outb[4 * i] = (uint8_t)(sb[i] >> 24); outb[4 * i + 1] = (uint8_t)(sb[i] >> 16);
        outb[4 * i + 2] = (uint8_t)(sb[i] >> 8); outb[4 * i + 3] = (uint8_t)sb[i];
    }
}
inline void digest(uint8_t out[32], const std::string& s) {
    digest3(out, (const uint8_t*)s.data(), s.size(), nullptr, 0, nullptr, 0);
}

}  // namespace sha
}  // namespace bppp
