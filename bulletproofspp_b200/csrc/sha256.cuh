// SHA-256 on the device (sm_100a) and the byte-level encodings of the reference's Fiat-Shamir
// transcript -- SURVEY 8 f4 ("transcript on device"):
//   hash              app/Main.hs:64-65      SHA256 (cryptohash-sha256), standard FIPS 180-4
//   show of a field   app/Main.hs:79         decimal rendering of the coordinate, "P <dec>" or "<dec>"
//   digest -> field   src/Encoding.hs:75-79  four big-endian Word64, FIRST word least significant, then mod p
// The host implementation of the same functions is csrc/host/{sha256,transcript}.hpp; tests compare the two
// and the oracle (oracle/transcript.py) byte for byte.
#pragma once
#include "fp.cuh"

namespace bppp {
namespace dsha {

__device__ __constant__ uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

__device__ __forceinline__ void init(uint32_t st[8]) {
    st[0] = 0x6a09e667; st[1] = 0xbb67ae85; st[2] = 0x3c6ef372; st[3] = 0xa54ff53a;
    st[4] = 0x510e527f; st[5] = 0x9b05688c; st[6] = 0x1f83d9ab; st[7] = 0x5be0cd19;
}
// one 64-byte block, w[0..15] = the block as big-endian words (destroyed)
__device__ __forceinline__ void compress(uint32_t st[8], uint32_t w[16]) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            const uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
            const uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
            w[i & 15] += s0 + w[(i + 9) & 15] + s1;
        }
        const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        const uint32_t ch = (e & f) ^ (~e & g);
        const uint32_t t1 = h + S1 + ch + K256[i] + w[i & 15];
        const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        const uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// The two halves of `compress`, for hashing ONE long message as fast as one thread can (k_tr_squeeze_coop): the
// message schedule of a block depends on the block alone, so other threads expand it ahead of time
// (kw[i] = K[i] + W[i]) and the thread that owns the chaining value runs only the 64 rounds -- 15 instructions per
// round with a dependent chain of three (funnel shift, xor3, three-input add) through e.
__device__ __forceinline__ void expand_kw(uint32_t w[16], uint32_t* __restrict__ kw) {
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            const uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
            const uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
            w[i & 15] += s0 + w[(i + 9) & 15] + s1;
        }
        kw[i] = w[i & 15] + K256[i];
    }
}
__device__ __forceinline__ void rounds_kw(uint32_t st[8], const uint32_t* __restrict__ kw) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i4 = 0; i4 < 16; i4++) {
        const uint4 q = *reinterpret_cast<const uint4*>(kw + 4 * i4);
        const uint32_t k4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t x = h + k4[r];                              // off the chain: h was e three rounds ago
            const uint32_t dx = d + x;
            const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t en = dx + S1 + ch;
            const uint32_t an = (x + S0 + mj) + (S1 + ch);
            h = g; g = f; f = e; e = en; d = c; c = b; b = a; a = an;
        }
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// ... and the same two halves for 32 hashes side by side: cell i4 of a lane holds K + W of rounds 4 i4 .. 4 i4 + 3, cells
// of consecutive lanes are consecutive in shared memory (stride 32 cells between i4 and i4 + 1)
__device__ __forceinline__ void expand_kw_cells(uint32_t w[16], uint4* __restrict__ cell) {
#pragma unroll
    for (int i4 = 0; i4 < 16; i4++) {
        uint32_t o[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int i = 4 * i4 + r;
            if (i >= 16) {
                const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
                const uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
                const uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
                w[i & 15] += s0 + w[(i + 9) & 15] + s1;
            }
            o[r] = w[i & 15] + K256[i];
        }
        cell[32 * i4] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
__device__ __forceinline__ void rounds_kw_cells(uint32_t st[8], const uint4* __restrict__ cell) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i4 = 0; i4 < 16; i4++) {
        const uint4 q = cell[32 * i4];
        const uint32_t k4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t x = h + k4[r];
            const uint32_t dx = d + x;
            const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t en = dx + S1 + ch;
            const uint32_t an = (x + S0 + mj) + (S1 + ch);
            h = g; g = f; f = e; e = en; d = c; c = b; b = a; a = an;
        }
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// Streaming hasher of one thread.  The 64-byte block buffer lives in SHARED memory, word-major across the
// CTA (blk[word * blockDim.x + tid]: conflict-free) so that the running word index can be dynamic; bytes
// arrive through a 64-bit shift register, so pieces of any length and alignment can be appended.
struct Stream {
    uint32_t st[8];
    uint32_t* blk;           // this thread's column of the CTA's block buffer
    uint32_t stride;         // blockDim.x
    uint64_t pend;           // pending bytes, left-aligned
    uint32_t npend;          // 0..3
    uint32_t widx;           // words already in the current block (0..15)
    uint64_t total;          // bytes so far

    __device__ __forceinline__ void begin(uint32_t* cta_buf) {
        init(st);
        blk = cta_buf + threadIdx.x;
        stride = blockDim.x;
        pend = 0; npend = 0; widx = 0; total = 0;
    }
    __device__ __forceinline__ void flush_block() {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = blk[i * stride];
        compress(st, w);
        widx = 0;
    }
    __device__ __forceinline__ void push_word(uint32_t w) {
        blk[widx * stride] = w;
        if (++widx == 16) flush_block();
    }
    // k (1..4) bytes, left-aligned in `be` (first byte in bits 31..24); bytes beyond k must be zero
    __device__ __forceinline__ void put(uint32_t be, uint32_t k) {
        pend |= (uint64_t)be << (32 - 8 * npend);
        npend += k;
        total += k;
        if (npend >= 4) {
            push_word((uint32_t)(pend >> 32));
            pend <<= 32;
            npend -= 4;
        }
    }
    __device__ __forceinline__ void put_byte(uint32_t b) { put(b << 24, 1); }
    // `len` bytes starting at a 4-byte aligned address (memory order = message order)
    __device__ __forceinline__ void put_aligned(const uint32_t* src, uint32_t len) {
        uint32_t i = 0;
        for (; i + 4 <= len; i += 4) put(__byte_perm(src[i >> 2], 0, 0x0123), 4);
        if (i < len) {
            const uint32_t k = len - i;
            put(__byte_perm(src[i >> 2], 0, 0x0123) & (0xffffffffu << (8 * (4 - k))), k);
        }
    }
    // decimal rendering of a small unsigned integer (`show n`)
    __device__ __forceinline__ void put_uint(uint64_t x) {
        char buf[20];
        int n = 0;
        do { buf[n++] = (char)('0' + (int)(x % 10)); x /= 10; } while (x);
        while (n) put_byte((uint32_t)(unsigned char)buf[--n]);
    }
    __device__ __forceinline__ void finish(uint32_t digest[8]) {
        const uint64_t bits = total * 8;
        put_byte(0x80);
        while (npend != 0) put(0, 1);                     // pad to a word
        total = 0;                                        // (padding is not message length)
        while (widx != 14) push_word(0);
        push_word((uint32_t)(bits >> 32));
        push_word((uint32_t)bits);                        // widx wraps to 0 -> compressed
#pragma unroll
        for (int i = 0; i < 8; i++) digest[i] = st[i];
    }
};

// digest (big-endian words d[0..7]) -> integer as 8 x 32-bit little-endian limbs (Encoding.hs:75-79):
// Word64 i = (d[2i] << 32) | d[2i+1], the FIRST Word64 is the least significant
__device__ __forceinline__ u256 digest_to_u256(const uint32_t d[8]) {
    u256 r;
#pragma unroll
    for (int i = 0; i < 4; i++) { r.v[2 * i] = d[2 * i + 1]; r.v[2 * i + 1] = d[2 * i]; }
    return r;
}
// ... mod r / mod q (the value is < 2^256 < 2 * modulus: one conditional subtraction)
__device__ __forceinline__ u256 digest_to_fr(const uint32_t d[8]) { return fr::cond_sub(digest_to_u256(d), 0); }
__device__ __forceinline__ u256 digest_to_fq(const uint32_t d[8]) { return fq::cond_sub(digest_to_u256(d), 0); }

// Decimal rendering of a canonical 256-bit integer into `out` (at most 78 digits, no terminator), most
// significant digit first, no leading zeros ("0" for zero); returns the length.  Nine digits at a time:
// repeated division of the 8-limb number by 10^9.
__device__ __forceinline__ int decimal_u256(const u256& a, unsigned char* out) {
    uint32_t t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = a.v[i];
    uint32_t chunks[9];
    int nc = 0, top = 7;
    while (top >= 0 && t[top] == 0) top--;
    while (top >= 0) {
        uint64_t rem = 0;
        for (int i = top; i >= 0; i--) {
            const uint64_t cur = (rem << 32) | t[i];
            t[i] = (uint32_t)(cur / 1000000000u);
            rem = cur % 1000000000u;
        }
        chunks[nc++] = (uint32_t)rem;
        while (top >= 0 && t[top] == 0) top--;
    }
    if (nc == 0) { out[0] = '0'; return 1; }
    int len = 0;
    for (int c = nc - 1; c >= 0; c--) {
        uint32_t x = chunks[c];
        unsigned char d[9];
#pragma unroll
        for (int k = 8; k >= 0; k--) { d[k] = (unsigned char)('0' + x % 10); x /= 10; }
        int k0 = 0;
        if (c == nc - 1) while (k0 < 8 && d[k0] == '0') k0++;    // no leading zeros on the top chunk
        for (int k = k0; k < 9; k++) out[len++] = d[k];
    }
    return len;
}

}  // namespace dsha
}  // namespace bppp
