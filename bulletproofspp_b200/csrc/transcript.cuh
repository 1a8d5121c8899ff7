// The reference's hash-derived objects on the device (sm_100a) -- SURVEY 8 f4:
//   getPoints seed                 app/Main.hs:68-72    generators: x = hash(seed <> show n), kept when x^3 + 7 is a square
//   shaOracle / oracle / oracle'   app/Main.hs:75-80, src/ZKP.hs:96-101   Fiat-Shamir challenges
//   random                         src/ZKP.hs:90-93, app/Main.hs:177      blinders: hash(randomSeed <> show n)
// Bit-identical to the host implementation (csrc/host/transcript.hpp) and to oracle/transcript.py; the
// `show` format of a field element and the square-root policy are the same switchable policies.
#pragma once
#include "ec.cuh"
#include "sha256.cuh"

namespace bppp {

#define TR_THREADS 128
enum { TR_PREFIXED_P = 0, TR_BARE_DECIMAL = 1 };
enum { TR_ROOT_EXP = 0, TR_ROOT_EVEN = 1, TR_ROOT_SMALLER = 2 };

// a^((q+1)/4): the square root candidate for q = 3 (mod 4); 253 squarings + 13 multiplications
__device__ __forceinline__ u256 fq_sqrt_candidate(const u256& a) {
    u256 x2 = fq::mul(fq::sqr(a), a);
    u256 x3 = fq::mul(fq::sqr(x2), a);
    u256 x6 = fq::mul(fq::sqr_n(x3, 3), x3);
    u256 x9 = fq::mul(fq::sqr_n(x6, 3), x3);
    u256 x11 = fq::mul(fq::sqr_n(x9, 2), x2);
    u256 x22 = fq::mul(fq::sqr_n(x11, 11), x11);
    u256 x44 = fq::mul(fq::sqr_n(x22, 22), x22);
    u256 x88 = fq::mul(fq::sqr_n(x44, 44), x44);
    u256 x176 = fq::mul(fq::sqr_n(x88, 88), x88);
    u256 x220 = fq::mul(fq::sqr_n(x176, 44), x44);
    u256 x223 = fq::mul(fq::sqr_n(x220, 3), x3);
    u256 t = fq::mul(fq::sqr_n(x223, 23), x22);
    t = fq::mul(fq::sqr_n(t, 6), x2);
    return fq::sqr_n(t, 2);
}

// candidate n0 + t of getPoints: out[t] = (x, y) when x = hash(seed <> show n) has x^3 + 7 a square, else (0, 0)
__global__ void __launch_bounds__(TR_THREADS) k_hash_to_curve(const unsigned char* __restrict__ seed, int seed_len, uint64_t n0,
                                                              size_t count, int root_policy, Affine* __restrict__ out) {
    __shared__ uint32_t buf[16 * TR_THREADS];
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= count) return;
    dsha::Stream S;
    S.begin(buf);
    for (int i = 0; i < seed_len; i++) S.put_byte(seed[i]);
    S.put_uint(n0 + t);
    uint32_t d[8];
    S.finish(d);
    const u256 x = dsha::digest_to_fq(d);
    u256 seven = u256_zero();
    seven.v[0] = 7;
    const u256 rhs = fq::add(fq::mul(fq::sqr(x), x), seven);
    u256 y = fq_sqrt_candidate(rhs);
    Affine p = aff_inf();
    if (u256_eq(fq::sqr(y), rhs)) {
        const u256 ny = fq::neg(y);
        if (root_policy == TR_ROOT_EVEN && (y.v[0] & 1u)) y = ny;
        else if (root_policy == TR_ROOT_SMALLER && !u256_geq(ny, y)) y = ny;
        p.x = x; p.y = y;
    }
    st_aff(out + t, p);
}

// ---- transcript store: one record per commitment, `show x <> show y` as ASCII (app/Main.hs:78-80)
#define TR_REC_BYTES 160             // 2 * (2 + 78); records are 4-byte aligned, readers mask the tail
// rec[b][slot0 + j] = show of pts[b * pts_stride + j], j < npts; len likewise.  One thread per point.
__global__ void __launch_bounds__(TR_THREADS) k_tr_render(const Affine* __restrict__ pts, size_t pts_stride, int npts, size_t batch,
                                                          int fmt, unsigned char* __restrict__ rec, unsigned char* __restrict__ len,
                                                          size_t cap, size_t slot0) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= batch * (size_t)npts) return;
    const size_t b = t / npts, j = t % npts;
    const Affine p = ld_aff(pts + b * pts_stride + j);
    const size_t slot = b * cap + slot0 + j;
    unsigned char* o = rec + slot * TR_REC_BYTES;
    int l = 0;
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        if (fmt == TR_PREFIXED_P) { o[l++] = 'P'; o[l++] = ' '; }
        l += dsha::decimal_u256(k ? p.y : p.x, o + l);
    }
    len[slot] = (unsigned char)l;
}

// The absorb calls of a transcript so far, oldest first: call c put `npts[c]` records at slot `first[c]`.
// `oracle xs` PREPENDS xs to the commitment list (cs' = xs ++ cs, src/ZKP.hs:98), so a message walks the
// calls from the newest to the oldest, and the points of one call in the order they were given.
#define TR_MAX_CALLS 48
struct TrCalls {
    int n;
    unsigned short first[TR_MAX_CALLS], npts[TR_MAX_CALLS];
};

// Reader of one message  show i <> show (length cs') <> concat [show x <> show y | A x y <- cs']  as
// big-endian 32-bit words followed by the SHA-256 padding (0x80, zeros); the caller places the length.
struct TrReader {
    const unsigned char* rec; const unsigned char* len;        // this proof's records
    const TrCalls* calls;
    int ci, pj;                  // current call (descending), point inside it
    const unsigned char* cur; unsigned cur_len, off;
    unsigned long long pend; unsigned npend;                   // left-aligned pending bytes
    unsigned long long total;                                   // message bytes consumed
    bool src_done, pad_done;

    __device__ __forceinline__ void open_record() {
        const size_t slot = (size_t)calls->first[ci] + pj;
        cur = rec + slot * TR_REC_BYTES;
        cur_len = len[slot];
        off = 0;
    }
    __device__ __forceinline__ void begin(const unsigned char* rec_, const unsigned char* len_, const TrCalls* calls_,
                                          unsigned long long header, unsigned header_len) {
        rec = rec_; len = len_; calls = calls_;
        pend = header; npend = header_len; total = header_len;
        src_done = false; pad_done = false;
        ci = calls->n - 1; pj = 0; cur_len = 0; off = 0; cur = rec_;
        while (ci >= 0 && calls->npts[ci] == 0) ci--;
        if (ci < 0) src_done = true; else open_record();
    }
    __device__ __forceinline__ unsigned next_word() {
        while (npend < 4 && !src_done) {
            if (off >= cur_len) {                               // next record, newest call first
                if (++pj >= (int)calls->npts[ci]) {
                    pj = 0;
                    do { ci--; } while (ci >= 0 && calls->npts[ci] == 0);
                    if (ci < 0) { src_done = true; break; }
                }
                open_record();
                continue;
            }
            const unsigned k = min(4u, cur_len - off);
            unsigned w = __byte_perm(*reinterpret_cast<const unsigned*>(cur + off), 0, 0x0123);
            if (k < 4) w &= 0xffffffffu << (8 * (4 - k));
            off += 4;
            pend |= (unsigned long long)w << (32 - 8 * npend);
            npend += k;
            total += k;
        }
        if (npend < 4 && !pad_done) {                           // the source is exhausted: one 0x80, then zeros
            pend |= 0x8000000000000000ull >> (8 * npend);
            npend += 1;
            pad_done = true;
        }
        const unsigned out = (unsigned)(pend >> 32);
        pend <<= 32;
        npend = npend >= 4 ? npend - 4 : 0;
        return out;
    }
};

// challenge i (1-based, i <= 9) of proof b after the absorbs in `calls`: out[b * count + i - 1], canonical
// scalar (digest -> Fr, Encoding.hs:75-79).  One thread per (proof, challenge); the threads of a warp fill
// their blocks independently and compress in lock-step.
__global__ void __launch_bounds__(TR_THREADS) k_tr_squeeze(const unsigned char* __restrict__ rec, const unsigned char* __restrict__ len,
                                                           size_t cap, TrCalls calls, size_t batch, int count, unsigned n_coms,
                                                           u256* __restrict__ out) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool live = t < batch * (size_t)count;
    const size_t b = live ? t / count : 0;
    const unsigned i = live ? (unsigned)(t % count) + 1 : 1;
    // header: show i <> show n_coms, at most 1 + 6 digits, left-aligned in 64 bits
    unsigned long long header = (unsigned long long)('0' + i) << 56;
    unsigned hl = 1;
    {
        char buf[8];
        int n = 0;
        unsigned x = n_coms;
        do { buf[n++] = (char)('0' + x % 10); x /= 10; } while (x && n < 6);
        while (n) { header |= (unsigned long long)(unsigned char)buf[--n] << (56 - 8 * hl); hl++; }
    }
    TrReader rd;
    rd.begin(rec + b * cap * TR_REC_BYTES, len + b * cap, &calls, header, hl);
    uint32_t st[8];
    dsha::init(st);
    bool done = !live;
    unsigned long long blocks = 0;
    while (!__all_sync(0xffffffffu, done)) {
        uint32_t w[16];
        bool last = false;
        if (!done) {
#pragma unroll
            for (int k = 0; k < 16; k++) w[k] = rd.next_word();
            blocks++;
            if (rd.pad_done && rd.total + 1 + 8 <= blocks * 64) {                 // message + 0x80 + length fit: final block
                const unsigned long long bits = rd.total * 8;
                w[14] = (uint32_t)(bits >> 32);
                w[15] = (uint32_t)bits;
                last = true;
            }
            dsha::compress(st, w);
        }
        if (last) done = true;
    }
    if (live) st_u256(out + t, dsha::digest_to_fr(st));
}

// `random` (src/ZKP.hs:90-93 with h = hashToScalar rn . show, app/Main.hs:177): out[b * count + j] =
// hash(seed_b <> show (n0 + j)) as a canonical scalar.  seeds = [batch][64] bytes, seed_len <= 40 so that the
// message is a single block.
__global__ void __launch_bounds__(TR_THREADS) k_tr_random(const unsigned char* __restrict__ seeds, const unsigned char* __restrict__ seed_len,
                                                          unsigned long long n0, size_t batch, size_t count, u256* __restrict__ out) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= batch * count) return;
    const size_t b = t / count, j = t % count;
    unsigned char m[64];
#pragma unroll
    for (int k = 0; k < 64; k++) m[k] = 0;
    int l = seed_len[b];
    for (int k = 0; k < l; k++) m[k] = seeds[b * 64 + k];
    {
        char buf[20];
        int n = 0;
        unsigned long long x = n0 + j;
        do { buf[n++] = (char)('0' + (int)(x % 10)); x /= 10; } while (x);
        while (n) m[l++] = (unsigned char)buf[--n];
    }
    m[l] = 0x80;
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 16; k++) w[k] = ((uint32_t)m[4 * k] << 24) | ((uint32_t)m[4 * k + 1] << 16) | ((uint32_t)m[4 * k + 2] << 8) | m[4 * k + 3];
    w[15] = (uint32_t)l * 8;
    uint32_t st[8];
    dsha::init(st);
    dsha::compress(st, w);
    st_u256(out + t, dsha::digest_to_fr(st));
}

}  // namespace bppp
