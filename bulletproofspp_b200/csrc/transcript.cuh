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

// ---- transcript store.  `oracle xs` PREPENDS xs to the commitment list (cs' = xs ++ cs, src/ZKP.hs:98) and a
// challenge hashes  show i <> show (length cs') <> concat [show x <> show y | A x y <- cs']  (app/Main.hs:75-80):
// the newest commitments come FIRST, so no hash state can be carried from one call to the next -- every
// challenge re-hashes the whole list.  The list is kept per proof as ONE contiguous byte string, right-aligned
// in a fixed buffer of `SC` bytes: absorbing a call writes its rendering in front of what is there.  The body of
// the message after call c is buf[start_c .. SC); every earlier transcript is a suffix of the latest one, which
// is what lets the verifier render all commitments once and squeeze the challenges of all stages in one launch.
#define TR_PT_BYTES 160              // 2 * (2 + 78): the longest rendering of one point
#define TR_MAX_CALLS 80
#define TR_PREPEND_THREADS 128

// number of decimal digits of x (< 10^9)
__device__ __forceinline__ int dec_len9(uint32_t x) {
    int l = 1;
    if (x >= 100000000u) l = 9; else if (x >= 10000000u) l = 8; else if (x >= 1000000u) l = 7; else if (x >= 100000u) l = 6;
    else if (x >= 10000u) l = 5; else if (x >= 1000u) l = 4; else if (x >= 100u) l = 3; else if (x >= 10u) l = 2;
    return l;
}
// a -> base-10^9 chunks, least significant first; returns the number of chunks (0 for a = 0).  Nine full passes
// with static indices: everything stays in registers (two inlined copies of a variant with data-dependent
// loop bounds and locally indexed arrays rendered the second coordinate wrongly on sm_100a, nvcc 12.9).
__device__ __noinline__ int dec_chunks(const u256& a, uint32_t chunks[9]) {
    uint32_t t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = a.v[i];
#pragma unroll
    for (int c = 0; c < 9; c++) {
        uint64_t rem = 0;
#pragma unroll
        for (int i = 7; i >= 0; i--) {
            const uint64_t cur = (rem << 32) | t[i];
            t[i] = (uint32_t)(cur / 1000000000u);
            rem = cur % 1000000000u;
        }
        chunks[c] = (uint32_t)rem;
    }
    int nc = 0;
#pragma unroll
    for (int c = 0; c < 9; c++)
        if (chunks[c]) nc = c + 1;
    return nc;
}
__device__ __forceinline__ int dec_chunks_len(const uint32_t chunks[9], int nc) {
    uint32_t topc = 0;
#pragma unroll
    for (int c = 0; c < 9; c++)
        if (c == nc - 1) topc = chunks[c];
    return nc ? 9 * (nc - 1) + dec_len9(topc) : 1;
}
// the digits of the chunk form, most significant first, no leading zeros ("0" for zero); returns the length
__device__ __noinline__ int dec_chunks_write(const uint32_t chunks[9], int nc, unsigned char* out) {
    if (nc == 0) { out[0] = '0'; return 1; }
    int len = 0;
#pragma unroll 1
    for (int c = nc - 1; c >= 0; c--) {
        uint32_t x = chunks[c];
        const int nd = c == nc - 1 ? dec_len9(x) : 9;
#pragma unroll 1
        for (int k = nd - 1; k >= 0; k--) { out[len + k] = (unsigned char)('0' + x % 10); x /= 10; }
        len += nd;
    }
    return len;
}

// One absorb call for every proof of the batch: CTA b renders the `npts` points pts[b * pts_stride + j] (in the
// order given: they become the FRONT of the list) in front of proof b's transcript.  start = [B][TR_MAX_CALLS + 1]
// body offsets; state `call` is read, state `call + 1` written.
__global__ void __launch_bounds__(TR_PREPEND_THREADS) k_tr_prepend(const Affine* __restrict__ pts, size_t pts_stride, int npts, int fmt,
                                                                   unsigned char* __restrict__ buf, unsigned SC,
                                                                   unsigned* __restrict__ start, int call) {
    __shared__ unsigned s_scan[TR_PREPEND_THREADS];
    __shared__ unsigned s_total;
    const size_t b = blockIdx.x;
    const int tid = threadIdx.x;
    const unsigned prev = start[b * (TR_MAX_CALLS + 1) + call];
    // pass 1: total length of the call
    unsigned mine = 0;
    for (int j = tid; j < npts; j += TR_PREPEND_THREADS) {
        const Affine p = ld_aff(pts + b * pts_stride + j);
        uint32_t ch[9];
        int nc = dec_chunks(p.x, ch);
        mine += dec_chunks_len(ch, nc);
        nc = dec_chunks(p.y, ch);
        mine += dec_chunks_len(ch, nc) + (fmt == TR_PREFIXED_P ? 4 : 0);
    }
    s_scan[tid] = mine;
    __syncthreads();
    if (tid == 0) {
        unsigned t = 0;
        for (int k = 0; k < TR_PREPEND_THREADS; k++) t += s_scan[k];
        s_total = t;
    }
    __syncthreads();
    const unsigned new_start = prev - s_total;
    unsigned char* base = buf + b * (size_t)SC + new_start;
    // pass 2: chunks of TR_PREPEND_THREADS consecutive points, offsets by a block scan with a running carry
    unsigned carry = 0;
    for (int j0 = 0; j0 < npts; j0 += TR_PREPEND_THREADS) {
        const int j = j0 + tid;
        uint32_t cx[9], cy[9];
        int nx = 0, ny = 0;
        unsigned len = 0;
        if (j < npts) {
            const Affine p = ld_aff(pts + b * pts_stride + j);
            nx = dec_chunks(p.x, cx);
            ny = dec_chunks(p.y, cy);
            len = dec_chunks_len(cx, nx) + dec_chunks_len(cy, ny) + (fmt == TR_PREFIXED_P ? 4 : 0);
        }
        __syncthreads();
        s_scan[tid] = len;
        __syncthreads();
        for (int d = 1; d < TR_PREPEND_THREADS; d <<= 1) {          // inclusive Hillis-Steele scan
            const unsigned v = tid >= d ? s_scan[tid - d] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const unsigned off = carry + s_scan[tid] - len;
        carry += s_scan[TR_PREPEND_THREADS - 1];
        if (j < npts) {
            unsigned char* o = base + off;
            int l = 0;
            if (fmt == TR_PREFIXED_P) { o[l++] = 'P'; o[l++] = ' '; }
            l += dec_chunks_write(cx, nx, o + l);
            if (fmt == TR_PREFIXED_P) { o[l++] = 'P'; o[l++] = ' '; }
            l += dec_chunks_write(cy, ny, o + l);
        }
    }
    if (tid == 0) start[b * (TR_MAX_CALLS + 1) + call + 1] = new_start;
}

// Which challenges one squeeze produces: challenge j is scalar `idx[j]` (1-based, <= 9) of the transcript after
// `state[j]` absorb calls, which then holds `ncoms[j]` commitments.
#define TR_MAX_CHAL 48
struct TrPlan {
    int count;
    unsigned char idx[TR_MAX_CHAL], state[TR_MAX_CHAL];
    unsigned ncoms[TR_MAX_CHAL];
};

// out[b * count + j] = challenge j of proof b, canonical scalar (digest -> Fr, Encoding.hs:75-79).  A block that lies
// inside the body is read with 17 aligned word loads and byte permutes (the body's alignment is fixed for the whole
// message); the first block (header) and the last one or two (tail, 0x80, bit length) are assembled byte by byte.
// Hashing is a serial chain per message and a warp that runs it alone is bound by its own issue rate (~1.7 cycles per
// instruction), so both kernels split a block's work over two warps: message schedule ahead, rounds behind.
//   k_tr_squeeze_coop   launches of <= TR_COOP_MAX hashes: a CTA per hash (one lane runs the rounds)
//   k_tr_squeeze_pair   larger launches: 32 hashes per CTA, lane = hash in both warps
// (Round 2's first version, one thread per hash doing both halves, took 1.85x longer per launch.)

// A FEW hashes (one proof alone: the drop-in seams' batch size).  A thread that hashes a message by itself issues
// ~1900 instructions per 64-byte block, half of them message assembly and schedule; a 22 KB transcript took 0.6 ms
// that way, and a proof has 14 of them in a row.  Here a CTA of two warps owns one hash:
// warp 1 assembles and expands the blocks of the next tile (one block per lane, kw = K + W into shared memory),
// lane 0 of warp 0 runs the rounds of the current tile.
#define TRC_TILE 32                    // blocks per tile (one per lane of the scheduling warp)
#define TRC_STRIDE 68                  // words per block in shared memory: 64 + 4 (16-byte rows, 4-way store conflicts)
__global__ void __launch_bounds__(64) k_tr_squeeze_coop(const unsigned char* __restrict__ buf, unsigned SC, const unsigned* __restrict__ start,
                                                        TrPlan plan, size_t batch, u256* __restrict__ out) {
    __shared__ __align__(16) uint32_t kw[2][TRC_TILE * TRC_STRIDE];
    const size_t t = blockIdx.x;
    const size_t b = t / plan.count;
    const int j = (int)(t % plan.count);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char hdr[8];
    unsigned hl = 0;
    hdr[hl++] = (unsigned char)('0' + plan.idx[j]);
    {
        char tmp[8];
        int n = 0;
        unsigned x = plan.ncoms[j];
        do { tmp[n++] = (char)('0' + x % 10); x /= 10; } while (x && n < 7);
        while (n) hdr[hl++] = (unsigned char)tmp[--n];
    }
    const unsigned st0 = start[b * (TR_MAX_CALLS + 1) + plan.state[j]];
    const unsigned char* body = buf + b * (size_t)SC + st0;
    const unsigned L = SC - st0;
    const unsigned long long T = (unsigned long long)hl + L;
    const unsigned nblk = (unsigned)((T + 9 + 63) / 64);
    const unsigned ntile = (nblk + TRC_TILE - 1) / TRC_TILE;
    const size_t a0 = (size_t)(body - hl);
    const unsigned sh = (unsigned)(a0 & 3);
    const unsigned sel = (sh + 3) | ((sh + 2) << 4) | ((sh + 1) << 8) | (sh << 12);
    auto schedule = [&](unsigned tile) {
        const unsigned k = tile * TRC_TILE + lane;
        if (k >= nblk) return;
        uint32_t w[16];
        const unsigned long long m0 = (unsigned long long)k * 64;
        if (k >= 1 && m0 + 64 <= T) {
            const uint32_t* q = reinterpret_cast<const uint32_t*>((a0 + m0) & ~(size_t)3);
            uint32_t nx[17];
#pragma unroll
            for (int i = 0; i < 17; i++) nx[i] = q[i];
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = __byte_perm(nx[i], nx[i + 1], sel);
        } else {
#pragma unroll 1
            for (int i = 0; i < 16; i++) {
                uint32_t x = 0;
#pragma unroll 1
                for (int c = 0; c < 4; c++) {
                    const unsigned long long m = m0 + 4 * i + c;
                    unsigned byte = 0;
                    if (m < hl) byte = hdr[m];
                    else if (m < T) byte = body[m - hl];
                    else if (m == T) byte = 0x80;
                    x = (x << 8) | byte;
                }
                w[i] = x;
            }
            if (k == nblk - 1) {
                const unsigned long long bits = T * 8;
                w[14] = (uint32_t)(bits >> 32);
                w[15] = (uint32_t)bits;
            }
        }
        dsha::expand_kw(w, &kw[tile & 1][lane * TRC_STRIDE]);
    };
    uint32_t st[8];
    dsha::init(st);
    if (warp == 1) schedule(0);
    __syncthreads();
    for (unsigned tile = 0; tile < ntile; tile++) {
        if (warp == 1) {
            if (tile + 1 < ntile) schedule(tile + 1);
        } else if (lane == 0) {
            const unsigned nb = min((unsigned)TRC_TILE, nblk - tile * TRC_TILE);
#pragma unroll 1
            for (unsigned i = 0; i < nb; i++) dsha::rounds_kw(st, &kw[tile & 1][i * TRC_STRIDE]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) st_u256(out + t, dsha::digest_to_fr(st));
}

// The batch form of the same split (launches of more than TR_COOP_MAX hashes): 32 hashes per CTA, lane = hash in both
// warps.  Warp 1 assembles and expands block k + 1 of its 32 messages (the next block's words already in flight) while
// warp 0 runs the rounds of block k; K + W goes through shared memory as [4 rounds][lane] 16-byte cells (conflict-free
// either way).  The instruction count of one thread doing both halves, on two warps instead of one: a launch of 512 hashes is 16 warps
// on 148 SMs, so its duration is one warp's issue time -- 1.85x less this way.
__global__ void __launch_bounds__(64) k_tr_squeeze_pair(const unsigned char* __restrict__ buf, unsigned SC, const unsigned* __restrict__ start,
                                                        TrPlan plan, size_t batch, u256* __restrict__ out) {
    __shared__ uint4 kw[2][16][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t n_hash = batch * (size_t)plan.count;
    const size_t t = blockIdx.x * (size_t)32 + lane;
    const bool live = t < n_hash;
    const size_t b = live ? t / plan.count : 0;
    const int j = live ? (int)(t % plan.count) : 0;
    unsigned char hdr[8];
    unsigned hl = 0;
    hdr[hl++] = (unsigned char)('0' + plan.idx[j]);
    {
        char tmp[8];
        int n = 0;
        unsigned x = plan.ncoms[j];
        do { tmp[n++] = (char)('0' + x % 10); x /= 10; } while (x && n < 7);
        while (n) hdr[hl++] = (unsigned char)tmp[--n];
    }
    const unsigned st0 = start[b * (TR_MAX_CALLS + 1) + plan.state[j]];
    const unsigned char* body = buf + b * (size_t)SC + st0;
    const unsigned L = SC - st0;
    const unsigned long long T = (unsigned long long)hl + L;
    const unsigned nblk = live ? (unsigned)((T + 9 + 63) / 64) : 0;
    const unsigned nblk_max = __reduce_max_sync(0xffffffffu, nblk);            // the same in both warps (lane = hash)
    const size_t a0 = (size_t)(body - hl);
    const unsigned sh = (unsigned)(a0 & 3);
    const unsigned sel = (sh + 3) | ((sh + 2) << 4) | ((sh + 1) << 8) | (sh << 12);
    uint32_t nx[17];
    bool nx_valid = false;
    auto interior = [&](unsigned k) { return k >= 1 && (unsigned long long)k * 64 + 64 <= T; };
    auto fetch = [&](unsigned k) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>((a0 + (unsigned long long)k * 64) & ~(size_t)3);
#pragma unroll
        for (int i = 0; i < 17; i++) nx[i] = q[i];
    };
    auto schedule = [&](unsigned k) {                                          // block k of this lane's message -> kw[k & 1]
        if (k >= nblk) return;
        uint32_t w[16];
        const unsigned long long m0 = (unsigned long long)k * 64;
        if (interior(k)) {
            if (!nx_valid) fetch(k);
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = __byte_perm(nx[i], nx[i + 1], sel);
            nx_valid = false;
            if (interior(k + 1)) { fetch(k + 1); nx_valid = true; }
        } else {
#pragma unroll 1
            for (int i = 0; i < 16; i++) {
                uint32_t x = 0;
#pragma unroll 1
                for (int c = 0; c < 4; c++) {
                    const unsigned long long m = m0 + 4 * i + c;
                    unsigned byte = 0;
                    if (m < hl) byte = hdr[m];
                    else if (m < T) byte = body[m - hl];
                    else if (m == T) byte = 0x80;
                    x = (x << 8) | byte;
                }
                w[i] = x;
            }
            if (k == nblk - 1) {
                const unsigned long long bits = T * 8;
                w[14] = (uint32_t)(bits >> 32);
                w[15] = (uint32_t)bits;
            }
            if (interior(k + 1)) { fetch(k + 1); nx_valid = true; }
        }
        dsha::expand_kw_cells(w, &kw[k & 1][0][lane]);
    };
    uint32_t st[8];
    dsha::init(st);
    if (warp == 1) schedule(0);
    __syncthreads();
    for (unsigned k = 0; k < nblk_max; k++) {
        if (warp == 1) schedule(k + 1);
        else if (k < nblk) dsha::rounds_kw_cells(st, &kw[k & 1][0][lane]);
        __syncthreads();
    }
    if (warp == 0 && live) st_u256(out + t, dsha::digest_to_fr(st));
}

// `random` (src/ZKP.hs:90-93 with h = hashToScalar rn . show, app/Main.hs:177): out[b * out_stride + j] =
// hash(seed_b <> show (n0_b + j)) as a canonical scalar, n0_b = n0s[b] (or n0 for every proof when n0s is
// null).  seeds = [batch][64] bytes, seed_len <= 40 so that the message is a single block.
__global__ void __launch_bounds__(TR_THREADS) k_tr_random(const unsigned char* __restrict__ seeds, const unsigned char* __restrict__ seed_len,
                                                          unsigned long long n0, const unsigned long long* __restrict__ n0s, size_t batch,
                                                          size_t count, u256* __restrict__ out, size_t out_stride) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= batch * count) return;
    const size_t b = t / count, j = t % count;
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 16; k++) w[k] = 0;
    int l = seed_len[b];
    auto put = [&](int pos, unsigned byte) { w[pos >> 2] |= byte << (24 - 8 * (pos & 3)); };
    for (int k = 0; k < l; k++) put(k, seeds[b * 64 + k]);
    {
        char buf[20];
        int n = 0;
        unsigned long long x = (n0s ? n0s[b] : n0) + j;
        do { buf[n++] = (char)('0' + (int)(x % 10)); x /= 10; } while (x);
        while (n) put(l++, (unsigned char)buf[--n]);
    }
    put(l, 0x80);
    w[15] = (uint32_t)l * 8;
    uint32_t st[8];
    dsha::init(st);
    dsha::compress(st, w);
    st_u256(out + b * out_stride + j, dsha::digest_to_fr(st));
}

}  // namespace bppp
