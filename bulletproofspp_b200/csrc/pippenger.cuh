// Size-aware Pippenger MSM over points in global memory (sm_100a) -- SURVEY K4.
//
// Replaces `commit = innerProduct . openToList` (src/Commitment.hs:416-417, 325-335) for LONG openings
// whose bases are not the resident fixed-base table: every round of one large norm argument (the
// folded generators are per proof), the verifier's collapsed MSM at large N (Bulletproof.hs:362-378),
// and bppp_msm.  The batched small-proof path keeps its shared-memory kernels (kernels.cuh).
//
// Every scalar is first split with the secp256k1 endomorphism phi(x, y) = (beta x, y) = lambda (x, y)
// (GLV): k = k1 + k2 lambda (mod r) with |k1|, |k2| < 2^128, so an n-term MSM becomes 2n terms of 128
// bits over the points P_i and phi(P_i) (one extra Fq multiplication when the entry is used).  The
// number of additions is unchanged, but the windows -- and with them the bucket reductions and the
// final Horner, the one inherently serial part (c doublings per window) -- are halved.
// The window width follows the length, c = clamp(floor(log2 2n) - 4, 5, 15), so a window has about
// 2n/32 buckets of ~32 entries: W = ceil(129/c) windows of signed digits (|d| <= 2^(c-1)), about
// W * 2n * 1.1 point additions in total against 43 * n for the fixed 6-bit kernel.
// Skewed digit distributions (small or repeated scalars, the short top window) are handled, not assumed
// away: the histogram / scatter atomics are warp-aggregated (__match_any_sync), the accumulation is
// balanced by construction, and a bucket cut into many pieces is merged by a whole CTA.
//
//   k_pip_glv       k -> (|k1|, |k2|, signs): two 256x256 products with the precomputed round(2^384 b/r),
//                   two 128x128 products, one multiplication by lambda mod r
//   k_pip_count     digit histogram: one global atomic per non-zero (term, window) digit
//   k_pip_scan_*    exclusive scan of the histogram -> bucket offsets (three small kernels)
//   k_pip_scatter   (term | sign) entries written bucket by bucket: a counting sort by (problem, window, bucket)
//   k_pip_accum     the sorted entry list is cut into EQUAL ranges, one per thread (balanced whatever the
//                   digit distribution); a thread adds its entries with XYZZ mixed additions (8M + 2S) and
//                   flushes one sum per bucket run: complete runs go straight to the bucket array, the
//                   first / last (possibly shared) runs of a range to per-thread slots
//   k_pip_merge     one thread per bucket adds the pieces of a bucket that spans several ranges; buckets with
//                   more than PIP_HEAVY pieces are queued for k_pip_merge_heavy (one CTA per bucket, tree)
//   k_pip_reduce1   one thread per segment of SEG buckets: S = sum B, T = sum (k+1) B by running sums
//   k_pip_reduce2   one warp per (problem, window): sum_b (b+1) B_b from the segment sums (shuffle
//                   suffix scan + tree), converted to Jacobian
//   k_pip_horner    one thread per problem: sum_j 2^(c j) W_j  (the only inherently serial part:
//                   ~256 doublings)
// A "problem" is one (proof, output) MSM; all kernels run every problem of a call in one launch.
// Group-law associativity makes the result independent of the (atomic) entry order: bit-exact.
#pragma once
#include "kernels.cuh"

namespace bppp {

struct PipArgs {
    const Affine* pts; size_t pts_stride;        // bases: pts[p*pts_stride + i]  (stride 0: shared by all proofs)
    const u256* sc; size_t sc_stride, sc_out_stride;   // canonical scalars sc[p*sc_stride + o*sc_out_stride + i]
    u256* dec;                                   // [n_prob * n] GLV halves: limbs 0..3 = |k1|, 4..7 = |k2|
    unsigned char* dsgn;                         // [n_prob * n] bit 0: k1 < 0, bit 1: k2 < 0
    unsigned n;                                  // terms per problem (scalars; 2n half-length terms after the split)
    unsigned n_out, n_prob;                      // outputs per proof; problems = proofs * n_out (prob = p*n_out + o)
    int c, W, NBK;                               // window bits, windows, buckets per window (2^(c-1))
    unsigned NT;                                 // n_prob * W * NBK buckets in total
    unsigned* off;                               // [NT + 1] histogram, then exclusive offsets
    unsigned* cur;                               // [NT] scatter cursors
    unsigned* ent;                               // [<= n_prob * 2n * W] entries: term index | phi << 30 | sign << 31
    unsigned* bsum;                              // scan scratch: one total per block of PIP_SCAN_BLOCK counters
    Xyzz* bucket;                                // [NT] bucket sums
    Xyzz* slotF; Xyzz* slotL;                    // [threads of k_pip_accum] first / last run of each range
    unsigned* heavy;                             // [0] = count, [1..] = buckets with more than PIP_HEAVY pieces
    int L;                                       // entries per thread in k_pip_accum
    int SEG, NS;                                 // buckets per segment, segments per window
    Xyzz* segS; Xyzz* segT;                      // [n_prob * W * NS]
    Jac* win;                                    // [n_prob * W] window sums
    Jac* out;                                    // [n_prob]
};

// ---- GLV split.  lambda^3 = 1 (mod r), beta^3 = 1 (mod q), lambda (x, y) = (beta x, y); lattice basis
// (a1, b1), (a2, b2) with a1 + b1 lambda = a2 + b2 lambda = 0 (mod r), b2 = a1.  With g1 = round(2^384 b2 / r),
// g2 = round(2^384 (-b1) / r):  c1 = round(k g1 / 2^384), c2 = round(k g2 / 2^384),
// k2 = c1 (-b1) - c2 b2,  k1 = k - k2 lambda (mod r), centred.  |k1|, |k2| < 2^128 (checked exhaustively on
// edge values and 2*10^5 random scalars against big-integer arithmetic by tests/test_host.py).
__device__ __forceinline__ u256 glv_const(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6, uint32_t a7) {
    u256 r;
    r.v[0] = a0; r.v[1] = a1; r.v[2] = a2; r.v[3] = a3; r.v[4] = a4; r.v[5] = a5; r.v[6] = a6; r.v[7] = a7;
    return r;
}
#define GLV_G1 glv_const(0x45DBB031u, 0xE893209Au, 0x71E8CA7Fu, 0x3DAA8A14u, 0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u)
#define GLV_G2 glv_const(0x8AC47F71u, 0x1571B4AEu, 0x9DF506C6u, 0x221208ACu, 0x0ABFE4C4u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u)
#define GLV_MB1 glv_const(0x0ABFE4C3u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u, 0, 0, 0, 0)
#define GLV_B2 glv_const(0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u, 0, 0, 0, 0)
#define GLV_LAMBDA glv_const(0x1B23BD72u, 0xDF02967Cu, 0x20816678u, 0x122E22EAu, 0x8812645Au, 0xA5261C02u, 0xC05C30E0u, 0x5363AD4Cu)
#define GLV_BETA glv_const(0x719501EEu, 0xC1396C28u, 0x12F58995u, 0x9CF04975u, 0xAC3434E9u, 0x6E64479Eu, 0x657C0710u, 0x7AE96A2Bu)
// round(k * g / 2^384): a 128-bit value in the low limbs
__device__ __forceinline__ u256 glv_mul_shift384(const u256& k, const u256& g) {
    uint32_t t[16];
    mul_wide(t, k, g);
    u256 c = u256_zero(), one = u256_zero(), r;
    c.v[0] = t[12]; c.v[1] = t[13]; c.v[2] = t[14]; c.v[3] = t[15];
    one.v[0] = t[11] >> 31;
    u256_add(r, c, one);
    return r;
}
__device__ __forceinline__ u256 glv_mul_lo(const u256& a, const u256& b) {      // a, b < 2^128: the 256-bit product
    uint32_t t[16];
    mul_wide(t, a, b);
    u256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    return r;
}
__global__ void __launch_bounds__(256) k_pip_glv(PipArgs A) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)A.n_prob * A.n) return;
    const unsigned prob = (unsigned)(t / A.n), i = (unsigned)(t % A.n);
    const unsigned p = prob / A.n_out, o = prob % A.n_out;
    const u256 k = ld_u256(A.sc + (size_t)p * A.sc_stride + (size_t)o * A.sc_out_stride + i);
    u256 out = u256_zero();
    unsigned sg = 0;
    if (!u256_is_zero(k)) {
        const u256 c1 = glv_mul_shift384(k, GLV_G1), c2 = glv_mul_shift384(k, GLV_G2);
        const u256 t1 = glv_mul_lo(c1, GLV_MB1), t2 = glv_mul_lo(c2, GLV_B2);
        u256 m2;
        const bool neg2 = u256_sub(m2, t1, t2) != 0;                  // k2 = t1 - t2
        if (neg2) u256_sub(m2, t2, t1);
        const u256 prod = fr::mul(fr::to_mont(m2), GLV_LAMBDA);       // |k2| lambda mod r (canonical)
        u256 k1 = neg2 ? fr::add(k, prod) : fr::sub(k, prod);         // k - k2 lambda
        u256 half = fr::modulus();                                    // (r - 1) / 2
#pragma unroll
        for (int j = 0; j < 8; j++) half.v[j] = (half.v[j] >> 1) | (j < 7 ? half.v[j + 1] << 31 : 0);
        const bool neg1 = !u256_geq(half, k1);
        if (neg1) u256_sub(k1, fr::modulus(), k1);
#pragma unroll
        for (int j = 0; j < 4; j++) { out.v[j] = k1.v[j]; out.v[4 + j] = m2.v[j]; }
        sg = (neg1 ? 1u : 0u) | (neg2 ? 2u : 0u);
    }
    st_u256(A.dec + t, out);
    A.dsgn[t] = (unsigned char)sg;
}
// signed window digit j (width c <= 16) of a 128-bit magnitude m[0..3]
__device__ __forceinline__ int pip_digit128(const uint32_t m[4], int j, int c, int& carry) {
    const int bit = j * c;
    uint32_t w = 0;
    if (bit < 128) {
        const int limb = bit >> 5, sh = bit & 31;
        uint64_t two = m[limb];
        if (limb + 1 < 4) two |= (uint64_t)m[limb + 1] << 32;
        w = (uint32_t)(two >> sh) & ((1u << c) - 1u);
    }
    int d = (int)w + carry;
    if (d > (1 << (c - 1))) { d -= (1 << c); carry = 1; } else carry = 0;
    return d;
}
__global__ void __launch_bounds__(256) k_pip_count(PipArgs A) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool valid = t < (size_t)A.n_prob * A.n;
    const unsigned prob = valid ? (unsigned)(t / A.n) : 0;
    u256 s = u256_zero();
    if (valid) s = ld_u256(A.dec + t);
    const unsigned lane = threadIdx.x & 31;
#pragma unroll 1
    for (int h = 0; h < 2; h++) {                                 // k1 on P, k2 on phi(P)
        int carry = 0;
#pragma unroll 1
        for (int j = 0; j < A.W; j++) {                           // all 32 lanes stay in the loop (warp-wide match)
            const int d = pip_digit128(s.v + 4 * h, j, A.c, carry);
            const unsigned key = d ? (prob * (unsigned)A.W + j) * (unsigned)A.NBK + (d < 0 ? -d : d) - 1 : 0xffffffffu;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            if (d && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(A.off + key, (unsigned)__popc(peers));
        }
    }
}

// ---- exclusive scan of off[0..NT) in place, off[NT] = total, cur = copy of the offsets
#define PIP_SCAN_BLOCK 2048          // counters per CTA (256 threads x 8)
__device__ __forceinline__ unsigned pip_block_scan(unsigned v, unsigned* sm, unsigned& total) {   // exclusive, 256 threads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        unsigned o = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= s) inc += o;
    }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned w = lane < 8 ? sm[lane] : 0, wi = w;
#pragma unroll
        for (int s = 1; s < 8; s <<= 1) {
            unsigned o = __shfl_up_sync(0xffffffffu, wi, s);
            if (lane >= s) wi += o;
        }
        if (lane < 8) sm[8 + lane] = wi - w;          // exclusive warp offsets
        if (lane == 7) sm[16] = wi;
    }
    __syncthreads();
    total = sm[16];
    return inc - v + sm[8 + warp];
}
__global__ void __launch_bounds__(256) k_pip_scan_blocks(PipArgs A) {
    __shared__ unsigned sm[17];
    const size_t base = (size_t)blockIdx.x * PIP_SCAN_BLOCK + threadIdx.x * 8;
    unsigned v = 0;
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (base + k < A.NT) v += A.off[base + k];
    unsigned total;
    pip_block_scan(v, sm, total);
    if (threadIdx.x == 0) A.bsum[blockIdx.x] = total;
}
// one CTA: exclusive scan of the block totals
__global__ void __launch_bounds__(256) k_pip_scan_top(PipArgs A, unsigned n_blocks) {
    __shared__ unsigned sm[17];
    __shared__ unsigned carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (unsigned b0 = 0; b0 < n_blocks; b0 += 256) {
        const unsigned i = b0 + threadIdx.x;
        const unsigned v = i < n_blocks ? A.bsum[i] : 0;
        unsigned total;
        const unsigned ex = pip_block_scan(v, sm, total);
        const unsigned carry = carry_s;
        if (i < n_blocks) A.bsum[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) A.off[A.NT] = carry_s;
}
__global__ void __launch_bounds__(256) k_pip_scan_apply(PipArgs A) {
    __shared__ unsigned sm[17];
    const size_t base = (size_t)blockIdx.x * PIP_SCAN_BLOCK + threadIdx.x * 8;
    unsigned c[8], v = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        c[k] = base + k < A.NT ? A.off[base + k] : 0;
        v += c[k];
    }
    unsigned total;
    unsigned run = pip_block_scan(v, sm, total) + A.bsum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (base + k < A.NT) {
            A.off[base + k] = run;
            A.cur[base + k] = run;
            run += c[k];
        }
}

__global__ void __launch_bounds__(256) k_pip_scatter(PipArgs A) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool valid = t < (size_t)A.n_prob * A.n;
    const unsigned prob = valid ? (unsigned)(t / A.n) : 0, i = valid ? (unsigned)(t % A.n) : 0;
    u256 s = u256_zero();
    unsigned sg = 0;
    if (valid) { s = ld_u256(A.dec + t); sg = A.dsgn[t]; }
    const unsigned lane = threadIdx.x & 31;
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
        const bool neg = (sg >> h) & 1u;
        int carry = 0;
#pragma unroll 1
        for (int j = 0; j < A.W; j++) {
            const int d = pip_digit128(s.v + 4 * h, j, A.c, carry);
            const unsigned key = d ? (prob * (unsigned)A.W + j) * (unsigned)A.NBK + (d < 0 ? -d : d) - 1 : 0xffffffffu;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            if (d) {
                const int leader = __ffs(peers) - 1;
                unsigned base = 0;
                if (lane == (unsigned)leader) base = atomicAdd(A.cur + key, (unsigned)__popc(peers));
                base = __shfl_sync(peers, base, leader);
                const unsigned pos = base + __popc(peers & ((1u << lane) - 1u));
                A.ent[pos] = i | ((unsigned)h << 30) | (((d < 0) != neg) ? 0x80000000u : 0u);
            }
        }
    }
}

// bases of the problem that owns global bucket g
__device__ __forceinline__ const Affine* pip_bases(const PipArgs& A, unsigned g) {
    const unsigned prob = g / (unsigned)(A.W * A.NBK);
    return A.pts + (size_t)(prob / A.n_out) * A.pts_stride;
}

__global__ void __launch_bounds__(128, 5) k_pip_accum(PipArgs A) {
    const unsigned E = A.off[A.NT];
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t a64 = t * (size_t)A.L;
    if (a64 >= E) return;
    const unsigned a = (unsigned)a64, b = (unsigned)min((size_t)E, a64 + A.L);
    const unsigned* __restrict__ off = A.off;
    unsigned lo = 0, hi = A.NT - 1;                       // last bucket with off[g] <= a
    while (lo < hi) {
        const unsigned mid = (lo + hi + 1) >> 1;
        if (off[mid] <= a) lo = mid; else hi = mid - 1;
    }
    unsigned g = lo;
    unsigned run_end = min(off[g + 1], b);
    const Affine* pts = pip_bases(A, g);
    const u256 beta = GLV_BETA;
    Xyzz acc = xyzz_inf();
    for (unsigned pos = a; pos < b; pos++) {
        if (pos == run_end) {                                 // flush the finished run (it cannot be the last one)
            const bool first = off[g] <= a;
            st_xyzz(first ? (A.slotF + t) : (A.bucket + g), acc);
            acc = xyzz_inf();
            do { g++; } while (off[g + 1] <= pos);
            run_end = min(off[g + 1], b);
            pts = pip_bases(A, g);
        }
        const unsigned e = A.ent[pos];
        if (pos + 1 < b)                                          // the gather of the next base overlaps this addition
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pts + (A.ent[pos + 1] & 0x3fffffffu)));
        Affine P = ld_aff(pts + (e & 0x3fffffffu));
        if (e & 0x40000000u) P.x = fq::mul(P.x, beta);            // phi(P) = (beta x, y) = lambda P
        if (e >> 31) P.y = fq::neg(P.y);
        acc = xyzz_madd(acc, P);
    }
    const bool first = off[g] <= a, last = off[g + 1] >= b;
    // the final run: a range that starts inside it owns slotF; else if the bucket continues past b, slotL;
    // else the run is the complete bucket
    st_xyzz(first ? (A.slotF + t) : (last && off[g + 1] > b ? (A.slotL + t) : (A.bucket + g)), acc);
}

// bucket g = sum of its pieces.  A bucket [k0, k1) that lies strictly inside one thread's range was
// written to bucket[g] by k_pip_accum itself; every other bucket is assembled here -- by this thread
// when it has at most PIP_HEAVY pieces, else by a whole CTA of k_pip_merge_heavy.
#define PIP_HEAVY 12
__device__ __forceinline__ Xyzz pip_piece(const PipArgs& A, unsigned t, unsigned k0, unsigned k1, unsigned E) {
    const unsigned a = t * (unsigned)A.L, b = min(E, a + (unsigned)A.L);
    if (k0 <= a) return ld_xyzz(A.slotF + t);
    if (k1 > b) return ld_xyzz(A.slotL + t);
    return xyzz_inf();                                        // unreachable for a bucket that spans ranges
}
__global__ void __launch_bounds__(128) k_pip_merge(PipArgs A) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A.NT) return;
    const unsigned E = A.off[A.NT];
    const unsigned k0 = A.off[g], k1 = A.off[g + 1];
    if (k0 == k1) { st_xyzz(A.bucket + g, xyzz_inf()); return; }
    const unsigned L = (unsigned)A.L;
    const unsigned t0 = k0 / L, t1 = (k1 - 1) / L;
    if (t0 == t1 && k0 > t0 * L && k1 <= min(E, t0 * L + L)) return;   // complete run: already in bucket[g]
    if (t1 - t0 + 1 > PIP_HEAVY) {
        A.heavy[1 + atomicAdd(A.heavy, 1u)] = g;
        return;
    }
    Xyzz sum = xyzz_inf();
    for (unsigned t = t0; t <= t1; t++) sum = xyzz_add(sum, pip_piece(A, t, k0, k1, E));
    st_xyzz(A.bucket + g, sum);
}
// one CTA per queued bucket: threads add pieces t0 + tid, t0 + tid + 256, ..., then a shared-memory tree
#define PIP_HEAVY_THREADS 256
__global__ void __launch_bounds__(PIP_HEAVY_THREADS) k_pip_merge_heavy(PipArgs A) {
    __shared__ Xyzz sm[PIP_HEAVY_THREADS / 2];
    const unsigned n_heavy = A.heavy[0], E = A.off[A.NT], L = (unsigned)A.L;
    for (unsigned h = blockIdx.x; h < n_heavy; h += gridDim.x) {
        const unsigned g = A.heavy[1 + h];
        const unsigned k0 = A.off[g], k1 = A.off[g + 1];
        const unsigned t0 = k0 / L, t1 = (k1 - 1) / L;
        Xyzz sum = xyzz_inf();
        for (unsigned t = t0 + threadIdx.x; t <= t1; t += PIP_HEAVY_THREADS) sum = xyzz_add(sum, pip_piece(A, t, k0, k1, E));
        for (unsigned half = PIP_HEAVY_THREADS / 2; half >= 1; half >>= 1) {
            __syncthreads();
            if (threadIdx.x >= half && threadIdx.x < 2 * half) sm[threadIdx.x - half] = sum;
            __syncthreads();
            if (threadIdx.x < half) sum = xyzz_add(sum, sm[threadIdx.x]);
        }
        if (threadIdx.x == 0) st_xyzz(A.bucket + g, sum);
        __syncthreads();
    }
}

// one thread per segment of SEG consecutive buckets of one (problem, window)
__global__ void __launch_bounds__(128) k_pip_reduce1(PipArgs A) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned n_seg = A.n_prob * (unsigned)A.W * (unsigned)A.NS;
    if (s >= n_seg) return;
    const Xyzz* B = A.bucket + (size_t)s * A.SEG;             // segments tile the bucket array in order
    Xyzz run = xyzz_inf(), T = xyzz_inf();
#pragma unroll 1
    for (int k = A.SEG - 1; k >= 0; k--) {
        run = xyzz_add_t<FqInl>(run, ld_xyzz(B + k));
        T = xyzz_add_t<FqInl>(T, run);                                 // sum (k+1) B_k
    }
    st_xyzz(A.segS + s, run);
    st_xyzz(A.segT + s, T);
}

__device__ __forceinline__ Xyzz shfl_xyzz(const Xyzz& a, int src) {
    Xyzz r;
    r.X = shfl_u256(a.X, src); r.Y = shfl_u256(a.Y, src); r.ZZ = shfl_u256(a.ZZ, src); r.ZZZ = shfl_u256(a.ZZZ, src);
    return r;
}
// one warp per (problem, window):  sum_b (b+1) B_b = sum_s T_s + SEG * sum_s s S_s ; lane l owns `per`
// consecutive segments:  sum_s s S_s = sum_l Bw_l + per * sum_{l>=1} suffix_l(A)
__global__ void __launch_bounds__(128) k_pip_reduce2(PipArgs A) {
    const unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= A.n_prob * (unsigned)A.W) return;
    const int lane = threadIdx.x & 31;
    const int per = A.NS >= 32 ? A.NS / 32 : 1;
    const Xyzz* S = A.segS + (size_t)w * A.NS;
    const Xyzz* T = A.segT + (size_t)w * A.NS;
    Xyzz run = xyzz_inf(), Bw = xyzz_inf(), Tt = xyzz_inf();
    if (lane * per < A.NS) {
        const int base = lane * per;
#pragma unroll 1
        for (int k = per - 1; k >= 1; k--) {
            run = xyzz_add_t<FqInl>(run, ld_xyzz(S + base + k));
            Bw = xyzz_add_t<FqInl>(Bw, run);                           // sum_{k>=1} k S_k
        }
        run = xyzz_add_t<FqInl>(run, ld_xyzz(S + base));               // A_l
#pragma unroll 1
        for (int k = 0; k < per; k++) Tt = xyzz_add_t<FqInl>(Tt, ld_xyzz(T + base + k));
    }
    Xyzz suf = run;
#pragma unroll 1
    for (int s2 = 1; s2 < 32; s2 <<= 1) {
        Xyzz other = shfl_xyzz(suf, (lane + s2) & 31);
        if (lane + s2 < 32) suf = xyzz_add_t<FqInl>(suf, other);
    }
    Xyzz v = lane == 0 ? xyzz_inf() : suf;
    for (int k = per; k > 1; k >>= 1) v = xyzz_dbl_t<FqInl>(v);        // * per
    v = xyzz_add_t<FqInl>(v, Bw);
    for (int k = A.SEG; k > 1; k >>= 1) v = xyzz_dbl_t<FqInl>(v);      // * SEG
    v = xyzz_add_t<FqInl>(v, Tt);
#pragma unroll 1
    for (int s2 = 16; s2 >= 1; s2 >>= 1) {
        Xyzz other = shfl_xyzz(v, (lane + s2) & 31);
        if (lane < s2) v = xyzz_add_t<FqInl>(v, other);
    }
    if (lane == 0) st_jac(A.win + w, xyzz_to_jac(v));
}

// out[prob] = sum_j 2^(c j) W_j, from the top window down
__global__ void __launch_bounds__(32) k_pip_horner(PipArgs A) {
    const unsigned prob = blockIdx.x * blockDim.x + threadIdx.x;
    if (prob >= A.n_prob) return;
    const Jac* win = A.win + (size_t)prob * A.W;
    Jac acc = ld_jac(win + A.W - 1);
#pragma unroll 1
    for (int j = A.W - 2; j >= 0; j--) {
#pragma unroll 1
        for (int k = 0; k < A.c; k++) acc = jac_dbl_t<FqInl>(acc);
        acc = jac_add_t<FqInl>(acc, ld_jac(win + j));
    }
    st_jac(A.out + prob, acc);
}

}  // namespace bppp
