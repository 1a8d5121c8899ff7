// Fixed-base MSM over a FULL-MULTIPLES table in HBM (sm_100a) -- the B200-sized version of the fixed-base kernel.
//
// `commit = innerProduct . openToList` (src/Commitment.hs:416-417, 325-335) over the SHARED generator list of a batch of
// range proofs: every commitment of proveTRRPM, both commitments of every argument round (tensor mode), the
// verifier's collapsed MSM.  k_msm_gens (kernels.cuh) folds the 29 nine-bit windows of a scalar into one bucket set
// over a 2.4 MB table T[i][j] = 2^(9j) P_i: 29 mixed additions per term plus sort, bucket merge and reduction.
// A B200 has 180 GB of HBM and sustains 18 G random 64-byte reads per second (tools/micro/rand64.cu), so the table
// can hold EVERY multiple a window can ask for,
//     L[(i W + w) NB + m - 1] = m * 2^(c w) * P_i,   m = 1 .. NB = 2^(c-1),   w < W = ceil(256 / c),
// plus the carry point 2^(c W) P_i: 43 GB for the 1286 generators of examples/128by64 at c = 16.  An MSM is then
// nothing but W = 16 table lookups and mixed additions per term -- no buckets, no sort, no reduction kernel, 45 %
// fewer additions -- and the arithmetic (8M + 2S per lookup) stays the bound: at 10 k proofs/s the lookups are a
// third of what HBM delivers at random.  The window width follows the memory budget (c = 9 .. 16).
//
//   k_lut_bases   2^(c w) P_i for w = 0 .. W (Jacobian; made affine by k_batch_to_affine)
//   k_lut_fill    one thread per run of LUT_RUN consecutive multiples of one base: start point by double-and-add,
//                 then a chain of mixed additions, made affine 32 at a time (Montgomery trick in place)
//   k_msm_lut     one CTA of 64 threads per (MSM, chunk): threads take (scalar, half of its windows) units from a shared
//                 counter, recode the signed c-bit digits, prefetch the entries one unit ahead, add them (XYZZ
//                 accumulator); the CTA tree-sums its threads' partial sums
// Results are group elements, independent of the table layout: bit-identical to k_msm_gens (tests).
#pragma once
#include "kernels.cuh"

namespace bppp {

#define LUT_RUN 256                    // multiples per k_lut_fill thread (or NB when NB is smaller)
#define LUT_THREADS 64                 // few threads per MSM: the tree sum at the end is 6 levels, 2 % of a thread's work
#ifndef LUT_MIN_CTAS
#define LUT_MIN_CTAS 10                // resident CTAs per SM the register allocation aims at
#endif
#ifndef LUT_FQ
#define LUT_FQ FqCall                  // FqInl: the multiplications of the loop's one mixed addition inlined
#endif

struct LutDesc {
    Affine* tbl;                       // [P0][W][NB]
    Affine* carry;                     // [P0]   2^(c W) P_i
    int c, W, NB;
    size_t P0;
};

// bases[i * (W + 1) + w] = 2^(c w) P_i, w = 0 .. W
__global__ void __launch_bounds__(64) k_lut_bases(const Affine* __restrict__ pts, size_t n, int c, int W, Jac* __restrict__ bases) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Jac cur = jac_from_aff(ld_aff(pts + i));
    for (int w = 0; w <= W; w++) {
        st_jac(bases + i * (W + 1) + w, cur);
        if (w < W)
            for (int k = 0; k < c; k++) cur = jac_dbl(cur);
    }
}

// thread t = ((i * W + w) * runs + r): multiples r * run + 1 .. (r + 1) * run of base (i, w)
__global__ void __launch_bounds__(128) k_lut_fill(LutDesc D, const Affine* __restrict__ bases, size_t i0, size_t n_threads, int run) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    const int runs = D.NB / run;
    const size_t iw = t / runs;
    const int r = (int)(t % runs);
    const size_t i = i0 + iw / D.W;
    const int w = (int)(iw % D.W);
    const Affine B = ld_aff(bases + i * (D.W + 1) + w);
    // (r * run) * B by double-and-add (r < 2^15)
    Jac acc = jac_inf();
    {
        const unsigned k = (unsigned)r * (unsigned)run;
        for (int bit = 31 - __clz(k | 1u); bit >= 0; bit--) {
            acc = jac_dbl(acc);
            if ((k >> bit) & 1u) acc = jac_madd(acc, B);
        }
        if (k == 0) acc = jac_inf();
    }
    Affine* out = D.tbl + ((i * D.W + w) * (size_t)D.NB + (size_t)r * run);
    u256 zs[32], pre[32];
    for (int m0 = 0; m0 < run; m0 += 32) {
        u256 prod = u256_one();
#pragma unroll 1
        for (int k = 0; k < 32; k++) {
            acc = jac_madd(acc, B);                        // (r run + m0 + k + 1) B: never the identity (the order is prime and huge)
            st_u256(&out[m0 + k].x, acc.X);
            st_u256(&out[m0 + k].y, acc.Y);
            zs[k] = acc.Z;
            pre[k] = prod;
            prod = fq::mul(prod, acc.Z);
        }
        u256 inv = fq::inv(prod);
#pragma unroll 1
        for (int k = 31; k >= 0; k--) {
            const u256 zi = fq::mul(inv, pre[k]);
            inv = fq::mul(inv, zs[k]);
            const u256 zi2 = fq::sqr(zi);
            Affine a;
            a.x = fq::mul(ld_u256(&out[m0 + k].x), zi2);
            a.y = fq::mul(ld_u256(&out[m0 + k].y), fq::mul(zi2, zi));
            st_aff(out + m0 + k, a);
        }
    }
}

struct LutMsmArgs {
    LutDesc D;
    const u256* sc; size_t sc_stride, sc_out_stride;   // canonical scalars sc[p * sc_stride + o * sc_out_stride + i]
    int n_terms, chunk_terms;
    Jac* out; size_t out_pstride;                      // out[p * out_pstride + o * n_chunks + chunk]
    int n_out, n_chunks;
    unsigned long long* count;                         // profiling: += lookups (mixed additions) made; may be null
};

// Work unit = (scalar, half of its windows): units are handed out by a shared-memory counter, so the CTA stays balanced
// whatever the pattern of zero scalars (the R commitment of a round uses every other generator), and the last thread
// to finish is at most W/2 additions behind.  A thread recodes and prefetches its NEXT unit before it adds the current
// one: the random HBM reads of a unit are in flight for the ~20 us the previous unit's additions take.
#define LUT_HALF_MAX 13                // windows per unit: ceil(26 / 2) at c = 10
struct LutUnit {
    const Affine* row;                 // entries of window w0 start at row
    int n;                             // windows in this unit
    int term;                          // generator index, for the carry point; -1: no carry point in this unit
    unsigned short mag[LUT_HALF_MAX];  // |digit| - 1 (0xffff: digit 0)
    unsigned neg;                      // sign bits
};
__device__ __forceinline__ bool lut_next_unit(const LutMsmArgs& A, const u256* sc, int base, int n_units, int* counter, LutUnit& U) {
    const int c = A.D.c, W = A.D.W, W0 = (W + 1) / 2;
    const size_t NB = (size_t)A.D.NB;
    for (;;) {
        const int u = atomicAdd(counter, 1);
        if (u >= n_units) return false;
        const int i = u >> 1, h = u & 1;
        const u256 s = ld_u256(sc + i);
        if (u256_is_zero(s)) continue;
        int carry = 0;
        const int w0 = h ? W0 : 0, w1 = h ? W : W0;
        for (int w = 0; w < w0; w++) (void)signed_digit(s, w, c, carry);      // the carry into this half
        U.row = A.D.tbl + ((size_t)(base + i) * W + w0) * NB;
        U.n = w1 - w0;
        U.neg = 0;
        bool any = false;
#pragma unroll 1
        for (int w = w0; w < w1; w++) {
            const int d = signed_digit(s, w, c, carry);
            const int m = d < 0 ? -d : d;
            U.mag[w - w0] = (unsigned short)(m - 1);                           // 0xffff for a zero digit (m <= 32768)
            if (d < 0) U.neg |= 1u << (w - w0);
            if (d) {
                any = true;
#ifdef BPPP_LUT_PREFETCH_UNIT
                asm volatile("prefetch.global.L2 [%0];" ::"l"(U.row + (size_t)(w - w0) * NB + m - 1));
#endif
            }
        }
        U.term = -1;
        if (h && carry) {                                                      // top carry: the point 2^(c W) P_i
            U.term = base + i;
            any = true;
        }
        if (!any) continue;
        return true;
    }
}
__global__ void __launch_bounds__(LUT_THREADS, LUT_MIN_CTAS) k_msm_lut(LutMsmArgs A) {
    __shared__ Xyzz sm[LUT_THREADS / 2];
    __shared__ int counter;
    const int chunk = blockIdx.x, o = blockIdx.y, p = blockIdx.z, tid = threadIdx.x;
    const int base = chunk * A.chunk_terms;
    const int n = min(A.chunk_terms, A.n_terms - base);
    const u256* sc = A.sc + (size_t)p * A.sc_stride + (size_t)o * A.sc_out_stride + base;
    const size_t NB = (size_t)A.D.NB;
    if (tid == 0) counter = 0;
    __syncthreads();
    Xyzz acc = xyzz_inf();
    LutUnit cur, nxt;
    unsigned n_add = 0;
    bool have = lut_next_unit(A, sc, base, 2 * n, &counter, cur);
    while (have) {
        const bool have_next = lut_next_unit(A, sc, base, 2 * n, &counter, nxt);
        const int n_slots = cur.n + (cur.term >= 0 ? 1 : 0);              // the top carry 2^(c W) P_i is one more entry
#pragma unroll 1
        for (int k = 0; k < n_slots; k++) {
            const unsigned m = k < cur.n ? cur.mag[k] : 0u;
#ifndef BPPP_LUT_PREFETCH_UNIT
            {   // two additions (~6 us) ahead: the entry is in L2 when its turn comes, and it is still there
                const int k2 = k + 2;
                const unsigned m2 = k2 < cur.n ? cur.mag[k2] : (have_next && k2 >= n_slots && k2 - n_slots < nxt.n ? nxt.mag[k2 - n_slots] : 0xffffu);
                const Affine* r2 = k2 < cur.n ? cur.row + (size_t)k2 * NB : nxt.row + (size_t)(k2 - n_slots) * NB;
                if (m2 != 0xffffu) asm volatile("prefetch.global.L2 [%0];" ::"l"(r2 + m2));
            }
#endif
            if (m == 0xffffu) continue;
            Affine P = ld_aff(k < cur.n ? cur.row + (size_t)k * NB + m : A.D.carry + cur.term);
            if ((cur.neg >> k) & 1u) P.y = fq::neg(P.y);                  // bit cur.n is never set
            acc = xyzz_madd_t<LUT_FQ>(acc, P);                              // the ONE addition site of the loop
            n_add++;
        }
        cur = nxt;
        have = have_next;
    }
    // CTA tree sum of the per-thread partial sums
#pragma unroll 1
    for (int half = LUT_THREADS / 2; half >= 1; half >>= 1) {
        __syncthreads();
        if (tid >= half && tid < 2 * half) sm[tid - half] = acc;
        __syncthreads();
        if (tid < half) acc = xyzz_add(acc, sm[tid]);
    }
    if (tid == 0) st_jac(A.out + (size_t)p * A.out_pstride + (size_t)o * A.n_chunks + chunk, xyzz_to_jac(acc));
    if (A.count) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) n_add += __shfl_down_sync(0xffffffffu, n_add, d);
        if ((tid & 31) == 0) atomicAdd(A.count, (unsigned long long)n_add);
    }
}

}  // namespace bppp
