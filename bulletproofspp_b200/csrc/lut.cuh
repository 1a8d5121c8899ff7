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
//   k_msm_lut     one CTA per (MSM, chunk): a thread recodes its scalars into signed c-bit digits, prefetches the W
//                 entries, adds them (XYZZ accumulator), and the CTA tree-sums its threads' partial sums
// Results are group elements, independent of the table layout: bit-identical to k_msm_gens (tests).
#pragma once
#include "kernels.cuh"

namespace bppp {

#define LUT_RUN 256                    // multiples per k_lut_fill thread (or NB when NB is smaller)
#define LUT_THREADS 256

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
};

__global__ void __launch_bounds__(LUT_THREADS, 2) k_msm_lut(LutMsmArgs A) {
    __shared__ Xyzz sm[LUT_THREADS / 2];
    const int chunk = blockIdx.x, o = blockIdx.y, p = blockIdx.z, tid = threadIdx.x;
    const int base = chunk * A.chunk_terms;
    const int n = min(A.chunk_terms, A.n_terms - base);
    const u256* sc = A.sc + (size_t)p * A.sc_stride + (size_t)o * A.sc_out_stride + base;
    const int c = A.D.c, W = A.D.W;
    const size_t NB = (size_t)A.D.NB;
    Xyzz acc = xyzz_inf();
#pragma unroll 1
    for (int i = tid; i < n; i += LUT_THREADS) {
        const u256 s = ld_u256(sc + i);
        if (u256_is_zero(s)) continue;
        const Affine* row = A.D.tbl + (size_t)(base + i) * W * NB;
        // pass 1: the digits, and a prefetch of every entry this scalar needs (HBM latency hides behind the additions)
        int carry = 0;
        unsigned long long neg = 0;                     // sign bits of the W digits
        int dig[32];                                    // W <= 26 (c >= 10)
#pragma unroll 1
        for (int w = 0; w < W; w++) {
            const int d = signed_digit(s, w, c, carry);
            dig[w] = d < 0 ? -d : d;
            if (d < 0) neg |= 1ull << w;
            if (d) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + (size_t)w * NB + dig[w] - 1));
        }
#pragma unroll 1
        for (int w = 0; w < W; w++) {
            if (!dig[w]) continue;
            Affine P = ld_aff(row + (size_t)w * NB + dig[w] - 1);
            if ((neg >> w) & 1ull) P.y = fq::neg(P.y);
            acc = xyzz_madd(acc, P);
        }
        if (carry) acc = xyzz_madd(acc, ld_aff(A.D.carry + base + i));
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
}

}  // namespace bppp
