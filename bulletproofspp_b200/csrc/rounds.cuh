// Per-proof round state of the norm-linear argument on the device (sm_100a).
//
// proveRoundM / proveBPM (src/Bulletproof.hs:346-359) alternate two commitments with a challenge-dependent
// fold.  The per-proof scalars of a round -- rationalReduceScalar of the challenge (src/Commitment.hs:242-288),
// the fold factors of NormArgument.hs:56-71,113-129, the running normalisations n, the weight q and the opening
// scalar s (Bulletproof.hs:352-353) -- are a few hundred field operations per proof.  Kept on the host they cost
// two stream synchronisations and a dozen small uploads per round; here they are three small kernels, so a
// whole argument (commitments, transcript, folds) is one stream of launches with a single synchronisation at
// its end.  Bit-identical to the host sequencing in capi.cu (bppp_nl_round_commit / bppp_nl_round_fold), which
// stays the reference for the step-by-step ABI.
#pragma once
#include "kernels.cuh"
#include "host_math.hpp"

namespace bppp {

struct RoundState {             // device arrays, one entry per proof unless noted; Montgomery form unless noted
    u256 *q, *qinv, *nn, *nl, *s;
    u256 *rho, *k1, *k2, *coef;                     // coef: [B][8]
    u256 *au, *bu, *al, *bl, *ac, *bc, *a0n, *b0n, *a0l, *b0l;
    u256 *kb, *ka;                                  // [B][2] magnitudes of b, a (canonical integers) for k_pair_fold
    unsigned char* sgn;                             // [B][2] bit 0: b < 0, bit 1: a < 0
    u256* inv;                                      // [B][2] scratch: 1 / b0 of the norm and the linear part
    const u256* chal;                               // [B] this round's challenge, canonical
    const u256* dots;                               // [B][2] sX, sR
    int B;
    // tensor mode: the generators are never folded, only their coefficients coef_idx *= (a0 | b0) and the scalar
    // vectors x' = (xL + e q xR) / b0 with n' = n b0 / q.  Every output (X, R, s, the final n * x) depends on
    // a0 / b0 = e / q (norm) and e (linear) alone, so the representative a0 = e / q, b0 = 1 gives the same bits
    // without rationalReduceScalar and without an inversion; the short (a', b') of src/Commitment.hs:242-288
    // only pays for itself where generators are really folded (k_pair_fold's half-length scalars).
    int tensor;
    // a contiguous shard of a larger vector: the pair weights of this shard start at rho^pair_off
    unsigned long long pair_off;
};

// constants of a round's commitments (NormArgument.hs:113): rho = q^4, k1 = 2 n^2 q^3, k2 = n^2 q^4;
// X scalars: q * xR on the left generators, q^-1 * xL on the right ones
__global__ void __launch_bounds__(128) k_round_pre(RoundState S) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.B) return;
    const u256 q = ld_u256(S.q + b), nn = ld_u256(S.nn + b);
    const u256 q2 = fr::sqr(q), q3 = fr::mul(q2, q), q4 = fr::sqr(q2), n2 = fr::sqr(nn);
    st_u256(S.rho + b, q4);
    u256 k1 = fr::dbl(fr::mul(n2, q3)), k2 = fr::mul(n2, q4);
    if (S.pair_off) {                               // weights of a shard start at (q^4)^(first pair index)
        u256 off = fr::one(), base = q4;
        for (unsigned long long e = S.pair_off; e; e >>= 1) {
            if (e & 1) off = fr::mul(off, base);
            base = fr::sqr(base);
        }
        k1 = fr::mul(k1, off);
        k2 = fr::mul(k2, off);
    }
    st_u256(S.k1 + b, k1);
    st_u256(S.k2 + b, k2);
    st_u256(S.coef + (size_t)b * 8 + 1, q);
    st_u256(S.coef + (size_t)b * 8 + 2, ld_u256(S.qinv + b));
}

// thread (b, seg): seg 0 = norm part, (a', b') = rationalReduceScalar (e * qInv)  (NormArgument.hs:125);
//                  seg 1 = linear part, (a', b') = rationalReduceScalar e          (NormArgument.hs:66)
__global__ void __launch_bounds__(64) k_round_ratio(RoundState S) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * S.B) return;
    const int b = t >> 1, seg = t & 1;
    const u256 e = ld_u256(S.chal + b);
    u256 x = e;
    if (seg == 0) x = fr::from_mont(fr::mul(fr::to_mont(e), ld_u256(S.qinv + b)));
    const host::Ratio r = host::rational_reduce(x);
    st_u256(S.kb + t, r.b);
    st_u256(S.ka + t, r.a);
    S.sgn[t] = (unsigned char)((r.b_neg ? 1 : 0) | (r.a_neg ? 2 : 0));
    const u256 b0 = host::fr_from_signed(r.b, r.b_neg), a0 = host::fr_from_signed(r.a, r.a_neg);
    st_u256((seg ? S.b0l : S.b0n) + b, b0);
    st_u256((seg ? S.a0l : S.a0n) + b, a0);
    st_u256(S.inv + t, fr::inv(b0));
}

// the fold factors and the state update (the second half of bppp_nl_round_fold's host loop)
__global__ void __launch_bounds__(128) k_round_post(RoundState S) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.B) return;
    const u256 em = fr::to_mont(ld_u256(S.chal + b));
    const u256 q = ld_u256(S.q + b), qinv = ld_u256(S.qinv + b);
    u256 inv_n, inv_l, b0n, b0l, a0l;
    if (S.tensor) {
        inv_n = inv_l = b0n = b0l = fr::one();
        a0l = em;
        st_u256(S.a0n + b, fr::mul(em, qinv)); st_u256(S.b0n + b, b0n);
        st_u256(S.a0l + b, a0l); st_u256(S.b0l + b, b0l);
    } else {
        inv_n = ld_u256(S.inv + 2 * b); inv_l = ld_u256(S.inv + 2 * b + 1);
        b0n = ld_u256(S.b0n + b); b0l = ld_u256(S.b0l + b); a0l = ld_u256(S.a0l + b);
    }
    // x' = b0Inv*xL + e*q*b0Inv*xR   (NormArgument.hs:129);  l' = b0Inv*xL + e*b0Inv*xR  (:71);  c' = b0*cL + a0*cR
    st_u256(S.au + b, inv_n);
    st_u256(S.bu + b, fr::mul(fr::mul(em, q), inv_n));
    st_u256(S.al + b, inv_l);
    st_u256(S.bl + b, fr::mul(em, inv_l));
    st_u256(S.ac + b, b0l);
    st_u256(S.bc + b, a0l);
    // s' = s + e*sX + (e^2 - 1)*sR   (Bulletproof.hs:352-353, makeEs NormArgument.hs:109)
    const u256 e1 = fr::sub(fr::sqr(em), fr::one());
    const u256 s = fr::add(ld_u256(S.s + b), fr::add(fr::mul(em, ld_u256(S.dots + 2 * b)), fr::mul(e1, ld_u256(S.dots + 2 * b + 1))));
    st_u256(S.s + b, s);
    // n <- n*b0*qInv ; q <- q^2 ; linear n <- n*b0
    st_u256(S.nn + b, fr::mul(fr::mul(ld_u256(S.nn + b), b0n), qinv));
    st_u256(S.nl + b, fr::mul(ld_u256(S.nl + b), b0l));
    const u256 qn = fr::sqr(q);
    st_u256(S.q + b, qn);
    st_u256(S.qinv + b, fr::sqr(qinv));
    st_u256(S.rho + b, fr::sqr(fr::sqr(qn)));
}

// One large argument sharded over GPUs: a rank's partial commitments (X, R as Jacobian points) and
// partial scalar parts (sX, sR) packed into one 256-byte record for the all-gather ...
#define SHARD_REC_BYTES 256            // 2 * 96 + 2 * 32
__global__ void k_shard_pack(const Jac* __restrict__ res, const u256* __restrict__ dots, unsigned char* __restrict__ rec) {
    if (threadIdx.x || blockIdx.x) return;
    Jac* o = reinterpret_cast<Jac*>(rec);
    st_jac(o, ld_jac(res));
    st_jac(o + 1, ld_jac(res + 1));
    u256* d = reinterpret_cast<u256*>(rec + 192);
    st_u256(d, ld_u256(dots));
    st_u256(d + 1, ld_u256(dots + 1));
}
// ... and the sums over all ranks in rank order (EC addition is not an NCCL reduction): thread 0 -> X, 1 -> R, 2 -> sX, sR
__global__ void k_shard_combine(const unsigned char* __restrict__ recs, int world, Jac* __restrict__ res, u256* __restrict__ dots) {
    const int t = threadIdx.x;
    if (blockIdx.x || t > 2) return;
    if (t < 2) {
        Jac acc = jac_inf();
        for (int r = 0; r < world; r++) acc = jac_add(acc, ld_jac(reinterpret_cast<const Jac*>(recs + (size_t)r * SHARD_REC_BYTES) + t));
        st_jac(res + t, acc);
    } else {
        u256 a = u256_zero(), b = u256_zero();
        for (int r = 0; r < world; r++) {
            const u256* d = reinterpret_cast<const u256*>(recs + (size_t)r * SHARD_REC_BYTES + 192);
            a = fr::add(a, ld_u256(d));
            b = fr::add(b, ld_u256(d + 1));
        }
        st_u256(dots, a);
        st_u256(dots + 1, b);
    }
}

// getWitness of the final round (NormArgument.hs:147-163 with the stored normalisation): out[b][i] = n_b * v[b][i],
// canonical; `scale` may be null (plain conversion)
__global__ void __launch_bounds__(128) k_scale_rows(const u256* __restrict__ v, size_t v_stride, const u256* __restrict__ scale, int n, int B,
                                                   u256* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * B) return;
    const int b = t / n, i = t % n;
    u256 x = ld_u256(v + (size_t)b * v_stride + i);
    if (scale) x = fr::mul(x, ld_u256(scale + b));
    st_u256(out + t, fr::from_mont(x));
}

}  // namespace bppp
