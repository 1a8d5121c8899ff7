// Host build (g++) of the __host__ __device__ arithmetic in fp.cuh / ec.cuh / host_math.hpp so the
// CPU test-suite can check the exact code the kernels run (portable multiply path) without a GPU.
#include <string.h>
#include "ec.cuh"
#include "host_math.hpp"
using namespace bppp;
static u256 ld(const uint8_t* b) { return host::from_bytes(b); }
static void st(uint8_t* b, const u256& a) { host::to_bytes(b, a); }
static Affine lda(const uint8_t* b) { Affine p; p.x = ld(b); p.y = ld(b + 32); return p; }
static void sta(uint8_t* b, const Affine& p) { st(b, p.x); st(b + 32, p.y); }
extern "C" {
int ht_field(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    u256 x = ld(a), y = ld(b), r;
    switch (op) {
        case 0: r = fq::mul(x, y); break;
        case 1: r = fq::add(x, y); break;
        case 2: r = fq::sub(x, y); break;
        case 3: r = fq::inv(x); break;
        case 4: r = fr::mul(x, y); break;
        case 5: r = fr::add(x, y); break;
        case 6: r = fr::sub(x, y); break;
        case 7: r = fr::to_mont(x); break;
        case 8: r = fr::from_mont(x); break;
        case 10: r = fr::from_mont(host::fr_inv(fr::to_mont(x))); break;
        default: return 1;
    }
    st(out, r);
    return 0;
}
int ht_ec(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    Affine x = lda(a), y = lda(b);
    Jac r;
    switch (op) {
        case 0: r = jac_madd(jac_from_aff(x), y); break;
        case 1: r = jac_dbl(jac_from_aff(x)); break;
        case 2: r = jac_add(jac_dbl(jac_from_aff(x)), jac_madd(jac_from_aff(y), x)); break;
        default: return 1;
    }
    sta(out, jac_to_aff(r));
    return 0;
}
int ht_jsf(const uint8_t* k0, const uint8_t* k1, uint8_t* digits, int max_digits) {
    return jsf_recode(digits, ld(k0), ld(k1), max_digits);
}
int ht_signed_digits(const uint8_t* s, int c, int w, int* out) {
    int carry = 0;
    u256 v = ld(s);
    for (int j = 0; j < w; j++) out[j] = signed_digit(v, j, c, carry);
    return carry;
}
int ht_rational_reduce(const uint8_t* x, uint8_t* a, int* a_neg, uint8_t* b, int* b_neg) {
    host::Ratio r = host::rational_reduce(ld(x));
    st(a, r.a); st(b, r.b); *a_neg = r.a_neg; *b_neg = r.b_neg;
    return 0;
}
// the pair-fold chain exactly as k_pair_fold runs it (fast + slow path), single pair
int ht_pair_fold(const uint8_t* kb, int b_neg, const uint8_t* ka, int a_neg, const uint8_t* pl, const uint8_t* pr, uint8_t* out) {
    uint8_t dig[264];
    int ndig = jsf_recode(dig, ld(kb), ld(ka), 264);
    Affine PL = aff_cneg(lda(pl), b_neg), PR = aff_cneg(lda(pr), a_neg);
    Jac acc = jac_inf();
    bool fast = !aff_is_inf(PL) && !aff_is_inf(PR) && !u256_eq(PL.x, PR.x);
    u256 H = u256_one();
    Affine t[4];
    if (fast) {
        H = fq::sub(PR.x, PL.x);
        u256 HH = fq::sqr(H), HHH = fq::mul(H, HH);
        t[0].x = fq::mul(PL.x, HH); t[0].y = fq::mul(PL.y, HHH);
        t[1].x = fq::mul(PR.x, HH); t[1].y = fq::mul(PR.y, HHH);
        u256 rp = fq::sub(PR.y, PL.y), V2 = fq::dbl(t[0].x);
        t[2].x = fq::sub(fq::sub(fq::sqr(rp), HHH), V2);
        t[2].y = fq::sub(fq::mul(rp, fq::sub(t[0].x, t[2].x)), t[0].y);
        u256 rm = fq::neg(fq::add(PR.y, PL.y));
        t[3].x = fq::sub(fq::sub(fq::sqr(rm), HHH), V2);
        t[3].y = fq::sub(fq::mul(rm, fq::sub(t[0].x, t[3].x)), t[0].y);
    } else { t[0] = PL; t[1] = PR; }
    for (int j = ndig - 1; j >= 0; j--) {
        acc = jac_dbl(acc);
        int d = dig[j], u0 = (d & 3) - 1, u1 = ((d >> 2) & 3) - 1;
        if (!u0 && !u1) continue;
        if (fast) {
            int e; bool neg;
            if (u1 == 0) { e = 0; neg = u0 < 0; } else if (u0 == 0) { e = 1; neg = u1 < 0; }
            else if (u0 == u1) { e = 2; neg = u0 < 0; } else { e = 3; neg = u0 < 0; }
            acc = jac_madd(acc, aff_cneg(t[e], neg));
        } else {
            if (u0) acc = jac_madd(acc, aff_cneg(t[0], u0 < 0));
            if (u1) acc = jac_madd(acc, aff_cneg(t[1], u1 < 0));
        }
    }
    if (fast && !jac_is_inf(acc)) acc.Z = fq::mul(acc.Z, H);
    sta(out, jac_to_aff(acc));
    return fast ? 1 : 0;
}
}
