/* CPU oracle, C part (TEST INFRASTRUCTURE, NOT PRODUCT): the reference's elliptic-curve hot loops
 * restated in plain C so the oracle can time "the reference's algorithm as written" on host cores.
 *
 *   ref_msm      `innerProduct` (src/Commitment.hs:325-335): normalizeBasis (:364-367, strip signs,
 *                batch-normalise the bases, BatchInverse.hs:14-24) then 256 rows of `dbl'` + one
 *                projective mixed add `nrmlAdd` (:156-169, 11 field multiplications, incomplete:
 *                P + P gives z = 0) per set scalar bit.  Zero scalars still walk every row.
 *   ref_pair_ip  `projectivePairIP` (src/Commitment.hs:343-353): the same loop over two bases and
 *                129 rows (`rationalReducedScalarLength`, :286).
 * Field arithmetic: 4 x 64-bit limbs mod q = 2^256 - 2^32 - 977 (the reference uses galois-field's
 * Natural-backed `Prime`, i.e. GMP; this is faster per operation, which only flatters the baseline).
 * PARITY UNPINNED: see oracle/__init__.py. */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fe;
typedef struct { fe x, y, z; } pp;          /* projective (X:Y:Z), identity z = 0 */

static const uint64_t P0 = 0xFFFFFFFEFFFFFC2FULL;
#define FC 0x1000003D1ULL                    /* 2^256 mod q */

static int fe_is_zero(const fe* a) { return !(a->v[0] | a->v[1] | a->v[2] | a->v[3]); }
static int fe_geq_p(const fe* a) {
    return a->v[3] == ~0ULL && a->v[2] == ~0ULL && a->v[1] == ~0ULL && a->v[0] >= P0;
}
static void fe_sub_p(fe* a) {                /* a -= q  ==  a += FC (mod 2^256) */
    u128 c = (u128)a->v[0] + FC;
    a->v[0] = (uint64_t)c; c >>= 64;
    for (int i = 1; i < 4; i++) { c += a->v[i]; a->v[i] = (uint64_t)c; c >>= 64; }
}
static void fe_add(fe* r, const fe* a, const fe* b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->v[i] + b->v[i]; r->v[i] = (uint64_t)c; c >>= 64; }
    if (c || fe_geq_p(r)) fe_sub_p(r);
}
static void fe_sub(fe* r, const fe* a, const fe* b) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a->v[i] - b->v[i] - br; r->v[i] = (uint64_t)d; br = (d >> 127) & 1; }
    if (br) {                                /* += q  ==  -= FC */
        u128 d = (u128)r->v[0] - FC; r->v[0] = (uint64_t)d; br = (d >> 127) & 1;
        for (int i = 1; i < 4; i++) { d = (u128)r->v[i] - br; r->v[i] = (uint64_t)d; br = (d >> 127) & 1; }
    }
}
static void fe_mul(fe* r, const fe* a, const fe* b) {
    uint64_t t[8] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a->v[j] * b->v[i] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
        t[i + 4] = (uint64_t)c;
    }
    /* fold high half: lo + hi * FC */
    uint64_t s[5];
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)t[4 + i] * FC + t[i]; s[i] = (uint64_t)c; c >>= 64; }
    s[4] = (uint64_t)c;
    c = (u128)s[4] * FC + s[0];
    r->v[0] = (uint64_t)c; c >>= 64;
    for (int i = 1; i < 4; i++) { c += s[i]; r->v[i] = (uint64_t)c; c >>= 64; }
    if (c) { fe_sub_p(r); }                  /* wrapped past 2^256: add FC once more */
    if (fe_geq_p(r)) fe_sub_p(r);
}
static void fe_sqr(fe* r, const fe* a) { fe_mul(r, a, a); }
static void fe_inv(fe* r, const fe* a) {     /* a^(q-2), square-and-multiply */
    static const uint64_t E[4] = {0xFFFFFFFEFFFFFC2DULL, ~0ULL, ~0ULL, ~0ULL};
    fe acc = {{1, 0, 0, 0}};
    for (int i = 255; i >= 0; i--) {
        fe_sqr(&acc, &acc);
        if ((E[i >> 6] >> (i & 63)) & 1) fe_mul(&acc, &acc, a);
    }
    *r = acc;
}
static void fe_neg(fe* r, const fe* a) { fe z = {{0, 0, 0, 0}}; fe_sub(r, &z, a); }

/* `nrmlAdd (A x2 y2) (P x1 y1 z1)` -- src/Commitment.hs:156-169 */
static void nrml_add(pp* v, const fe* x2, const fe* y2) {
    if (fe_is_zero(&v->z)) { v->x = *x2; v->y = *y2; memset(&v->z, 0, sizeof(fe)); v->z.v[0] = 1; return; }
    fe u, uu, w, vv, vvv, r, a, t, t2;
    fe_mul(&t, y2, &v->z); fe_sub(&u, &t, &v->y);
    fe_sqr(&uu, &u);
    fe_mul(&t, x2, &v->z); fe_sub(&w, &t, &v->x);
    fe_sqr(&vv, &w);
    fe_mul(&vvv, &w, &vv);
    fe_mul(&r, &vv, &v->x);
    fe_mul(&t, &uu, &v->z); fe_sub(&t, &t, &vvv); fe_add(&t2, &r, &r); fe_sub(&a, &t, &t2);
    fe_mul(&v->x, &w, &a);
    fe_sub(&t, &r, &a); fe_mul(&t, &u, &t); fe_mul(&t2, &vvv, &v->y); fe_sub(&v->y, &t, &t2);
    fe_mul(&v->z, &vvv, &v->z);
}
/* projective doubling, a = 0 (dbl-2007-bl) */
static void pp_dbl(pp* p) {
    if (fe_is_zero(&p->z)) return;
    fe xx, w, s, ss, sss, R, RR, B, h, t, t2;
    fe_sqr(&xx, &p->x);
    fe_add(&w, &xx, &xx); fe_add(&w, &w, &xx);
    fe_mul(&s, &p->y, &p->z); fe_add(&s, &s, &s);
    fe_sqr(&ss, &s);
    fe_mul(&sss, &s, &ss);
    fe_mul(&R, &p->y, &s);
    fe_sqr(&RR, &R);
    fe_add(&t, &p->x, &R); fe_sqr(&t, &t); fe_sub(&t, &t, &xx); fe_sub(&B, &t, &RR);
    fe_sqr(&h, &w); fe_add(&t, &B, &B); fe_sub(&h, &h, &t);
    fe_mul(&p->x, &h, &s);
    fe_sub(&t, &B, &h); fe_mul(&t, &w, &t); fe_add(&t2, &RR, &RR); fe_sub(&p->y, &t, &t2);
    p->z = sss;
}
static void load_fe(fe* r, const uint8_t* b) { memcpy(r->v, b, 32); }
static void store_fe(uint8_t* b, const fe* a) { memcpy(b, a->v, 32); }
static void to_affine(uint8_t out[64], const pp* p) {
    if (fe_is_zero(&p->z)) { memset(out, 0, 64); return; }
    fe zi, x, y;
    fe_inv(&zi, &p->z);
    fe_mul(&x, &p->x, &zi); fe_mul(&y, &p->y, &zi);
    store_fe(out, &x); store_fe(out + 32, &y);
}
/* `normalizes` (src/Commitment.hs:151-154): batch inversion of the z's then 2 mults per point.
 * The oracle hands bases over affine (z = 1), so this performs the reference's arithmetic on 1s. */
static void normalize_bases(size_t n, fe* xs, fe* ys) {
    fe* pre = (fe*)malloc((n + 1) * sizeof(fe));
    fe* zs = (fe*)malloc((n + 1) * sizeof(fe));
    fe acc = {{1, 0, 0, 0}};
    for (size_t i = 0; i < n; i++) { memset(&zs[i], 0, sizeof(fe)); zs[i].v[0] = 1; pre[i] = acc; fe_mul(&acc, &acc, &zs[i]); }
    fe inv;
    fe_inv(&inv, &acc);
    for (size_t i = n; i-- > 0;) {
        fe zi;
        fe_mul(&zi, &inv, &pre[i]);
        fe_mul(&inv, &inv, &zs[i]);
        fe_mul(&xs[i], &xs[i], &zi);
        fe_mul(&ys[i], &ys[i], &zi);
    }
    free(pre); free(zs);
}
/* scalars: n x 32-byte LE magnitudes of the centred lift, neg[i] = sign; points: n x (x||y), the
 * identity is 64 zero bytes.  rows = 256 (innerProduct) or 129 (projectivePairIP). */
static void straus(size_t n, const uint8_t* mags, const uint8_t* neg, const uint8_t* pts, int rows, size_t n_pad,
                   uint8_t out[64]) {
    /* n_pad = bases that carry a zero scalar (dotWith's padding, src/Commitment.hs:423-424): the
     * reference still batch-normalises them on every call and tests their bit in every row */
    if (n_pad) {
        fe* px = (fe*)malloc(n_pad * sizeof(fe));
        fe* py = (fe*)malloc(n_pad * sizeof(fe));
        for (size_t i = 0; i < n_pad; i++) { memset(&px[i], 0, sizeof(fe)); px[i].v[0] = 2 + i; py[i] = px[i]; }
        normalize_bases(n_pad, px, py);
        volatile uint64_t sink = px[n_pad - 1].v[0] ^ py[0].v[0];
        (void)sink;
        free(px); free(py);
    }
    fe* xs = (fe*)malloc((n + 1) * sizeof(fe));
    fe* ys = (fe*)malloc((n + 1) * sizeof(fe));
    uint8_t* inf = (uint8_t*)malloc(n + 1);
    for (size_t i = 0; i < n; i++) {
        load_fe(&xs[i], pts + 64 * i); load_fe(&ys[i], pts + 64 * i + 32);
        inf[i] = fe_is_zero(&xs[i]) && fe_is_zero(&ys[i]);
        if (neg[i]) fe_neg(&ys[i], &ys[i]);
    }
    normalize_bases(n, xs, ys);
    pp v;
    memset(&v, 0, sizeof v);
    v.y.v[0] = 1;
    for (int row = rows - 1; row >= 0; row--) {
        pp_dbl(&v);
        for (size_t i = 0; i < n; i++) {
            if ((mags[32 * i + (row >> 3)] >> (row & 7)) & 1) {
                if (!inf[i]) nrml_add(&v, &xs[i], &ys[i]);     /* nrmlAdd O p = p */
            }
        }
    }
    to_affine(out, &v);
    free(xs); free(ys); free(inf);
}
/* ---- an honest CPU algorithm for the same sum (baseline (ii)): Pippenger's bucket
 * method with signed windows over the same projective formulas -- what a CPU implementer would
 * write instead of the reference's 256-row Straus.  Complete additions (the bucket method adds
 * points that may coincide). */
static void pp_add(pp* r, const pp* a, const pp* b) {                /* add-1998-cmo-2, complete */
    if (fe_is_zero(&a->z)) { *r = *b; return; }
    if (fe_is_zero(&b->z)) { *r = *a; return; }
    fe y1z2, x1z2, z1z2, u, uu, v, vv, vvv, R, A, t, t2;
    fe_mul(&y1z2, &a->y, &b->z); fe_mul(&x1z2, &a->x, &b->z); fe_mul(&z1z2, &a->z, &b->z);
    fe_mul(&t, &b->y, &a->z); fe_sub(&u, &t, &y1z2);
    fe_mul(&t, &b->x, &a->z); fe_sub(&v, &t, &x1z2);
    if (fe_is_zero(&v)) {
        if (fe_is_zero(&u)) { *r = *a; pp_dbl(r); return; }
        memset(r, 0, sizeof *r); r->y.v[0] = 1; return;
    }
    fe_sqr(&uu, &u); fe_sqr(&vv, &v); fe_mul(&vvv, &v, &vv); fe_mul(&R, &vv, &x1z2);
    fe_mul(&t, &uu, &z1z2); fe_sub(&t, &t, &vvv); fe_add(&t2, &R, &R); fe_sub(&A, &t, &t2);
    fe_mul(&r->x, &v, &A);
    fe_sub(&t, &R, &A); fe_mul(&t, &u, &t); fe_mul(&t2, &vvv, &y1z2); fe_sub(&r->y, &t, &t2);
    fe_mul(&r->z, &vvv, &z1z2);
}
static void pp_madd(pp* v, const fe* x2, const fe* y2) {             /* mixed, complete */
    if (!fe_is_zero(&v->z)) {
        fe t, u, w;
        fe_mul(&t, y2, &v->z); fe_sub(&u, &t, &v->y);
        fe_mul(&t, x2, &v->z); fe_sub(&w, &t, &v->x);
        if (fe_is_zero(&w)) {
            if (fe_is_zero(&u)) { pp_dbl(v); return; }
            memset(v, 0, sizeof *v); v->y.v[0] = 1; return;
        }
    }
    nrml_add(v, x2, y2);
}
/* same interface as ref_msm (magnitudes of the centred lift + signs, affine points, identity = zeros) */
int pip_msm(size_t n, const uint8_t* mags, const uint8_t* neg, const uint8_t* pts, uint8_t out[64]) {
    int c = 4;
    while (c < 14 && ((size_t)1 << (c + 3)) <= n) c++;               /* c ~ log2(n) - 2 */
    const int windows = (256 + c - 1) / c + 1;                       /* +1 for the signed-digit carry */
    const size_t nb = (size_t)1 << (c - 1);
    fe* xs = (fe*)malloc((n + 1) * sizeof(fe));
    fe* ys = (fe*)malloc((n + 1) * sizeof(fe));
    fe* nys = (fe*)malloc((n + 1) * sizeof(fe));
    int* dig = (int*)malloc((n + 1) * windows * sizeof(int));
    for (size_t i = 0; i < n; i++) {
        load_fe(&xs[i], pts + 64 * i); load_fe(&ys[i], pts + 64 * i + 32);
        if (neg[i]) fe_neg(&ys[i], &ys[i]);
        fe_neg(&nys[i], &ys[i]);
        const int inf = fe_is_zero(&xs[i]) && fe_is_zero(&ys[i]);
        int carry = 0;
        for (int w = 0; w < windows; w++) {                           /* signed digits in [-2^(c-1), 2^(c-1)] */
            int d = carry;
            for (int k = 0; k < c; k++) {
                int bit = w * c + k;
                if (bit < 256) d += ((mags[32 * i + (bit >> 3)] >> (bit & 7)) & 1) << k;
            }
            carry = 0;
            if (d > (int)nb) { d -= 1 << c; carry = 1; }
            dig[i * windows + w] = inf ? 0 : d;
        }
    }
    pp* bucket = (pp*)malloc(nb * sizeof(pp));
    pp acc;
    memset(&acc, 0, sizeof acc); acc.y.v[0] = 1;
    for (int w = windows - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) pp_dbl(&acc);
        for (size_t m = 0; m < nb; m++) { memset(&bucket[m], 0, sizeof(pp)); bucket[m].y.v[0] = 1; }
        for (size_t i = 0; i < n; i++) {
            int d = dig[i * windows + w];
            if (d > 0) pp_madd(&bucket[d - 1], &xs[i], &ys[i]);
            else if (d < 0) pp_madd(&bucket[-d - 1], &xs[i], &nys[i]);
        }
        pp run, sum;                                                  /* sum_m m * B_m by running sums */
        memset(&run, 0, sizeof run); run.y.v[0] = 1;
        sum = run;
        for (size_t m = nb; m-- > 0;) { pp_add(&run, &run, &bucket[m]); pp_add(&sum, &sum, &run); }
        pp_add(&acc, &acc, &sum);
    }
    to_affine(out, &acc);
    free(xs); free(ys); free(nys); free(dig); free(bucket);
    return 0;
}
int ref_msm(size_t n, const uint8_t* mags, const uint8_t* neg, const uint8_t* pts, size_t n_pad, uint8_t out[64]) {
    straus(n, mags, neg, pts, 256, n_pad, out);
    return 0;
}
/* n_pairs independent two-term products sharing (b, a): out[i] = b*pts[2i] + a*pts[2i+1] */
int ref_pair_ip(size_t n_pairs, const uint8_t b_mag[32], int b_neg, const uint8_t a_mag[32], int a_neg,
                const uint8_t* pts, uint8_t* out) {
    uint8_t mags[64], neg[2] = {(uint8_t)b_neg, (uint8_t)a_neg};
    memcpy(mags, b_mag, 32); memcpy(mags + 32, a_mag, 32);
    for (size_t i = 0; i < n_pairs; i++) straus(2, mags, neg, pts + 128 * i, 129, 0, out + 64 * i);
    return 0;
}
