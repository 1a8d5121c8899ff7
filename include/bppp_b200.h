/* bppp_b200 -- C ABI of the B200-native Bulletproofs++ argument hot path.
 *
 * The reference (Liam-Eagen/BulletproofsPP, pure Haskell) has no FFI of its own: its "plugin
 * API" is a set of typeclasses.  These entry points are what a `foreign import ccall` shim
 * under those classes binds; each cites the reference interface it replaces.
 *
 * Conventions
 *   - scalars and coordinates are 32-byte LITTLE-endian canonical integers (< modulus);
 *     Fr = secp256k1 group order, Fq = secp256k1 base field;
 *   - an affine point is x||y (64 bytes); the identity is 64 zero bytes; output points are
 *     affine and fully reduced (the transcript hashes decimal x, y of `toA`, app/Main.hs:78-80);
 *   - every function returns 0 on success, non-zero on error (see bppp_last_error); nothing
 *     throws across the ABI; the reference's own convention is Maybe/panic (app/Main.hs:155-169);
 *   - caller owns every in/out buffer; the library owns device memory behind the opaque handles;
 *   - results are a pure function of the inputs; calls are thread-safe per ctx/handle (each call
 *     sets the device and uses the context's own stream), so they can be imported `safe` from
 *     a -threaded GHC RTS (package.yaml:74-78);
 *   - there is NO CPU fallback: without a CUDA device bppp_init fails.
 */
#ifndef BPPP_B200_H
#define BPPP_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bppp_ctx bppp_ctx;
typedef struct bppp_nl bppp_nl;

enum { BPPP_OK = 0, BPPP_ERR_ARG = 1, BPPP_ERR_CUDA = 2, BPPP_ERR_RANGE = 3, BPPP_ERR_STATE = 4 };
enum { BPPP_ARG_NL = 0, BPPP_ARG_IP = 1 };

int bppp_init(int device, bppp_ctx** out);
void bppp_free(bppp_ctx* ctx);
const char* bppp_last_error(bppp_ctx* ctx);
/* OPT-IN process-wide tuning for a dedicated batch-proving process (bench.py and the Python harness
 * call it; a library embedded under a GHC RTS should not): BPPP_TUNE_MALLOC keeps freed host memory
 * in the malloc arenas (mallopt), BPPP_TUNE_DEVICE makes contexts created afterwards request
 * cudaDeviceScheduleBlockingSync and pre-grow the device memory pool.  Nothing of this happens
 * without the call. */
enum { BPPP_TUNE_MALLOC = 1, BPPP_TUNE_DEVICE = 2 };
int bppp_tune_process(int flags);
/* ABI version; bumped on any signature change */
int bppp_abi_version(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t bppp_launch_count(bppp_ctx* ctx);
/* block until every queued operation of the context has finished */
int bppp_sync(bppp_ctx* ctx);

/* ---- measurement support (bench.py): per-kernel CUDA-event timing on the launching stream,
 * host<->device copy counters, a stream timer and the measured integer peak. */
int bppp_profile_enable(bppp_ctx* ctx, int on);
int bppp_profile_reset(bppp_ctx* ctx);
int bppp_profile_report(bppp_ctx* ctx, char* out_json, size_t cap);
int bppp_timer_start(bppp_ctx* ctx);
int bppp_timer_stop(bppp_ctx* ctx, double* ms);
int bppp_measure_imad_peak(bppp_ctx* ctx, double* wide_per_s, double* lo_per_s);

/* ---- MSM seam: `commit` = `innerProduct . openToList` (src/Commitment.hs:416-417, 325-335),
 * i.e. FastInnerProduct.innerProduct :: [(Scalar v, v)] -> v.  out = sum_i scalars[i] * points[i]. */
int bppp_msm(bppp_ctx* ctx, size_t n, const uint8_t* scalars, const uint8_t* points, uint8_t out[64]);
/* `batch` independent MSMs of n terms each.  scalars: batch*n*32 bytes.  points: n*64 bytes when
 * shared_points != 0 (the range-proof commitments over one generator list, commitRPW,
 * src/RangeProof/Internal.hs:43-48), else batch*n*64.  out: batch*64. */
int bppp_msm_batch(bppp_ctx* ctx, size_t batch, size_t n, const uint8_t* scalars, const uint8_t* points,
                   int shared_points, uint8_t* out);

/* ---- the shared generator list [g | G | H] resident on the device with its fixed-base window
 * table (2^(9j) * P_i): used by round 1 of every argument, every range-proof commitment
 * (commitRPW, src/RangeProof/Internal.hs:43-48) and the verifier's collapsed MSM.  The reference's
 * analogue is the generator cache points.bin (app/Main.hs:147,259-263). */
typedef struct bppp_gens bppp_gens;
int bppp_gens_create(bppp_ctx* ctx, size_t N, size_t M, const uint8_t* g, const uint8_t* G, const uint8_t* H,
                     bppp_gens** out);
void bppp_gens_destroy(bppp_gens* g);
/* `batch` MSMs over the first n generators of the list: scalars batch*n*32, out batch*64 */
int bppp_gens_msm_batch(bppp_gens* g, size_t batch, size_t n, const uint8_t* scalars, uint8_t* out);
/* Full-multiples table for the list (csrc/lut.cuh): every multiple m * 2^(c w) * P_i a signed c-bit window can ask for,
 * resident in HBM -- an MSM over the list becomes ceil(256/c) table lookups + mixed additions per term (no buckets, no
 * reduction).  budget_gb bounds the table (43 GB for 1286 generators at c = 16; c = 10 .. 16 by what fits, and at most
 * half of the free device memory); tables are shared between generator sets over the same list on one device.
 * *c_out = the window width in use (0: no table, the nine-bit bucket kernel stays).  Results are unchanged.  Every batch
 * size takes the table, a lone proof included (32-term chunks: 118 us per commitment against 320 us). */
int bppp_gens_enable_lut(bppp_gens* g, double budget_gb, int* c_out);
/* host threads used by the round sequencing inside the device entry points (0 = all cores) */
void bppp_set_device_host_threads(int n);
void bppp_set_thread_host_threads(int n);   /* same, for the calling thread only */
int bppp_ctx_device(bppp_ctx* ctx);
/* page-locked host staging memory for the in/out buffers of the batch entry points */
int bppp_pinned_alloc(size_t bytes, void** out);
void bppp_pinned_free(void* p);

/* ---- the reference's hash-derived objects on the device ("transcript on device"); bit-identical
 * to the host transcript (bppp_host_*) and the oracle.
 * getPoints seed (app/Main.hs:68-72): the first `count` generators, out = count*64 bytes */
int bppp_get_points(bppp_ctx* ctx, const char* seed, size_t count, int root_policy, uint8_t* out);
/* ZKPT's commitment list (src/ZKP.hs:68-101) for `batch` proofs in lock-step, at most max_points commitments each */
typedef struct bppp_dtr bppp_dtr;
int bppp_dtr_create(bppp_ctx* ctx, size_t batch, size_t max_points, int show_format, bppp_dtr** out);
void bppp_dtr_destroy(bppp_dtr* t);
int bppp_dtr_reset(bppp_dtr* t);
/* `oracle xs` / `oracle'` (src/ZKP.hs:96-101, 60-63; shaOracle app/Main.hs:75-80): prepend pts = [batch][npts]
 * to every proof's list, then out = [batch][count] = the first `count` (<= 9) scalars of shaOracle */
int bppp_dtr_oracle(bppp_dtr* t, const uint8_t* pts, size_t npts, int count, uint8_t* out);
/* (bppp_dtr_squeeze_inv also returns inv_out[b][j] = 1 / out[b][j], inverted on the device: the provers and verifiers
 * need e^-1, q^-1, ... at once)
 * the two halves separately, for a verifier that knows every commitment up front (verifyBPM's `oracle'` calls,
 * src/Bulletproof.hs:370-378): bppp_dtr_absorb is cs' = xs ++ cs alone -- row b of the call is the `npts` points at
 * pts + 64 * stride_points * b; bppp_dtr_squeeze then hashes any number of stages in ONE launch: out[b][j] = scalar
 * idx[j] (1-based, <= 9) of the transcript as it was after state[j] absorb calls (state NULL or 0: all of them).
 * Every earlier transcript is a suffix of the latest one (newest commitments first). */
int bppp_dtr_absorb(bppp_dtr* t, const uint8_t* pts, size_t stride_points, size_t npts);
int bppp_dtr_squeeze(bppp_dtr* t, size_t n_chal, const uint8_t* idx, const uint8_t* state, uint8_t* out);
int bppp_dtr_squeeze_inv(bppp_dtr* t, size_t n_chal, const uint8_t* idx, const uint8_t* state, uint8_t* out, uint8_t* inv_out);
int bppp_dtr_fits(bppp_dtr* t, size_t batch, size_t max_points, int show_format);
/* the rendered list of one proof, concat [show x <> show y] newest first (what app/Main.hs:78-80 feeds the hash) */
int bppp_dtr_export(bppp_dtr* t, size_t proof, uint8_t* out, size_t cap, size_t* len);
/* `random` (src/ZKP.hs:90-93 with hashToScalar randomSeed . show, app/Main.hs:177): out[b][j] = draw n0 + j of proof b */
int bppp_dev_random(bppp_ctx* ctx, size_t batch, const char* const* seeds, uint64_t n0, size_t count, uint8_t* out);

/* ---- fixed-base MSMs over a handful of generators shared by every call: the range proofs'
 * input commitments value*g + type*hs0 + blind*hs1 (scalarRPW' / scalarPairRPW' + commitRPW,
 * src/RangeProof/Internal.hs:43-57; app/Main.hs:287-288,315).  Precomputed 8-bit window tables. */
typedef struct bppp_fb bppp_fb;
int bppp_fb_create(bppp_ctx* ctx, size_t n_bases, const uint8_t* points, bppp_fb** out);
/* scalars: batch*n_bases*32; out: batch*64 */
int bppp_fb_msm_batch(bppp_fb* fb, size_t batch, const uint8_t* scalars, uint8_t* out);
void bppp_fb_destroy(bppp_fb* fb);

/* ---- generator fold: collapsePoints b a gL gR = projectivePairIP (b, gL) (a, gR)
 * (src/Bulletproof.hs:213-214, src/Commitment.hs:343-353), for a whole vector with one (a, b):
 * out[i] = (+-b)*in[2i] + (+-a)*in[2i+1]; an odd tail pairs with the identity.
 * a, b: magnitudes (ReducedScalar), *_neg their signs.  points_in: n_in*64, points_out: ceil(n_in/2)*64. */
int bppp_pair_fold(bppp_ctx* ctx, size_t n_in, const uint8_t a[32], int a_neg, const uint8_t b[32], int b_neg,
                   const uint8_t* points_in, uint8_t* points_out);

/* ---- rationalReduceScalar (src/Commitment.hs:242-255): host half-GCD, a = b*x (mod r), a^2 <= 2r */
int bppp_rational_reduce(const uint8_t x[32], uint8_t a[32], int* a_neg, uint8_t b[32], int* b_neg);

/* ---- Argument seam: a device-resident NormLinear argument state, lock-step over `batch` proofs.
 * Replaces NormLinearBP / BPOpening (src/Bulletproof.hs:179-207, 276-291) for
 * NL.NormLinear (src/Bulletproof/NormArgument.hs:153-178) when kind = BPPP_ARG_NL and
 * IP.NormLinear (src/Bulletproof/InnerProductArgument.hs:239-267) when kind = BPPP_ARG_IP.
 *
 * bppp_nl_create = makeNormLinearBP' 1 q cs nss ngs lss lgs + makePSV sc g:
 *   g (64), G (N*64), H (M*64): generators, shared by the whole batch;
 *   per proof b: q[b] (for IP: the r with q = r^4), s[b] (scalar on g), w[b][N] (norm witness),
 *   l[b][M] (linear witness), c[b][M] (public linear coefficients). */
int bppp_nl_create(bppp_ctx* ctx, int kind, size_t batch, size_t N, size_t M, const uint8_t* g, const uint8_t* G,
                   const uint8_t* H, const uint8_t* q, const uint8_t* s, const uint8_t* w, const uint8_t* l,
                   const uint8_t* c, bppp_nl** out);
/* same, over a resident generator list (no table rebuild per call) */
int bppp_nl_create_gens(bppp_gens* gens, int kind, size_t batch, const uint8_t* q, const uint8_t* s, const uint8_t* w,
                        const uint8_t* l, const uint8_t* c, bppp_nl** out);
/* makeScalarsComs + the two `commit`s of proveRoundM (src/Bulletproof.hs:346-350):
 * X[b], R[b] (64 bytes each; L, R for the IP argument). */
int bppp_nl_round_commit(bppp_nl* h, uint8_t* X, uint8_t* R);
/* the same plus E = [batch] challenges  e <- head <$> oracle [X, R]  (Bulletproof.hs:351) from the device
 * transcript the argument continues (arguments made by bppp_nl_create_trrp after bppp_trrp_*_tr calls) */
int bppp_nl_round_challenge(bppp_nl* h, uint8_t* X, uint8_t* R, uint8_t* E);
/* proveBPM (Bulletproof.hs:357-359) as one stream of launches: commitments, `oracle [X, R]` on the device
 * transcript, rationalReduceScalar and the fold factors (rounds.cuh), folds -- a single synchronisation at the end.
 * Fresh norm-linear handle with a transcript (from bppp_nl_create_trrp after the _tr phases, or attached with
 * bppp_nl_attach_transcript, e.g. after bppp_dtr_absorb of the initial commitment).  responses = [batch][rounds][2]
 * points and es = [batch][rounds] challenges (may be NULL), newest first; s / w / l as bppp_nl_final. */
int bppp_nl_attach_transcript(bppp_nl* h, bppp_dtr* t);
/* ---- one large argument over several GPUs: NCCL inside the library.  libnccl.so.2 is loaded at run
 * time (bppp_comm_load(path), or the copy already in the process / the system's), so single-GPU users never need it.
 * bppp_comm_unique_id on rank 0 -> hand the 128 bytes to every rank -> bppp_comm_create(ctx, world, rank, id).
 * bppp_nl_prove_sharded: rank r holds the contiguous slice [r * len, (r + 1) * len) of the norm vector and its
 * generators (bppp_nl_create over the slice + bppp_nl_set_shard(r * len); len a power of two), rank 0 also the linear
 * part (lin_len entries; the other ranks create M = 0); same q, s and transcript state everywhere.  `local_rounds`
 * rounds fold locally with the per-round partial commitments all-gathered (256 B per rank) and summed on the device;
 * then the slices are gathered once and every rank finishes on the short whole argument.  Outputs on every rank as
 * bppp_nl_prove_device; the proof equals the unsharded one bit for bit. */
typedef struct bppp_comm bppp_comm;
int bppp_comm_load(const char* libnccl_path);
const char* bppp_comm_last_error(void);
int bppp_comm_unique_id(uint8_t id[128]);
int bppp_comm_create(bppp_ctx* ctx, int world, int rank, const uint8_t id[128], bppp_comm** out);
void bppp_comm_destroy(bppp_comm* c);
int bppp_nl_prove_sharded(bppp_nl* h, bppp_comm* comm, size_t rounds, size_t local_rounds, size_t lin_len,
                          uint8_t* responses, uint8_t* es, uint8_t* s, uint8_t* w, uint8_t* l);
int bppp_nl_prove_device(bppp_nl* h, size_t rounds, uint8_t* responses, uint8_t* es, uint8_t* s, uint8_t* w, uint8_t* l);
/* the rest of proveRoundM (src/Bulletproof.hs:351-355): s' = s + e0*sX + e1*sR and `collapse e`
 * (NormArgument.hs:64-71,123-129 / InnerProductArgument.hs:86-101,155-170) with challenge e[b]. */
int bppp_nl_round_fold(bppp_nl* h, const uint8_t* e);
/* proveBPM (src/Bulletproof.hs:357-359): all `rounds` rounds on a device-resident argument with the
 * reference's Fiat-Shamir transcript on the host (e <- head <$> oracle [X, R], Bulletproof.hs:351; shaOracle,
 * app/Main.hs:75-80).  init_pts = [batch][n_init] commitments already in the transcript, newest first.
 * Outputs newest first: responses [batch][rounds][2] points, es [batch][rounds] challenges (may be NULL). */
int bppp_nl_prove(bppp_nl* h, size_t batch, int show_format, size_t n_init, const uint8_t* init_pts, size_t rounds,
                  uint8_t* responses, uint8_t* es);
/* the verifier's half of that transcript (verifyBPM, src/Bulletproof.hs:370-378): es from the responses */
int bppp_nl_challenges(size_t batch, int show_format, size_t n_init, const uint8_t* init_pts, size_t rounds,
                       const uint8_t* responses, uint8_t* es);
/* sharding one large argument: this handle holds the slice of the norm vector that
 * starts at `first_element`; bppp_nl_export hands the stored-form state to the rank that finishes */
int bppp_nl_set_shard(bppp_nl* h, size_t first_element);
int bppp_nl_export(bppp_nl* h, uint8_t* nn, uint8_t* nl, uint8_t* points, uint8_t* c);
/* current lengths after the folds so far (norm vector length as `getWitness` reports it) */
int bppp_nl_lengths(bppp_nl* h, size_t* n_norm, size_t* n_lin);
/* scalarCP s and getWitness (NormArgument.hs:62,121 / InnerProductArgument.hs:154,222-223): s[b], w[b][n_norm], l[b][n_lin] */
int bppp_nl_final(bppp_nl* h, uint8_t* s, uint8_t* w, uint8_t* l);
void bppp_nl_destroy(bppp_nl* h);

/* verifyBPM's collapsed check (src/Bulletproof.hs:370-378, 362-368; expandChallenges
 * NormArgument.hs:73-81,131-145 / InnerProductArgument.hs:103-124,172-181) for `batch` proofs over shared
 * generators: ok[b] = [ (s_pub - sc)*g + sum (pub_i - tensor_i)*G_i + sum (0 - tensor_j)*H_j
 *                       + sum_k init_s[b][k]*init_p[b][k] + sum_rounds (e0*X + e1*R) == 0 ].
 * es[b][k]: challenges NEWEST FIRST (as verifyBPM builds them); XR[b][k][2]: responses newest
 * first; pub_w[b][N], s_pub[b]: public norm vector / scalar; c[b][M]: public linear coefficients;
 * fw[b][n_norm], fl[b][n_lin]: the proof's final witness scalars; init_*: the opening of initCom. */
int bppp_nl_verify(bppp_ctx* ctx, int kind, size_t batch, size_t N, size_t M, size_t k, const uint8_t* g,
                   const uint8_t* G, const uint8_t* H, const uint8_t* q, const uint8_t* s_pub, const uint8_t* pub_w,
                   const uint8_t* c, const uint8_t* es, const uint8_t* XR, size_t n_norm, size_t n_lin,
                   const uint8_t* fw, const uint8_t* fl, size_t n_init, const uint8_t* init_s, const uint8_t* init_p,
                   int* ok);

int bppp_nl_verify_gens(bppp_gens* gens, int kind, size_t batch, size_t k, const uint8_t* q, const uint8_t* s_pub,
                        const uint8_t* pub_w, const uint8_t* c, const uint8_t* es, const uint8_t* XR, size_t n_norm,
                        size_t n_lin, const uint8_t* fw, const uint8_t* fl, size_t n_init, const uint8_t* init_s,
                        const uint8_t* init_p, int* ok);

/* ---- TypedReciprocal scalar phases on the device.
 * Replaces the per-entry ("norm" part, one slot per Phase-1 entry) arithmetic of proveTRRPM,
 * src/RangeProof/TypedReciprocal.hs:399-444: makePhase2s' reciprocals r_i = ps_i/(e+d_i) and
 * c_i (:171-195), makeErrorTerms (:217-233), makePublicConsts (:236-263) and the witness
 * combination pub + bl + t m + t^2 dm + t^3 r (:439).  The transcript, the blinders and the few
 * scalar/"linear" slots stay with the caller.  Static per-entry description (from the schema):
 *   ent_desc[i] = {u32 flags, i32 range index, i32 shared-base index (x^(3+2j), bases sorted), i32 0}
 *   flags: 1 typing entry, 2 output, 4 assumed, 8 inline digit, 16 symbol s_i != 0
 *   ent_b[i] = digit coefficient b_i, ent_s[i] = symbol s_i        (32-byte canonical scalars)
 * Calls must follow phase1 -> phase2 -> phase3 -> commit_bl -> phase4 -> bppp_nl_create_trrp;
 * n_entries must equal the generator set's N.  All scalars are 32-byte little-endian canonical. */
typedef struct bppp_trrp bppp_trrp;
int bppp_trrp_create(bppp_gens* gens, size_t n_entries, const uint8_t* ent_desc, const uint8_t* ent_b,
                     const uint8_t* ent_s, size_t n_ranges, size_t n_bases, bppp_trrp** out);
void bppp_trrp_destroy(bppp_trrp* h);
/* phase 1 (:399-410): sc_dm_m = [batch][2][1+N+M] commitment scalars of the digit and multiplicity
 * witnesses (scalar slot, norm slots, linear slots); amounts = [batch][n_ranges] committed values.
 * coms = [batch][2] points (dmCom, mCom). */
int bppp_trrp_phase1(bppp_trrp* h, size_t batch, const uint8_t* sc_dm_m, const uint8_t* amounts, uint8_t* coms);
/* phase 2 (:412-419): chal = [batch][4] = (e, 1/e, x, 1/r0); r_sclin = [batch][1+M] scalar slot and
 * linear slots of the reciprocal witness (blindErrWitness, Internal.hs:142-152) with the err7 slot
 * zero; err7_slot = its index in the linear part.  rcom = [batch] points, err7 = [batch] scalars. */
int bppp_trrp_phase2(bppp_trrp* h, const uint8_t* chal, const uint8_t* r_sclin, size_t err7_slot, uint8_t* rcom,
                     uint8_t* err7);
/* phase 3a (:421-433): chal = [batch][2] = (q-power base, x'); bls_nrm = [batch][N] norm-part
 * blinders.  errs = [batch][6] error terms summed over the norm entries (the caller adds the
 * shared-multiplicity term 2 sum cs_i bls_i to errs[3]). */
int bppp_trrp_phase3(bppp_trrp* h, const uint8_t* chal, const uint8_t* bls_nrm, uint8_t* errs);
/* phase 3b (:434): bl_sclin = [batch][1+M] scalar and linear slots of the blinding witness. */
int bppp_trrp_commit_bl(bppp_trrp* h, const uint8_t* bl_sclin, uint8_t* blcom);
/* The same phases with the Fiat-Shamir transcript on the device: after
 * bppp_trrp_set_transcript(h, show_format >= 0) the commitments are absorbed where they are produced
 * and each call also returns the challenges of the reference's next `oracle'` call --
 *   phase1_tr: n_coms = [batch][n_inputs] input commitments; chal = [batch][3] (e, x, r0)   (:411)
 *   phase2_tr: chal_out = [batch][3] (q, x', r1)                                            (:420)
 *   commit_bl_tr: chal_out = [batch] (t)                                                    (:435)
 * and bppp_nl_create_trrp hands the transcript on to the argument (bppp_nl_round_challenge).
 * phase3_rnd draws the N norm blinders on the device: `random` values n0[b] .. n0[b]+N-1 of seeds[b]. */
int bppp_trrp_set_transcript(bppp_trrp* h, int show_format);
/* one-shot: the next _tr call also writes 1 / challenge for each of its challenges to `out` (same layout) */
int bppp_trrp_want_inverses(bppp_trrp* h, uint8_t* out);
/* makeSharedCoeffs (TypedReciprocal.hs:204-206) on the device.  Once per setup: slot i belongs to shared base number
 * base_idx[i] (index into the sorted base list) and symbol sym[i].  Then, after bppp_trrp_phase2* (prover) or
 * bppp_trrp_verify_pub (verifier): out[b][i] = x^(3 + 2 base_idx[i]) * (1/e - 1/(e + sym[i])); montgomery != 0
 * returns residues times 2^256 mod r (the library's internal form) instead of canonical scalars. */
int bppp_trrp_set_shared(bppp_trrp* h, size_t n_slots, const int32_t* base_idx, const uint8_t* sym);
int bppp_trrp_shared_coeffs(bppp_trrp* h, int montgomery, uint8_t* out);
int bppp_trrp_phase1_tr(bppp_trrp* h, size_t batch, const uint8_t* sc_dm_m, const uint8_t* amounts, size_t n_inputs,
                        const uint8_t* n_coms, uint8_t* coms, uint8_t* chal);
int bppp_trrp_phase2_tr(bppp_trrp* h, const uint8_t* chal, const uint8_t* r_sclin, size_t err7_slot, uint8_t* rcom,
                        uint8_t* err7, uint8_t* chal_out);
int bppp_trrp_phase3_rnd(bppp_trrp* h, const uint8_t* chal, const char* const* seeds, const uint64_t* n0, uint8_t* errs);
int bppp_trrp_commit_bl_tr(bppp_trrp* h, const uint8_t* bl_sclin, uint8_t* blcom, uint8_t* chal_out);
/* phase 4 (:435-444): chal = [batch][2] = (t, 1/q0).  sums = [batch][3] = (sum q2_i p_i^2 over all
 * entries, sum q2_i and sum v_i over the digit entries); the combined witness stays on the device. */
int bppp_trrp_phase4(bppp_trrp* h, const uint8_t* chal, uint8_t* sums);
/* verifier (verifyTRRPM, :447-467): chal = [batch][8] = (e, 1/e, x, x', q-power base q0, 1/q0, t, 0).
 * sums as in phase 4; the norm part of makePublicConsts stays on the device for bppp_nl_verify_trrp,
 * which is bppp_nl_verify_gens (norm-linear argument) without the pub_w argument. */
int bppp_trrp_verify_pub(bppp_trrp* h, size_t batch, const uint8_t* chal, uint8_t* sums);
/* Batch verification across proofs (the reference's TODO, src/RangeProof/TypedReciprocal.hs:469-472,
 * src/RangeProof.hs:103-106): bppp_nl_verify_trrp / bppp_nl_verify_gens plus weights = [batch] scalars drawn at random
 * by the caller after seeing the proofs.  sum_b weight_b * (check of proof b) is ONE fixed-base MSM over the shared
 * generators plus ONE Pippenger over all per-proof points; if it is the identity every ok[b] = 1, otherwise the
 * per-proof checks run and locate the bad proofs -- the verdicts are exact either way. */
int bppp_nl_verify_trrp_rlc(bppp_trrp* h, size_t k, const uint8_t* q, const uint8_t* s_pub, const uint8_t* c,
                            const uint8_t* es, const uint8_t* XR, size_t n_norm, size_t n_lin, const uint8_t* fw,
                            const uint8_t* fl, size_t n_init, const uint8_t* init_s, const uint8_t* init_p,
                            const uint8_t* weights, int* ok);
int bppp_nl_verify_gens_rlc(bppp_gens* gens, int kind, size_t batch, size_t k, const uint8_t* q, const uint8_t* s_pub,
                            const uint8_t* pub_w, const uint8_t* c, const uint8_t* es, const uint8_t* XR,
                            size_t n_norm, size_t n_lin, const uint8_t* fw, const uint8_t* fl, size_t n_init,
                            const uint8_t* init_s, const uint8_t* init_p, const uint8_t* weights, int* ok);
int bppp_nl_verify_trrp(bppp_trrp* h, size_t k, const uint8_t* q, const uint8_t* s_pub, const uint8_t* c,
                        const uint8_t* es, const uint8_t* XR, size_t n_norm, size_t n_lin, const uint8_t* fw,
                        const uint8_t* fl, size_t n_init, const uint8_t* init_s, const uint8_t* init_p, int* ok);
/* the norm-linear argument (bppp_nl_create_gens) over the device-resident witness of phase 4 */
int bppp_nl_create_trrp(bppp_trrp* h, const uint8_t* q, const uint8_t* s, const uint8_t* l, const uint8_t* c,
                        bppp_nl** out);

/* ---- Range-proof layer (host C++ above the device entry points; the Fiat-Shamir transcript,
 * round sequencing and the scalar phases run on host threads, every group operation on the GPU).
 * Mirrors RPOpening / RangeProof (src/RangeProof.hs:25-101) for TypedReciprocal
 * (src/RangeProof/TypedReciprocal.hs) and Binary (src/RangeProof/Binary.hs), batched. */
typedef struct bppp_rp bppp_rp;
typedef struct {
    uint8_t min[16], max[16];          /* two's-complement little-endian 128-bit integers */
    uint32_t base;                     /* ignored for binary proofs */
    int32_t is_shared, is_output, is_assumed;
} bppp_range_spec;                     /* one entry per range AFTER `count` expansion (app/Parse.hs:126-165) */
typedef struct {
    uint8_t amount[16], type[16];
    int32_t is_output;
} bppp_public_spec;                    /* app/Parse.hs:210-235 */
enum { BPPP_SHOW_PREFIXED_P = 0, BPPP_SHOW_BARE_DECIMAL = 1 };
enum { BPPP_ROOT_EXP = 0, BPPP_ROOT_EVEN = 1, BPPP_ROOT_SMALLER = 2 };
/* TRRP.setup (TypedReciprocal.hs:332-359) / setupBRP (Binary.hs:143-156); generators from
 * getPoints basis_seed (app/Main.hs:68-72). */
int bppp_rp_setup(bppp_ctx* ctx, int binary, int arg_kind, int typed_or_conserved, const char* basis_seed,
                  int show_format, int root_policy, size_t n_ranges, const bppp_range_spec* ranges, size_t n_pub,
                  const bppp_public_spec* pubs, bppp_rp** out);
void bppp_rp_free(bppp_rp* s);
const char* bppp_rp_last_error(bppp_rp* s);
int bppp_rp_info(bppp_rp* s, size_t* n_inputs, size_t* num_rp_coms, size_t* nrm_len, size_t* lin_len, size_t* rounds,
                 size_t* fin_norm, size_t* fin_lin);
int bppp_rp_points(bppp_rp* s, size_t count, uint8_t* out);
int bppp_input_blind(const char* random_seed, uint64_t j, uint8_t out[32]);
void bppp_set_host_threads(int n);
/* where bppp_rp_prove_batch / bppp_rp_verify_batch run the Fiat-Shamir transcript: 0 = host (default, the
 * reference's arrangement), 1 = device (bit-identical proofs and verdicts; needs the device scalar
 * phases, i.e. TypedReciprocal over the norm-linear argument).  Environment default: BPPP_DEVICE_TRANSCRIPT=1. */
int bppp_rp_set_device_transcript(bppp_rp* s, int on);
/* bppp_rp_verify_batch by one random linear combination per lane sub-batch (weights from the OS entropy source);
 * falls back to the per-proof checks when a combination fails, so ok[] is exact.  Default off (BPPP_BATCH_VERIFY=1). */
int bppp_rp_set_batch_verify(bppp_rp* s, int on);
/* bppp_gens_enable_lut for the setup's generator list (all lanes share one table).  Environment default: BPPP_LUT_GB. */
int bppp_rp_enable_lut(bppp_rp* s, double budget_gb, int* c_out);
/* RangeProof.proveM for `batch` independent proofs (see rp_host.cpp for the buffer layout) */
int bppp_rp_prove_batch(bppp_rp* s, size_t batch, const uint8_t* values, const uint8_t* types, const uint8_t* blinds,
                        const char* const* random_seeds, uint8_t* coms, uint8_t* responses, uint8_t* finals);
/* RangeProof.verifyM for `batch` proofs */
int bppp_rp_verify_batch(bppp_rp* s, size_t batch, size_t rounds, size_t n_norm, size_t n_lin, const uint8_t* coms,
                         const uint8_t* responses, const uint8_t* finals, int* ok);
/* wire format: encodeProof' / decodeProof' (src/RangeProof.hs:60-85) over encodeScalarsCurvePoints /
 * decodeCommitments (src/Encoding.hs:75-134): proof.bin and commits.bin images per proof */
int bppp_rp_encoded_sizes(bppp_rp* s, size_t* proof_bytes, size_t* commits_bytes);
int bppp_rp_encode_batch(bppp_rp* s, size_t batch, const uint8_t* coms, const uint8_t* responses, const uint8_t* finals,
                         uint8_t* proof_bin, uint8_t* commits_bin);
int bppp_rp_decode_batch(bppp_rp* s, size_t batch, const uint8_t* proof_bin, const uint8_t* commits_bin, uint8_t* coms,
                         uint8_t* responses, uint8_t* finals, int* ok);
/* contexts of the concurrent lanes a setup runs its sub-batches on (BPPP_LANES, default 8) */
int bppp_rp_contexts(bppp_rp* s, bppp_ctx** out, size_t cap, size_t* count);
/* host-only self-test hooks (no device needed) */
int bppp_host_sha256(const uint8_t* data, size_t n, uint8_t out[32]);
int bppp_host_oracle(const uint8_t* pts, size_t npts, int count, int show_format, uint8_t* out);
int bppp_host_fr(int op, const uint8_t* a, const uint8_t* b, uint8_t* out);
/* the host job scheduler under concurrent submitters: 0 when every item of every job ran exactly once */
int bppp_host_scheduler_selftest(int lanes, int jobs, int items);
/* the batched transcript paths (two-stream SHA-256): n_random values of `random`, commitments `pts`,
 * then `rounds` (X,R) pairs (128 B each); out = randoms | round challenges | the same via the paired path */
int bppp_host_transcript(const char* seed, int show_format, size_t n_random, const uint8_t* pts, size_t npts,
                         const uint8_t* round_pts, size_t rounds, uint8_t* out);
int bppp_host_get_points(const char* seed, size_t count, int root_policy, uint8_t* out);

/* ---- debug / self-test entry points used by tests/ (element-wise device arithmetic) */
int bppp_dbg_field(bppp_ctx* ctx, int op, size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out);
int bppp_dbg_ec(bppp_ctx* ctx, int op, size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out /* n*64 affine */);

#ifdef __cplusplus
}
#endif
#endif
