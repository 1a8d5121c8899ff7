// C ABI (include/bppp_b200.h): context, device buffers, kernel launches and the host-side round
// sequencing of the NormLinear argument.  No CPU fallback: every compute entry point launches
// the sm_100a kernels of kernels.cuh.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <dlfcn.h>
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bppp_b200.h"
#include "host_math.hpp"
#include "host/fr64.hpp"
#include <atomic>
#include <thread>
#include "kernels.cuh"
#include "pippenger.cuh"
#include "transcript.cuh"
#include "rounds.cuh"
#include "lut.cuh"

using namespace bppp;

static thread_local cudaStream_t g_alloc_stream = nullptr;   // set at every API entry (ENTER)

enum KernelId { K_FR_CONVERT = 0, K_FOLD_DOTS, K_DOTS_FINISH, K_MSM_SCALARS, K_PAIR_FOLD, K_TO_AFFINE, K_MSM_BUCKET,
                K_MSM_FINISH, K_TENSOR, K_FB_BUILD, K_FB_MSM, K_BCAST, K_DBG, K_MSM_GENS, K_JAC_SUM, K_GT_BUILD, K_EXPAND, K_COEF, K_IP_MISC, K_POW_TABLE, K_TRRP, K_MSM_GROUPS, K_MSM_REDUCE, K_CHECK, K_PIP_SORT, K_PIP_ACCUM, K_PIP_MERGE, K_PIP_REDUCE, K_PIP_HORNER, K_TR_POINTS, K_TR_RENDER, K_TR_SQUEEZE, K_TR_RANDOM, K_ROUND_STATE, K_BATCH_WEIGHT, K_MSM_LUT, K_LUT_BUILD, K_COUNT };
static const char* const kKernelNames[K_COUNT] = {"k_fr_convert", "k_fold_dots", "k_dots_finish", "k_msm_scalars",
                                                   "k_pair_fold", "k_batch_to_affine", "k_msm_bucket", "k_msm_finish",
                                                   "k_tensor_expand", "k_fb_build", "k_fb_msm", "k_bcast_point", "k_dbg", "k_msm_gens",
                                                   "k_jac_sum", "k_gt_build", "k_expand_scalars", "k_coef_update", "k_ip_misc", "k_pow_table", "k_trrp_phases", "k_msm_gens_small", "k_msm_gens_reduce", "k_check_points",
                                                   "k_pip_sort", "k_pip_accum", "k_pip_merge", "k_pip_reduce", "k_pip_horner",
                                                   "k_hash_to_curve", "k_tr_render", "k_tr_squeeze", "k_tr_random", "k_round_state", "k_batch_weight", "k_msm_lut", "k_lut_build"};
struct ProfRec {
    int id;
    double work;                 // algorithmic units of this launch: IMADs or bytes
    cudaEvent_t a, b;
};
struct bppp_ctx {
    int dev = 0;
    cudaStream_t st = nullptr;
    std::string err;
    uint64_t launches = 0;
    uint64_t h2d = 0, d2h = 0;
    bool prof = false;
    std::vector<ProfRec> pending;
    std::vector<cudaEvent_t> pool;
    double k_ms[K_COUNT] = {0}, k_work[K_COUNT] = {0};
    double k_top_ms[K_COUNT] = {0}, k_top_work[K_COUNT] = {0};   // the single launch with the most algorithmic work
    uint64_t k_n[K_COUNT] = {0};
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    cudaEvent_t sync_ev = nullptr;   // blocking-sync event: waiting host threads sleep whatever the device's schedule flags
    unsigned long long* lut_count = nullptr;   // device counter: table lookups (= mixed additions) of k_msm_lut while profiling
};
// Wait for everything queued on the context's stream.  An event created with cudaEventBlockingSync
// makes the calling thread sleep even when the primary context was created by someone else (torch,
// the embedding application) without cudaDeviceScheduleBlockingSync: the lanes' driver threads
// must not spin on the cores the host phases need.
static cudaError_t ctx_sync(bppp_ctx* c) {
    if (!c->sync_ev) return cudaStreamSynchronize(c->st);
    cudaError_t e = cudaEventRecord(c->sync_ev, c->st);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(c->sync_ev);
}
static cudaEvent_t prof_event(bppp_ctx* c) {
    cudaEvent_t e;
    if (!c->pool.empty()) { e = c->pool.back(); c->pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {               // wraps one kernel launch with a pair of events on the launching stream
    bppp_ctx* c;
    ProfScope(bppp_ctx* ctx, int id, double work) : c(ctx) {
        c->launches++;
        if (!c->prof) return;
        ProfRec r;
        r.id = id; r.work = work; r.a = prof_event(c); r.b = prof_event(c);
        cudaEventRecord(r.a, c->st);
        c->pending.push_back(r);
    }
    ~ProfScope() {
        if (c->prof) cudaEventRecord(c->pending.back().b, c->st);
    }
};
static void prof_collect(bppp_ctx* c) {
    if (c->pending.empty()) return;
    cudaStreamSynchronize(c->st);
    for (auto& r : c->pending) {
        float ms = 0;
        cudaEventElapsedTime(&ms, r.a, r.b);
        c->k_ms[r.id] += ms; c->k_work[r.id] += r.work; c->k_n[r.id]++;
        if (r.work > c->k_top_work[r.id]) { c->k_top_work[r.id] = r.work; c->k_top_ms[r.id] = ms; }
        c->pool.push_back(r.a); c->pool.push_back(r.b);
    }
    c->pending.clear();
}
#define H2D(dst, src, bytes) (ctx->h2d += (bytes), cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->st))
#define D2H(dst, src, bytes) (ctx->d2h += (bytes), cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->st))

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            char buf_[256];                                                                     \
            snprintf(buf_, sizeof buf_, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            ctx->err = buf_;                                                                    \
            return BPPP_ERR_CUDA;                                                               \
        }                                                                                       \
    } while (0)
#define ENTER(c)                          \
    do {                                  \
        CK(cudaSetDevice((c)->dev));      \
        g_alloc_stream = (c)->st;         \
    } while (0)
#define FAIL(code, msg)  \
    do {                 \
        ctx->err = msg;  \
        return code;     \
    } while (0)


// algorithmic work per launch for the rooflines: IMADs (32x32->64 multiply-accumulates)
// for the group-law kernels, bytes for the scalar fold.
static double msm_alg_imads(double n, double bits = 256.0) {   // ceil(bits/c*) (n + 2^(c*-1)) * 11 * 136
    if (n < 1) return 0;
    int c = (int)floor(log2(n)) - 2;
    if (c < 4) c = 4;
    return ceil(bits / c) * (n + pow(2.0, c - 1)) * 11.0 * 136.0;
}
// one folded generator = 3.9k Fq-mul * 136 = 5.3e5 IMAD with 256-bit scalars; the
// half-length (a, b) of rationalReduceScalar halve it
static double fold_alg_imads(double n_points) { return n_points * 0.5 * 3.9e3 * 136.0; }
#define WORK_K_MSM_BUCKET g_work
#define WORK_K_MSM_FINISH 0
#define WORK_K_BATCH_TO_AFFINE 0
#define WORK_K_FB_BUILD 0
#define WORK_K_FB_MSM 0
#define WORK_K_PAIR_FOLD g_work
#define WORK_K_FOLD_DOTS g_work
#define WORK_K_BCAST_POINT 0
#define WORK_K_FR_CONVERT 0
#define WORK_K_DOTS_FINISH 0
#define WORK_K_MSM_SCALARS 0
#define WORK_K_TENSOR_EXPAND 0
#define WORK_K_DBG_FIELD 0
#define WORK_K_DBG_EC 0
static thread_local double g_work = 0;           // set by the caller right before a launch

int g_bppp_tune_device = 0;     // set by bppp_tune_process (rp_host.cpp)
int g_capi_threads = 0;
static thread_local int t_capi_threads = 0;
extern "C" void bppp_set_device_host_threads(int n) { g_capi_threads = n; }
extern "C" void bppp_set_thread_host_threads(int n) { t_capi_threads = n; }
extern "C" int bppp_ctx_device(bppp_ctx* ctx);
template <class F>
static void host_parallel_for(size_t n, F fn) {
    int nt = t_capi_threads > 0 ? t_capi_threads : (g_capi_threads > 0 ? g_capi_threads : (int)std::thread::hardware_concurrency());
    if (nt < 1) nt = 1;
    if (nt > 4) nt = 4;                      // short loops: a few threads suffice, lanes run concurrently
    if ((size_t)nt > n / 16) nt = (int)(n / 16);
    if (nt <= 1) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::atomic<size_t> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&]() {
            for (;;) {
                size_t i0 = next.fetch_add(16);
                if (i0 >= n) break;
                size_t i1 = i0 + 16 < n ? i0 + 16 : n;
                for (size_t i = i0; i < i1; i++) fn(i);
            }
        });
    for (auto& t : th) t.join();
}

namespace {

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    DBuf() {}
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { release(); }
    cudaStream_t st = nullptr;
    void release() {
        if (p) cudaFreeAsync(p, st);                 // stream-ordered: no device-wide sync
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        st = g_alloc_stream;
        cudaError_t e = cudaMallocAsync((void**)&p, count * sizeof(T), st);
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t ensure(size_t count) { return count <= n ? cudaSuccess : alloc(count); }
};

__global__ void k_bcast_point(const Affine* src, Affine* dst, size_t stride, size_t batch) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < batch) st_aff(dst + i * stride, ld_aff(src));
}

bool check_fr(const uint8_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) {
        uint64_t top;
        memcpy(&top, b + 32 * i + 24, 8);
        if (top != ~0ULL) continue;                  // the top word of r is all ones: anything below is canonical
        if (!host::fr_is_canonical(host::from_bytes(b + 32 * i))) return false;
    }
    return true;
}
bool check_fq(const uint8_t* b, size_t n) {
    for (size_t i = 0; i < n; i++)
        if (!host::fq_is_canonical(host::from_bytes(b + 32 * i))) return false;
    return true;
}

// ----------------------------------------------------------------------------- size-aware Pippenger (pippenger.cuh)
struct PipWork {                     // scratch of one call; stream-ordered pool allocations, reused across rounds
    DBuf<unsigned> off, cur, ent, bsum, heavy;
    DBuf<u256> dec;
    DBuf<unsigned char> dsgn;
    DBuf<Xyzz> bucket, slots, seg;
    DBuf<Jac> win;
};
// window width for n scalars = 2n half-length terms (GLV): a window has about 2n/32 buckets.  Widths whose
// top window would hold nothing but the carry (c = 4, 8, 16: (W-1)*c = 128) are avoided.
int pip_window_bits(size_t n) {
    int l = 0;
    while (((2 * n) >> l) > 1) l++;
    int c = l - 4;
    if (c < 5) c = 5;
    if (c > 15) c = 15;
    if (c == 8) c = 9;
    return c;
}
// `batch` x `n_out` MSMs of n terms: bases pts[p*pts_stride + i] (stride 0: shared), canonical scalars
// sc[p*sc_stride + o*sc_out_stride + i]; Jacobian results in d_res[p*n_out + o].
int run_msm_pip(bppp_ctx* ctx, PipWork& wk, const Affine* pts, size_t pts_stride, const u256* sc, size_t sc_stride,
                size_t sc_out_stride, size_t n, size_t batch, int n_out, Jac* d_res, double work_per_proof) {
    if (n == 0 || batch == 0 || n_out <= 0) FAIL(BPPP_ERR_ARG, "empty MSM");
    if (n >= ((size_t)1 << 30)) FAIL(BPPP_ERR_ARG, "MSM too long");
    const int c = pip_window_bits(n), W = (129 + c - 1) / c, NBK = 1 << (c - 1);
    const int SEG = 1 << std::max(0, (c - 1) / 2 - 2), NS = NBK / SEG;
    // proofs per pass: bucket ids and entry positions are 32-bit, and the entry list is bounded to 2 GiB
    const size_t ent_per_proof = (size_t)n_out * 2 * n * W, nt_per_proof = (size_t)n_out * W * NBK;
    size_t slab = std::min<size_t>(batch, std::max<size_t>(1, ((size_t)1 << 29) / std::max(ent_per_proof, nt_per_proof)));
    if (std::max(ent_per_proof, nt_per_proof) >= ((size_t)1 << 31)) FAIL(BPPP_ERR_ARG, "MSM too long for one pass");
    for (size_t p0 = 0; p0 < batch; p0 += slab) {
        const size_t nb = std::min(slab, batch - p0);
        PipArgs A;
        A.pts = pts + p0 * pts_stride; A.pts_stride = pts_stride;
        A.sc = sc + p0 * sc_stride; A.sc_stride = sc_stride; A.sc_out_stride = sc_out_stride;
        A.n = (unsigned)n; A.n_out = (unsigned)n_out; A.n_prob = (unsigned)(nb * n_out);
        A.c = c; A.W = W; A.NBK = NBK; A.NT = (unsigned)(nb * nt_per_proof);
        A.SEG = SEG; A.NS = NS;
        const size_t e_max = nb * ent_per_proof;
        // entries per accumulation thread: enough threads for a few waves on 148 SMs, at least 4 entries each
        size_t L = e_max / ((size_t)148 * 1024);
        L = std::min<size_t>(64, std::max<size_t>(4, L));
        A.L = (int)L;
        const size_t n_thr = (e_max + L - 1) / L;
        const unsigned n_scan = (unsigned)(((size_t)A.NT + PIP_SCAN_BLOCK - 1) / PIP_SCAN_BLOCK);
        const size_t n_seg = (size_t)A.n_prob * W * NS, n_win = (size_t)A.n_prob * W;
        CK(wk.off.ensure((size_t)A.NT + 1)); CK(wk.cur.ensure(A.NT)); CK(wk.ent.ensure(e_max)); CK(wk.bsum.ensure(n_scan));
        CK(wk.heavy.ensure((size_t)A.NT + 1));
        CK(wk.dec.ensure((size_t)A.n_prob * n)); CK(wk.dsgn.ensure((size_t)A.n_prob * n));
        A.dec = wk.dec.p; A.dsgn = wk.dsgn.p;
        CK(wk.bucket.ensure(A.NT)); CK(wk.slots.ensure(2 * n_thr)); CK(wk.seg.ensure(2 * n_seg)); CK(wk.win.ensure(n_win));
        A.off = wk.off.p; A.cur = wk.cur.p; A.ent = wk.ent.p; A.bsum = wk.bsum.p; A.heavy = wk.heavy.p;
        A.bucket = wk.bucket.p; A.slotF = wk.slots.p; A.slotL = wk.slots.p + n_thr;
        A.segS = wk.seg.p; A.segT = wk.seg.p + n_seg; A.win = wk.win.p; A.out = d_res + p0 * n_out;
        CK(cudaMemsetAsync(A.off, 0, ((size_t)A.NT + 1) * sizeof(unsigned), ctx->st));
        CK(cudaMemsetAsync(A.heavy, 0, sizeof(unsigned), ctx->st));
        const unsigned g_terms = (unsigned)(((size_t)A.n_prob * n + 255) / 256);
        {   ProfScope ps_(ctx, K_PIP_SORT, 0);
            k_pip_glv<<<g_terms, 256, 0, ctx->st>>>(A);
            k_pip_count<<<g_terms, 256, 0, ctx->st>>>(A);
            k_pip_scan_blocks<<<n_scan, 256, 0, ctx->st>>>(A);
            k_pip_scan_top<<<1, 256, 0, ctx->st>>>(A, n_scan);
            k_pip_scan_apply<<<n_scan, 256, 0, ctx->st>>>(A);
            k_pip_scatter<<<g_terms, 256, 0, ctx->st>>>(A);
            ctx->launches += 5;
        }
        CK(cudaGetLastError());
        {   ProfScope ps_(ctx, K_PIP_ACCUM, work_per_proof * (double)nb);
            k_pip_accum<<<(unsigned)((n_thr + 127) / 128), 128, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
        {   ProfScope ps_(ctx, K_PIP_MERGE, 0);
            k_pip_merge<<<(A.NT + 127) / 128, 128, 0, ctx->st>>>(A);
            k_pip_merge_heavy<<<148, PIP_HEAVY_THREADS, 0, ctx->st>>>(A);
            ctx->launches += 1;
        }
        CK(cudaGetLastError());
        {   ProfScope ps_(ctx, K_PIP_REDUCE, 0);
            k_pip_reduce1<<<(unsigned)((n_seg + 127) / 128), 128, 0, ctx->st>>>(A);
            k_pip_reduce2<<<(unsigned)((n_win + 3) / 4), 128, 0, ctx->st>>>(A);
            ctx->launches += 1;
        }
        CK(cudaGetLastError());
        {   ProfScope ps_(ctx, K_PIP_HORNER, 0);
            k_pip_horner<<<(A.n_prob + 31) / 32, 32, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    }
    return BPPP_OK;
}

// ----------------------------------------------------------------------------- MSM driver
struct MsmPlan {
    std::vector<MsmSlice> slices;
    DBuf<MsmSlice> d_slices;
    DBuf<Jac> d_partial;
    size_t smem = 0;
    PipWork pip;
    MsmSlice whole;                   // the undivided MSM (one add() per plan), for the size-aware path
    size_t whole_n = 0;
    void add(const Affine* pts, size_t pts_stride, const u256* sc, size_t sc_stride, size_t sc_out_stride, size_t n) {
        whole.pts = pts; whole.pts_stride = pts_stride; whole.sc = sc; whole.sc_stride = sc_stride;
        whole.sc_out_stride = sc_out_stride; whole.n = 0;
        whole_n = n;
        for (size_t off = 0; off < n; off += MSM_MAX_CHUNK) {
            MsmSlice s;
            s.pts = pts + off; s.pts_stride = pts_stride;
            s.sc = sc + off; s.sc_stride = sc_stride; s.sc_out_stride = sc_out_stride;
            s.n = (int)std::min<size_t>(MSM_MAX_CHUNK, n - off);
            slices.push_back(s);
        }
    }
};
size_t msm_smem_bytes(int max_n) {
    return (size_t)(2 * MSM_W * MSM_NB + 1) * 4 + (size_t)max_n * MSM_W * 2 + 16;
}
bool g_attr_set = false;

// runs the MSMs described by plan.slices for `batch` proofs x `n_out` outputs; result Jacobian
// points in d_res[(p*n_out + o)]
int run_msm(bppp_ctx* ctx, MsmPlan& plan, size_t batch, int n_out, Jac* d_res, double work_per_proof = 0) {
    int nch = (int)plan.slices.size();
    if (nch == 0) FAIL(BPPP_ERR_ARG, "empty MSM");
    // Few long MSMs (one large argument, bppp_msm): the size-aware global-memory Pippenger.  Many short ones
    // (a batch of small proofs, <= 2048 terms each): one CTA per MSM with shared-memory bucket lists below.
    if (nch > 1 || batch * (size_t)n_out <= 16)
        return run_msm_pip(ctx, plan.pip, plan.whole.pts, plan.whole.pts_stride, plan.whole.sc, plan.whole.sc_stride,
                           plan.whole.sc_out_stride, plan.whole_n, batch, n_out, d_res, work_per_proof);
    int max_n = 0;
    for (auto& s : plan.slices) max_n = std::max(max_n, s.n);
    CK(plan.d_slices.ensure(nch));
    CK(H2D(plan.d_slices.p, plan.slices.data(), nch * sizeof(MsmSlice)));
    CK(plan.d_partial.ensure(batch * n_out * nch * MSM_W));
    if (!g_attr_set) {
        CK(cudaFuncSetAttribute(k_msm_bucket, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)msm_smem_bytes(MSM_MAX_CHUNK)));
        CK(cudaFuncSetAttribute(k_msm_bucket, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        g_attr_set = true;
    }
    MsmArgs A;
    A.slices = plan.d_slices.p; A.n_chunks = nch; A.partial = plan.d_partial.p; A.n_out = n_out;
    // grid.z is limited to 65535: loop over batch slabs
    for (size_t b0 = 0; b0 < batch; b0 += 32768) {
        size_t nb = std::min<size_t>(32768, batch - b0);
        MsmArgs B = A;
        g_work = work_per_proof * (double)nb;
        // shift per-proof bases by b0 proofs: done by offsetting the partial pointer and using a
        // slab-local copy of the slices when b0 > 0
        if (b0 > 0) {
            std::vector<MsmSlice> sl = plan.slices;
            for (auto& s : sl) { s.pts += b0 * s.pts_stride; s.sc += b0 * s.sc_stride; }
            MsmSlice* d2;
            CK(cudaMallocAsync((void**)&d2, nch * sizeof(MsmSlice), ctx->st));
            CK(H2D(d2, sl.data(), nch * sizeof(MsmSlice)));
            CK(ctx_sync(ctx));
            B.slices = d2;
            B.partial = A.partial + b0 * n_out * nch * MSM_W;
            { ProfScope ps_(ctx, K_MSM_BUCKET, WORK_K_MSM_BUCKET);
            k_msm_bucket<<<dim3(nch, n_out, (unsigned)nb), MSM_THREADS, msm_smem_bytes(max_n), ctx->st>>>(B);
            }
            CK(cudaGetLastError());
            CK(cudaFreeAsync(d2, ctx->st));
        } else {
            { ProfScope ps_(ctx, K_MSM_BUCKET, WORK_K_MSM_BUCKET);
            k_msm_bucket<<<dim3(nch, n_out, (unsigned)nb), MSM_THREADS, msm_smem_bytes(max_n), ctx->st>>>(B);
            }
            CK(cudaGetLastError());
        }
    }
    size_t n_msm = batch * n_out;
    { ProfScope ps_(ctx, K_MSM_FINISH, WORK_K_MSM_FINISH);
    k_msm_finish<<<(unsigned)n_msm, 32, 0, ctx->st>>>(plan.d_partial.p, nch, d_res, n_msm);
    }
    CK(cudaGetLastError());
    return BPPP_OK;
}

int to_affine(bppp_ctx* ctx, const Jac* in, size_t in_stride, Affine* out, size_t out_stride, int out_off, int n_per,
              size_t total) {
    if (total == 0) return BPPP_OK;
    // enough threads to fill the chip, at most 32 points per thread
    int chunk = (int)std::min<size_t>(32, std::max<size_t>(1, total / (148 * 512)));
    size_t threads = (total + chunk - 1) / chunk;
    { ProfScope ps_(ctx, K_TO_AFFINE, WORK_K_BATCH_TO_AFFINE);
    k_batch_to_affine<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->st>>>(in, in_stride, out, out_stride, out_off,
                                                                              n_per, total, chunk);
    }
    CK(cudaGetLastError());
    return BPPP_OK;
}


// every point of pts[0..n) on the curve (or the identity)?  `flags` gets one int per group of `per`
// consecutive points (non-zero = an off-curve point in the group); no synchronisation here.
int check_points_async(bppp_ctx* ctx, const Affine* pts, size_t n, size_t per, DBuf<int>& flags) {
    const size_t groups = (n + per - 1) / per;
    CK(flags.alloc(std::max<size_t>(groups, 1)));
    CK(cudaMemsetAsync(flags.p, 0, std::max<size_t>(groups, 1) * sizeof(int), ctx->st));
    if (n == 0) return BPPP_OK;
    { ProfScope ps_(ctx, K_CHECK, 0);
    k_check_points<<<(unsigned)((n + 127) / 128), 128, 0, ctx->st>>>(pts, n, per, flags.p);
    }
    CK(cudaGetLastError());
    return BPPP_OK;
}
// the same for one group, synchronously: BPPP_ERR_RANGE when a point is off the curve
int check_points_sync(bppp_ctx* ctx, const Affine* pts, size_t n, const char* what) {
    DBuf<int> flags;
    int rc = check_points_async(ctx, pts, n, std::max<size_t>(n, 1), flags);
    if (rc) return rc;
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CK(ctx_sync(ctx));
    if (bad) FAIL(BPPP_ERR_RANGE, std::string(what) + ": point not on the curve");
    return BPPP_OK;
}

}  // namespace

// =============================================================================== context
extern "C" int bppp_abi_version(void) { return 1; }

extern "C" int bppp_init(int device, bppp_ctx** out) {
    if (!out) return BPPP_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return BPPP_ERR_CUDA;
    // bppp_tune_process(BPPP_TUNE_DEVICE) only: host threads waiting for the device sleep instead of
    // spinning (the cores are needed by the host phases of the other lanes; ignored if the primary
    // context is already active).  Without it the process-wide device flags are left alone; waits still
    // sleep on the context's blocking-sync event.
    if (g_bppp_tune_device) {
        cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync);
        cudaGetLastError();
    }
    if (cudaSetDevice(device) != cudaSuccess) return BPPP_ERR_CUDA;
    bppp_ctx* c = new bppp_ctx();
    c->dev = device;
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return BPPP_ERR_CUDA;
    }
    if (cudaEventCreateWithFlags(&c->sync_ev, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess) c->sync_ev = nullptr;
    {   // keep freed blocks in the stream-ordered pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thr = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        // Grow the pool once, up front (BPPP_POOL_PREWARM_MB; default 8192 after bppp_tune_process(BPPP_TUNE_DEVICE),
        // else off): when many lanes
        // overlap in a new way the pool otherwise grows in the middle of a batch, and mapping fresh
        // device memory stalls every stream for a long time.
        static std::mutex mu;
        static bool done[64] = {false};
        std::lock_guard<std::mutex> lk(mu);
        if (device < 64 && !done[device]) {
            done[device] = true;
            const char* ev = getenv("BPPP_POOL_PREWARM_MB");
            size_t mb = ev ? (size_t)atoll(ev) : (g_bppp_tune_device ? 8192 : 0), free_b = 0, total_b = 0;
            if (mb && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b / 4 > (mb << 20)) {
                void* p = nullptr;
                if (cudaMallocAsync(&p, mb << 20, c->st) == cudaSuccess) cudaFreeAsync(p, c->st);
                cudaStreamSynchronize(c->st);
                cudaGetLastError();
            }
        }
    }
    *out = c;
    return BPPP_OK;
}
extern "C" void bppp_free(bppp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->dev);
    if (ctx->sync_ev) cudaEventDestroy(ctx->sync_ev);
    if (ctx->lut_count) cudaFree(ctx->lut_count);
    if (ctx->st) cudaStreamDestroy(ctx->st);
    delete ctx;
}
extern "C" int bppp_ctx_device(bppp_ctx* ctx) { return ctx ? ctx->dev : -1; }
// page-locked host staging memory (async H2D/D2H at full PCIe speed); caller frees with bppp_pinned_free
extern "C" int bppp_pinned_alloc(size_t bytes, void** out) {
    if (!out) return BPPP_ERR_ARG;
    *out = nullptr;
    return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable) == cudaSuccess ? BPPP_OK : BPPP_ERR_CUDA;
}
extern "C" void bppp_pinned_free(void* p) {
    if (p) cudaFreeHost(p);
}
extern "C" const char* bppp_last_error(bppp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" uint64_t bppp_launch_count(bppp_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int bppp_sync(bppp_ctx* ctx) {
    if (!ctx) return BPPP_ERR_ARG;
    ENTER(ctx);
    CK(ctx_sync(ctx));
    return BPPP_OK;
}

// =============================================================================== profiling
extern "C" int bppp_profile_enable(bppp_ctx* ctx, int on) {
    if (!ctx) return BPPP_ERR_ARG;
    cudaSetDevice(ctx->dev);
    prof_collect(ctx);
    ctx->prof = on != 0;
    return BPPP_OK;
}
extern "C" int bppp_profile_reset(bppp_ctx* ctx) {
    if (ctx && ctx->lut_count) { cudaSetDevice(ctx->dev); cudaMemsetAsync(ctx->lut_count, 0, 8, ctx->st); }
    if (!ctx) return BPPP_ERR_ARG;
    cudaSetDevice(ctx->dev);
    prof_collect(ctx);
    for (int i = 0; i < K_COUNT; i++) { ctx->k_ms[i] = 0; ctx->k_work[i] = 0; ctx->k_n[i] = 0; ctx->k_top_ms[i] = 0; ctx->k_top_work[i] = 0; }
    ctx->h2d = ctx->d2h = 0;
    return BPPP_OK;
}
// JSON: {"kernels": {"name": {"launches": n, "ms": total, "work": algorithmic units}, ...}, "h2d_bytes":.., "d2h_bytes":..}
extern "C" int bppp_profile_report(bppp_ctx* ctx, char* out, size_t cap) {
    if (!ctx || !out) return BPPP_ERR_ARG;
    cudaSetDevice(ctx->dev);
    prof_collect(ctx);
    std::string j = "{\"kernels\": {";
    bool first = true;
    char buf[256];
    for (int i = 0; i < K_COUNT; i++) {
        if (!ctx->k_n[i]) continue;
        snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %llu, \"ms\": %.6f, \"work\": %.6e, \"top_ms\": %.6f, \"top_work\": %.6e}",
                 first ? "" : ", ", kKernelNames[i], (unsigned long long)ctx->k_n[i], ctx->k_ms[i], ctx->k_work[i], ctx->k_top_ms[i],
                 ctx->k_top_work[i]);
        j += buf;
        first = false;
    }
    unsigned long long lookups = 0;
    if (ctx->lut_count) {
        cudaMemcpyAsync(&lookups, ctx->lut_count, 8, cudaMemcpyDeviceToHost, ctx->st);
        cudaStreamSynchronize(ctx->st);
    }
    snprintf(buf, sizeof buf, "}, \"h2d_bytes\": %llu, \"d2h_bytes\": %llu, \"launches\": %llu, \"lut_lookups\": %llu}", (unsigned long long)ctx->h2d,
             (unsigned long long)ctx->d2h, (unsigned long long)ctx->launches, lookups);
    j += buf;
    if (j.size() + 1 > cap) FAIL(BPPP_ERR_ARG, "report buffer too small");
    memcpy(out, j.c_str(), j.size() + 1);
    return BPPP_OK;
}
// CUDA-event timer on the context's stream (the stream every kernel of this library is launched on)
extern "C" int bppp_timer_start(bppp_ctx* ctx) {
    if (!ctx) return BPPP_ERR_ARG;
    ENTER(ctx);
    if (!ctx->t0) { CK(cudaEventCreate(&ctx->t0)); CK(cudaEventCreate(&ctx->t1)); }
    CK(ctx_sync(ctx));
    CK(cudaEventRecord(ctx->t0, ctx->st));
    return BPPP_OK;
}
extern "C" int bppp_timer_stop(bppp_ctx* ctx, double* ms) {
    if (!ctx || !ms || !ctx->t0) return BPPP_ERR_ARG;
    ENTER(ctx);
    CK(cudaEventRecord(ctx->t1, ctx->st));
    CK(cudaEventSynchronize(ctx->t1));
    float f = 0;
    CK(cudaEventElapsedTime(&f, ctx->t0, ctx->t1));
    *ms = f;
    return BPPP_OK;
}
// Measured integer peak: dependency-free streams of 32x32->64 multiply-accumulates (the IMAD.WIDE
// the field multiplier is made of) and of 32-bit IMADs, on every SM.  Results in ops/second.
__global__ void k_imad_peak_wide(unsigned long long* out, unsigned a, unsigned b, int iters) {
    unsigned long long acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    unsigned x = a + threadIdx.x, y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = (unsigned long long)x * y + acc[i];
            x ^= (unsigned)acc[0];
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    if (s == 0x1234567ULL) out[0] = s;
}
__global__ void k_imad_peak_lo(unsigned* out, unsigned a, unsigned b, int iters) {
    unsigned acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    unsigned x = a + threadIdx.x, y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = x * y + acc[i];
            x ^= acc[0];
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    if (s == 0x1234567u) out[0] = s;
}
extern "C" int bppp_measure_imad_peak(bppp_ctx* ctx, double* wide_per_s, double* lo_per_s) {
    if (!ctx || !wide_per_s || !lo_per_s) return BPPP_ERR_ARG;
    ENTER(ctx);
    DBuf<unsigned long long> d;
    CK(d.alloc(4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    double best[2] = {0, 0};
    for (int rep = 0; rep < 4; rep++) {
        for (int kind = 0; kind < 2; kind++) {
            CK(cudaEventRecord(e0, ctx->st));
            if (kind == 0) k_imad_peak_wide<<<blocks, threads, 0, ctx->st>>>(d.p, 12345u + rep, 6789u, iters);
            else k_imad_peak_lo<<<blocks, threads, 0, ctx->st>>>((unsigned*)d.p, 12345u + rep, 6789u, iters);
            CK(cudaGetLastError());
            CK(cudaEventRecord(e1, ctx->st));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            double ops = (double)blocks * threads * iters * 64.0 / (ms * 1e-3);
            if (rep > 0 && ops > best[kind]) best[kind] = ops;
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *wide_per_s = best[0]; *lo_per_s = best[1];
    return BPPP_OK;
}

// =============================================================================== MSM seam
extern "C" int bppp_msm_batch(bppp_ctx* ctx, size_t batch, size_t n, const uint8_t* scalars, const uint8_t* points,
                              int shared_points, uint8_t* out) {
    if (!ctx) return BPPP_ERR_ARG;
    if (!scalars || !points || !out || batch == 0) FAIL(BPPP_ERR_ARG, "bppp_msm_batch: null/empty argument");
    ENTER(ctx);
    if (n == 0) {                                   // the reference's innerProduct calls `head` on []
        memset(out, 0, batch * 64);                 // (src/Commitment.hs:328); here: the identity
        return BPPP_OK;
    }
    size_t npts = shared_points ? n : batch * n;
    if (!check_fr(scalars, batch * n)) FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    if (!check_fq(points, npts * 2)) FAIL(BPPP_ERR_RANGE, "coordinate >= field modulus");
    DBuf<u256> d_sc;
    DBuf<Affine> d_pts, d_aff;
    DBuf<Jac> d_res;
    CK(d_sc.alloc(batch * n));
    CK(d_pts.alloc(npts));
    CK(d_res.alloc(batch));
    CK(d_aff.alloc(batch));
    CK(H2D(d_sc.p, scalars, batch * n * 32));
    CK(H2D(d_pts.p, points, npts * 64));
    { int rc0 = check_points_sync(ctx, d_pts.p, npts, "bppp_msm_batch"); if (rc0) return rc0; }
    MsmPlan plan;
    plan.add(d_pts.p, shared_points ? 0 : n, d_sc.p, n, 0, n);
    int rc = run_msm(ctx, plan, batch, 1, d_res.p, msm_alg_imads((double)n));
    if (rc) return rc;
    rc = to_affine(ctx, d_res.p, 1, d_aff.p, 1, 0, 1, batch);
    if (rc) return rc;
    CK(D2H(out, d_aff.p, batch * 64));
    CK(ctx_sync(ctx));
    return BPPP_OK;
}
extern "C" int bppp_msm(bppp_ctx* ctx, size_t n, const uint8_t* scalars, const uint8_t* points, uint8_t out[64]) {
    return bppp_msm_batch(ctx, 1, n, scalars, points, 1, out);
}

// =============================================================================== generators, transcript (device)
// getPoints seed (app/Main.hs:68-72) on the device: candidates n = 0, 1, ... are hashed and tested in
// parallel, the survivors are compacted in order on the host.  Bit-identical to bppp_host_get_points.
extern "C" int bppp_get_points(bppp_ctx* ctx, const char* seed, size_t count, int root_policy, uint8_t* out) {
    if (!ctx) return BPPP_ERR_ARG;
    if (!seed || !out) FAIL(BPPP_ERR_ARG, "bppp_get_points: null argument");
    const size_t sl = strlen(seed);
    if (sl > 200) FAIL(BPPP_ERR_ARG, "bppp_get_points: seed too long");
    ENTER(ctx);
    DBuf<unsigned char> d_seed;
    DBuf<Affine> d_cand;
    CK(d_seed.alloc(sl + 1));
    CK(H2D(d_seed.p, seed, sl));
    size_t found = 0;
    uint64_t n0 = 0;
    std::vector<Affine> cand;
    while (found < count) {
        // about half of the candidates are on the curve
        const size_t chunk = std::min<size_t>((count - found) * 2 + (count - found) / 8 + 256, (size_t)1 << 22);
        CK(d_cand.ensure(chunk));
        cand.resize(chunk);
        { ProfScope ps_(ctx, K_TR_POINTS, 0);
        k_hash_to_curve<<<(unsigned)((chunk + TR_THREADS - 1) / TR_THREADS), TR_THREADS, 0, ctx->st>>>(d_seed.p, (int)sl, n0, chunk, root_policy, d_cand.p);
        }
        CK(cudaGetLastError());
        CK(D2H(cand.data(), d_cand.p, chunk * sizeof(Affine)));
        CK(ctx_sync(ctx));
        for (size_t i = 0; i < chunk && found < count; i++)
            if (!aff_is_inf(cand[i])) memcpy(out + 64 * found++, &cand[i], 64);
        n0 += chunk;
    }
    return BPPP_OK;
}

// Device transcript of `batch` proofs in lock-step: the commitment list of ZKPT (src/ZKP.hs:68-101)
// rendered in device memory (one right-aligned byte string per proof, k_tr_prepend), challenges by SHA-256 on the
// device (k_tr_squeeze_pair / k_tr_squeeze_coop).  Bit-identical to the host transcript (csrc/host/transcript.hpp).
struct bppp_dtr {
    bppp_ctx* ctx;
    size_t B, cap;                      // proofs, commitments per proof the store can hold
    unsigned SC;                        // bytes per proof
    int fmt;
    int n_calls = 0;                    // absorb calls so far
    unsigned ncoms[TR_MAX_CALLS + 1];   // commitments per proof after c calls
    bool fresh = false;                 // start[.][0] initialised on the device
    DBuf<unsigned char> buf;
    DBuf<unsigned> start;               // [B][TR_MAX_CALLS + 1]
    DBuf<u256> chal;                    // [B][TR_MAX_CHAL]
    DBuf<u256> chal_inv;                // the same challenges inverted (filled on request)
    DBuf<Affine> stage;
};
extern "C" int bppp_dtr_create(bppp_ctx* ctx, size_t batch, size_t max_points, int show_format, bppp_dtr** out) {
    if (!ctx) return BPPP_ERR_ARG;
    if (!out || batch == 0 || max_points == 0 || max_points > 60000) FAIL(BPPP_ERR_ARG, "bppp_dtr_create: bad argument");
    *out = nullptr;
    ENTER(ctx);
    bppp_dtr* t = new bppp_dtr();
    t->ctx = ctx; t->B = batch; t->cap = max_points; t->fmt = show_format;
    t->SC = (unsigned)((max_points * TR_PT_BYTES + 63) & ~(size_t)63);
    t->ncoms[0] = 0;
    cudaError_t e;
    if ((e = t->buf.alloc(batch * (size_t)t->SC + 64)) || (e = t->start.alloc(batch * (TR_MAX_CALLS + 1))) || (e = t->chal.alloc(batch * TR_MAX_CHAL))) {
        delete t;
        ctx->err = cudaGetErrorString(e);
        return BPPP_ERR_CUDA;
    }
    *out = t;
    return BPPP_OK;
}
extern "C" void bppp_dtr_destroy(bppp_dtr* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->dev);
    cudaStreamSynchronize(t->ctx->st);
    g_alloc_stream = t->ctx->st;
    delete t;
}
// can this transcript serve `batch` proofs of up to `max_points` commitments in format `show_format`?
extern "C" int bppp_dtr_fits(bppp_dtr* t, size_t batch, size_t max_points, int show_format) {
    return t && t->B == batch && t->cap >= max_points && t->fmt == show_format;
}
extern "C" int bppp_dtr_reset(bppp_dtr* t) {
    if (!t) return BPPP_ERR_ARG;
    t->n_calls = 0;
    t->ncoms[0] = 0;
    return BPPP_OK;
}
namespace {
// `oracle xs` part 1 (src/ZKP.hs:96-98): cs' = xs ++ cs, xs = npts points per proof already on the device
int dtr_absorb_dev(bppp_dtr* t, const Affine* pts, size_t pts_stride, size_t npts) {
    bppp_ctx* ctx = t->ctx;
    if (npts == 0) return BPPP_OK;
    if (t->ncoms[t->n_calls] + npts > t->cap || t->n_calls >= TR_MAX_CALLS) FAIL(BPPP_ERR_STATE, "device transcript: capacity exceeded");
    if (!t->fresh) {                    // state 0 of every proof: the empty body at the end of its buffer
        { ProfScope ps_(ctx, K_TR_RENDER, 0);
        k_fill_u32_strided<<<(unsigned)((t->B + 127) / 128), 128, 0, ctx->st>>>(t->start.p, TR_MAX_CALLS + 1, t->SC, t->B);
        }
        CK(cudaGetLastError());
        t->fresh = true;
    }
    { ProfScope ps_(ctx, K_TR_RENDER, 0);
    k_tr_prepend<<<(unsigned)t->B, TR_PREPEND_THREADS, 0, ctx->st>>>(pts, pts_stride, (int)npts, t->fmt, t->buf.p, t->SC, t->start.p, t->n_calls);
    }
    CK(cudaGetLastError());
    t->ncoms[t->n_calls + 1] = t->ncoms[t->n_calls] + (unsigned)npts;
    t->n_calls++;
    return BPPP_OK;
}
// part 2 (app/Main.hs:75-80): challenge j = scalar idx[j] of the transcript after state[j] absorb calls
// (state 0 = all calls so far); left on the device in t->chal, canonical, [batch][n_chal]
#define TR_COOP_MAX 296           // hashes per launch up to which the CTA-per-hash kernel is used (two per SM)
int dtr_squeeze_dev(bppp_dtr* t, int n_chal, const unsigned char* idx, const unsigned char* state) {
    bppp_ctx* ctx = t->ctx;
    if (n_chal < 1 || n_chal > TR_MAX_CHAL) FAIL(BPPP_ERR_ARG, "device transcript: 1..48 challenges per call");
    if (!t->fresh) {
        k_fill_u32_strided<<<(unsigned)((t->B + 127) / 128), 128, 0, ctx->st>>>(t->start.p, TR_MAX_CALLS + 1, t->SC, t->B);
        CK(cudaGetLastError());
        t->fresh = true;
    }
    TrPlan plan;
    plan.count = n_chal;
    for (int j = 0; j < n_chal; j++) {
        const int stt = state && state[j] ? state[j] : t->n_calls;
        if (idx[j] < 1 || idx[j] > 9 || stt > t->n_calls) FAIL(BPPP_ERR_ARG, "device transcript: bad challenge plan");
        plan.idx[j] = idx[j]; plan.state[j] = (unsigned char)stt; plan.ncoms[j] = t->ncoms[stt];
    }
    const size_t n = t->B * (size_t)n_chal;
    { ProfScope ps_(ctx, K_TR_SQUEEZE, 0);
    if (n <= TR_COOP_MAX)       // a few long hashes: latency is all that counts, a CTA per hash splits schedule and rounds
        k_tr_squeeze_coop<<<(unsigned)n, 64, 0, ctx->st>>>(t->buf.p, t->SC, t->start.p, plan, t->B, t->chal.p);
    else
        k_tr_squeeze_pair<<<(unsigned)((n + 31) / 32), 64, 0, ctx->st>>>(t->buf.p, t->SC, t->start.p, plan, t->B, t->chal.p);   // 32 hashes per CTA: every CTA gets an SM of its own
    }
    CK(cudaGetLastError());
    return BPPP_OK;
}
// 1 / challenge for the n_chal challenges of the last squeeze (canonical, [batch][n_chal]) in t->chal_inv: the
// provers and verifiers need e^-1, r^-1, q^-1 right away, and a field inversion is the most expensive scalar
// operation the host would otherwise do per proof
int dtr_invert_dev(bppp_dtr* t, int n_chal) {
    bppp_ctx* ctx = t->ctx;
    const size_t n = t->B * (size_t)n_chal;
    CK(t->chal_inv.ensure(t->B * TR_MAX_CHAL));
    { ProfScope ps_(ctx, K_TR_SQUEEZE, 0);
    k_fr_inv_rows<<<(unsigned)((n + 127) / 128), 128, 0, ctx->st>>>(t->chal.p, t->chal_inv.p, n);
    }
    CK(cudaGetLastError());
    return BPPP_OK;
}
int dtr_squeeze_first(bppp_dtr* t, int count) {        // scalars 1..count of the current transcript
    unsigned char idx[9];
    if (count < 1 || count > 9) { t->ctx->err = "device transcript: 1..9 challenges per oracle call"; return BPPP_ERR_ARG; }
    for (int j = 0; j < count; j++) idx[j] = (unsigned char)(j + 1);
    return dtr_squeeze_dev(t, count, idx, nullptr);
}
}  // namespace
// cs' = xs ++ cs without a challenge: pts = [batch] rows of `npts` points, row b at pts + 64 * stride_points * b
extern "C" int bppp_dtr_absorb(bppp_dtr* t, const uint8_t* pts, size_t stride_points, size_t npts) {
    if (!t) return BPPP_ERR_ARG;
    bppp_ctx* ctx = t->ctx;
    if (!pts || npts == 0 || stride_points < npts) FAIL(BPPP_ERR_ARG, "bppp_dtr_absorb: bad argument");
    ENTER(ctx);
    if (t->ncoms[t->n_calls] + npts > t->cap || t->n_calls >= TR_MAX_CALLS) FAIL(BPPP_ERR_STATE, "device transcript: capacity exceeded");
    for (size_t b = 0; b < t->B; b++)
        if (!check_fq(pts + 64 * stride_points * b, 2 * npts)) FAIL(BPPP_ERR_RANGE, "coordinate >= field modulus");
    // every call gets its own staging area: the copies of consecutive calls are queued behind each other
    const size_t off = t->B * (size_t)t->ncoms[t->n_calls];
    CK(t->stage.ensure(t->B * t->cap));
    ctx->h2d += t->B * npts * 64;
    CK(cudaMemcpy2DAsync(t->stage.p + off, npts * 64, pts, stride_points * 64, npts * 64, t->B, cudaMemcpyHostToDevice, ctx->st));
    return dtr_absorb_dev(t, t->stage.p + off, npts, npts);
}
// the rendered commitment list of one proof as the hash sees it (concat of show x <> show y, newest first):
// for tests that compare the device's `show` byte for byte with the reference's
extern "C" int bppp_dtr_export(bppp_dtr* t, size_t proof, uint8_t* out, size_t cap, size_t* len) {
    if (!t) return BPPP_ERR_ARG;
    bppp_ctx* ctx = t->ctx;
    if (!out || !len || proof >= t->B) FAIL(BPPP_ERR_ARG, "bppp_dtr_export: bad argument");
    ENTER(ctx);
    unsigned st0 = t->SC;
    if (t->n_calls) {
        CK(cudaMemcpyAsync(&st0, t->start.p + proof * (TR_MAX_CALLS + 1) + t->n_calls, 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(ctx_sync(ctx));
    }
    if (st0 > t->SC) FAIL(BPPP_ERR_STATE, "bppp_dtr_export: corrupt transcript offset");
    *len = t->SC - st0;
    if (*len > cap) FAIL(BPPP_ERR_ARG, "bppp_dtr_export: buffer too small");
    if (*len) {
        CK(cudaMemcpyAsync(out, t->buf.p + proof * (size_t)t->SC + st0, *len, cudaMemcpyDeviceToHost, ctx->st));
        CK(ctx_sync(ctx));
    }
    return BPPP_OK;
}
// challenges of any earlier stage: out[b][j] = scalar idx[j] after state[j] absorb calls (state NULL / 0: all calls)
static int dtr_squeeze_impl(bppp_dtr* t, size_t n_chal, const uint8_t* idx, const uint8_t* state, uint8_t* out, uint8_t* inv_out) {
    if (!t) return BPPP_ERR_ARG;
    bppp_ctx* ctx = t->ctx;
    if (!idx || !out) FAIL(BPPP_ERR_ARG, "bppp_dtr_squeeze: null argument");
    ENTER(ctx);
    int rc = dtr_squeeze_dev(t, (int)n_chal, idx, state);
    if (rc) return rc;
    CK(D2H(out, t->chal.p, t->B * n_chal * 32));
    if (inv_out) {
        if ((rc = dtr_invert_dev(t, (int)n_chal))) return rc;
        CK(D2H(inv_out, t->chal_inv.p, t->B * n_chal * 32));
    }
    CK(ctx_sync(ctx));
    return BPPP_OK;
}
extern "C" int bppp_dtr_squeeze(bppp_dtr* t, size_t n_chal, const uint8_t* idx, const uint8_t* state, uint8_t* out) {
    return dtr_squeeze_impl(t, n_chal, idx, state, out, nullptr);
}
// the same plus inv_out[b][j] = 1 / out[b][j] (0 -> 0), inverted on the device
extern "C" int bppp_dtr_squeeze_inv(bppp_dtr* t, size_t n_chal, const uint8_t* idx, const uint8_t* state, uint8_t* out, uint8_t* inv_out) {
    if (!inv_out) return BPPP_ERR_ARG;
    return dtr_squeeze_impl(t, n_chal, idx, state, out, inv_out);
}
// `oracle xs` for host-resident commitments: pts = [batch][npts] points, out = [batch][count] challenges
extern "C" int bppp_dtr_oracle(bppp_dtr* t, const uint8_t* pts, size_t npts, int count, uint8_t* out) {
    if (!t) return BPPP_ERR_ARG;
    bppp_ctx* ctx = t->ctx;
    if (!out || (npts && !pts)) FAIL(BPPP_ERR_ARG, "bppp_dtr_oracle: null argument");
    ENTER(ctx);
    if (npts) {
        int rc = bppp_dtr_absorb(t, pts, npts, npts);
        if (rc) return rc;
    }
    int rc = dtr_squeeze_first(t, count);
    if (rc) return rc;
    CK(D2H(out, t->chal.p, t->B * (size_t)count * 32));
    CK(ctx_sync(ctx));
    return BPPP_OK;
}
// `random` (src/ZKP.hs:90-93, app/Main.hs:177): out[b][j] = hash(seed_b <> show (n0 + j)), j < count
extern "C" int bppp_dev_random(bppp_ctx* ctx, size_t batch, const char* const* seeds, uint64_t n0, size_t count, uint8_t* out) {
    if (!ctx) return BPPP_ERR_ARG;
    if (!seeds || !out || batch == 0 || count == 0) FAIL(BPPP_ERR_ARG, "bppp_dev_random: null/empty argument");
    ENTER(ctx);
    std::vector<unsigned char> hs(batch * 64, 0), hl(batch);
    for (size_t b = 0; b < batch; b++) {
        const size_t l = strlen(seeds[b]);
        if (l > 40) FAIL(BPPP_ERR_ARG, "bppp_dev_random: seed longer than 40 bytes");
        memcpy(&hs[64 * b], seeds[b], l);
        hl[b] = (unsigned char)l;
    }
    DBuf<unsigned char> ds, dl;
    DBuf<u256> d_out;
    CK(ds.alloc(batch * 64)); CK(dl.alloc(batch)); CK(d_out.alloc(batch * count));
    CK(H2D(ds.p, hs.data(), batch * 64));
    CK(H2D(dl.p, hl.data(), batch));
    { ProfScope ps_(ctx, K_TR_RANDOM, 0);
    k_tr_random<<<(unsigned)((batch * count + TR_THREADS - 1) / TR_THREADS), TR_THREADS, 0, ctx->st>>>(ds.p, dl.p, n0, nullptr, batch, count, d_out.p, count);
    }
    CK(cudaGetLastError());
    CK(D2H(out, d_out.p, batch * count * 32));
    CK(ctx_sync(ctx));
    return BPPP_OK;
}

// =============================================================================== fixed base
struct bppp_fb {
    bppp_ctx* ctx;
    size_t n_bases;
    DBuf<Affine> tbl;
};
extern "C" int bppp_fb_create(bppp_ctx* ctx, size_t n_bases, const uint8_t* points, bppp_fb** out) {
    if (!ctx) return BPPP_ERR_ARG;
    if (!out || !points || n_bases == 0 || n_bases > 16) FAIL(BPPP_ERR_ARG, "bppp_fb_create: bad argument");
    *out = nullptr;
    ENTER(ctx);
    if (!check_fq(points, 2 * n_bases)) FAIL(BPPP_ERR_RANGE, "coordinate >= field modulus");
    bppp_fb* fb = new bppp_fb();
    fb->ctx = ctx; fb->n_bases = n_bases;
    size_t total = n_bases * FB_WINDOWS * FB_ENTRIES;
    DBuf<Affine> d_b;
    DBuf<Jac> d_j;
    cudaError_t e;
    if ((e = d_b.alloc(n_bases)) || (e = d_j.alloc(total)) || (e = fb->tbl.alloc(total))) {
        delete fb;
        ctx->err = cudaGetErrorString(e);
        return BPPP_ERR_CUDA;
    }
    H2D(d_b.p, points, n_bases * 64);
    { int rc0 = check_points_sync(ctx, d_b.p, n_bases, "bppp_fb_create"); if (rc0) { delete fb; return rc0; } }
    int nt = (int)(n_bases * FB_WINDOWS);
    { ProfScope ps_(ctx, K_FB_BUILD, WORK_K_FB_BUILD);
    k_fb_build<<<(nt + 31) / 32, 32, 0, ctx->st>>>(d_b.p, (int)n_bases, d_j.p);
    }
    int rc = to_affine(ctx, d_j.p, total, fb->tbl.p, total, 0, (int)total, total);
    if (rc == 0 && (e = ctx_sync(ctx)) != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = BPPP_ERR_CUDA; }
    if (rc) { delete fb; return rc; }
    *out = fb;
    return BPPP_OK;
}
extern "C" int bppp_fb_msm_batch(bppp_fb* fb, size_t batch, const uint8_t* scalars, uint8_t* out) {
    if (!fb) return BPPP_ERR_ARG;
    bppp_ctx* ctx = fb->ctx;
    if (!scalars || !out || batch == 0) FAIL(BPPP_ERR_ARG, "bppp_fb_msm_batch: null/empty argument");
    ENTER(ctx);
    if (!check_fr(scalars, batch * fb->n_bases)) FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    DBuf<u256> d_sc;
    DBuf<Jac> d_res;
    DBuf<Affine> d_aff;
    CK(d_sc.alloc(batch * fb->n_bases)); CK(d_res.alloc(batch)); CK(d_aff.alloc(batch));
    CK(H2D(d_sc.p, scalars, batch * fb->n_bases * 32));
    { ProfScope ps_(ctx, K_FB_MSM, WORK_K_FB_MSM);
    k_fb_msm<<<(unsigned)((batch + 127) / 128), 128, 0, ctx->st>>>(fb->tbl.p, (int)fb->n_bases, d_sc.p, d_res.p, batch);
    }
    CK(cudaGetLastError());
    int rc = to_affine(ctx, d_res.p, 1, d_aff.p, 1, 0, 1, batch);
    if (rc) return rc;
    CK(D2H(out, d_aff.p, batch * 64));
    CK(ctx_sync(ctx));
    return BPPP_OK;
}
extern "C" void bppp_fb_destroy(bppp_fb* fb) { delete fb; }

// =============================================================================== pair fold
namespace {
int launch_pair_fold(bppp_ctx* ctx, const Affine* in, size_t in_stride, Jac* out, size_t out_stride,
                     const PairFoldSeg* segs, int n_seg, const u256* kb, const u256* ka, const unsigned char* sgn,
                     size_t batch) {
    static bool attr = false;
    const size_t smem = 4 * 16 * PF_THREADS * sizeof(uint32_t);
    if (!attr) {
        CK(cudaFuncSetAttribute(k_pair_fold, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_pair_fold, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr = true;
    }
    PairFoldArgs A;
    A.in = in; A.in_stride = in_stride; A.out = out; A.out_stride = out_stride;
    A.n_seg = 0; A.kb = kb; A.ka = ka; A.sgn = sgn;
    int total_blocks = 0;
    for (int s = 0; s < n_seg; s++) {
        A.seg[A.n_seg] = segs[s];
        int n_out = (segs[s].n_in + 1) / 2;
        A.blocks_per_seg[A.n_seg] = (n_out + PF_THREADS - 1) / PF_THREADS;
        total_blocks += A.blocks_per_seg[A.n_seg];
        A.n_seg++;
    }
    if (total_blocks == 0) return BPPP_OK;
    for (size_t b0 = 0; b0 < batch; b0 += 65535) {
        size_t nb = std::min<size_t>(65535, batch - b0);
        PairFoldArgs B = A;
        {   // half-length Shamir (129 dbl * 7 + 97 add * 11 Fq mults) * 136 IMAD per folded point
            double outs = 0;
            for (int s2 = 0; s2 < n_seg; s2++) outs += (segs[s2].n_in + 1) / 2;
            g_work = outs * (double)nb * (129.0 * 7 + 97.0 * 11) * 136.0;
        }
        B.in = in + b0 * in_stride; B.out = out + b0 * out_stride;
        B.kb = kb + b0 * n_seg; B.ka = ka + b0 * n_seg; B.sgn = sgn + b0 * n_seg;
        { ProfScope ps_(ctx, K_PAIR_FOLD, WORK_K_PAIR_FOLD);
        k_pair_fold<<<dim3(total_blocks, (unsigned)nb), PF_THREADS, smem, ctx->st>>>(B);
        }
        CK(cudaGetLastError());
    }
    return BPPP_OK;
}
}  // namespace

extern "C" int bppp_pair_fold(bppp_ctx* ctx, size_t n_in, const uint8_t a[32], int a_neg, const uint8_t b[32],
                              int b_neg, const uint8_t* points_in, uint8_t* points_out) {
    if (!ctx) return BPPP_ERR_ARG;
    if (!a || !b || !points_in || !points_out) FAIL(BPPP_ERR_ARG, "bppp_pair_fold: null argument");
    if (n_in == 0) return BPPP_OK;
    ENTER(ctx);
    if (!check_fq(points_in, n_in * 2)) FAIL(BPPP_ERR_RANGE, "coordinate >= field modulus");
    u256 ka = host::from_bytes(a), kb = host::from_bytes(b);
    if (ka.v[7] >> 31 || kb.v[7] >> 31) FAIL(BPPP_ERR_RANGE, "fold scalar magnitude >= 2^255");
    size_t n_out = (n_in + 1) / 2;
    DBuf<Affine> d_in, d_out;
    DBuf<Jac> d_j;
    DBuf<u256> d_k;
    DBuf<unsigned char> d_s;
    CK(d_in.alloc(n_in)); CK(d_out.alloc(n_out)); CK(d_j.alloc(n_out)); CK(d_k.alloc(2)); CK(d_s.alloc(1));
    unsigned char sg = (unsigned char)((b_neg ? 1 : 0) | (a_neg ? 2 : 0));
    u256 ks[2] = {kb, ka};
    CK(H2D(d_in.p, points_in, n_in * 64));
    CK(H2D(d_k.p, ks, 64));
    CK(H2D(d_s.p, &sg, 1));
    { int rc0 = check_points_sync(ctx, d_in.p, n_in, "bppp_pair_fold"); if (rc0) return rc0; }
    PairFoldSeg seg = {0, (int)n_in, 0};
    int rc = launch_pair_fold(ctx, d_in.p, 0, d_j.p, 0, &seg, 1, d_k.p, d_k.p + 1, d_s.p, 1);
    if (rc) return rc;
    rc = to_affine(ctx, d_j.p, n_out, d_out.p, n_out, 0, (int)n_out, n_out);
    if (rc) return rc;
    CK(D2H(points_out, d_out.p, n_out * 64));
    CK(ctx_sync(ctx));
    return BPPP_OK;
}

extern "C" int bppp_rational_reduce(const uint8_t x[32], uint8_t a[32], int* a_neg, uint8_t b[32], int* b_neg) {
    if (!x || !a || !b || !a_neg || !b_neg) return BPPP_ERR_ARG;
    u256 xv = host::from_bytes(x);
    if (!host::fr_is_canonical(xv)) return BPPP_ERR_RANGE;
    host::Ratio r = host::rational_reduce(xv);
    host::to_bytes(a, r.a); host::to_bytes(b, r.b);
    *a_neg = r.a_neg; *b_neg = r.b_neg;
    return BPPP_OK;
}

// =============================================================================== generator tables
// The shared generator list [g | G | H] with its fixed-base window table, resident on the device.
// full-multiples tables (lut.cuh) are large and depend only on the generator list: one per (device, list), shared by
// every bppp_gens over that list (the lanes of a setup, prover and verifier setups of one process)
struct LutTable {
    int dev = 0;
    LutDesc D;
    std::vector<uint8_t> key;        // the generator list, P0 * 64 bytes
    int refs = 0;
    double build_ms = 0;
};
static std::mutex g_lut_mu;
static std::vector<LutTable*> g_luts;

struct bppp_gens {
    bppp_ctx* ctx;
    LutTable* lut = nullptr;
    size_t N, M, P0;
    DBuf<Affine> base;               // [g | G | H]
    DBuf<Affine> tbl;                // [P0][GT_W]; empty for long lists (P0 > GT_TABLE_MAX_TERMS)
    std::vector<uint8_t> host;       // P0 * 64 bytes (for the bucket-kernel fallback paths)
    PipWork pip;                     // scratch of the size-aware MSM over `base` (long lists)
};
// The window table pays off for batches of short proofs (29 table points per generator, no doublings,
// one bucket reduction per MSM).  For a long list the size-aware Pippenger over the bare generators needs
// FEWER additions (16 windows of 16 bits at n = 2^20 against 29 of 9 bits) and no 1.9 GB table.
#define GT_TABLE_MAX_TERMS 8192
namespace {
size_t gt_smem_bytes(int n) { return (size_t)(2 * GT_KEYS + 1) * 4 + (size_t)n * GT_W * 2 + 16; }

// fixed-base MSMs over the first n_terms generators: scalars sc[p*sc_stride + o*sc_out_stride + i];
// result Jacobian points in d_out[p*n_out + o]
int run_msm_gens(bppp_gens* g, size_t n_terms, const u256* sc, size_t sc_stride, size_t sc_out_stride, size_t batch,
                 int n_out, Jac* d_out, double work_per_proof) {
    bppp_ctx* ctx = g->ctx;
    if (!g->tbl.p)
        return run_msm_pip(ctx, g->pip, g->base.p, 0, sc, sc_stride, sc_out_stride, n_terms, batch, n_out, d_out, work_per_proof);
    if (g->lut) {
        // full-multiples table: W lookups + mixed additions per term, one CTA per (MSM, chunk), no reduction kernel
        // CTAs of 64 threads: cut every MSM into enough chunks for BPPP_LUT_WAVES (default 3) x 148 x 8 CTAs -- 4 chunks for
        // 1024 MSMs, i.e. 2.8 waves of the 10 CTAs per SM that fit (CTAs that finish are replaced while others still
        // add: an R commitment has half the work of an X commitment).  Fewer chunks: less tree-summing of partials;
        // more: a fuller last wave when the launch runs alone.  At least 64 terms per chunk; k_jac_sum adds the chunk sums
        const size_t n_msm_all = batch * (size_t)n_out;
        static const int lut_waves = [] { const char* e = getenv("BPPP_LUT_WAVES"); return e && atoi(e) > 0 ? atoi(e) : 3; }();
        size_t want = ((size_t)lut_waves * 148 * 8 + n_msm_all - 1) / n_msm_all;
        // (a lone proof: 32 terms per chunk -- one (scalar, half) unit per thread, the latency of 8 additions and the tree)
        const size_t min_chunk = n_msm_all <= 8 ? 32 : 64;
        want = std::max<size_t>(1, std::min<size_t>(want, (n_terms + min_chunk - 1) / min_chunk));
        const size_t chunk_terms = (n_terms + want - 1) / want;
        const int nch = (int)((n_terms + chunk_terms - 1) / chunk_terms);
        DBuf<Jac> partsbuf;
        Jac* parts = d_out;
        if (nch > 1) { CK(partsbuf.alloc(batch * n_out * nch)); parts = partsbuf.p; }
        for (size_t b0 = 0; b0 < batch; b0 += 32768) {
            const size_t nb = std::min<size_t>(32768, batch - b0);
            LutMsmArgs A;
            A.D = g->lut->D; A.sc = sc + b0 * sc_stride; A.sc_stride = sc_stride; A.sc_out_stride = sc_out_stride;
            A.n_terms = (int)n_terms; A.chunk_terms = (int)chunk_terms;
            A.out = parts + b0 * n_out * nch; A.out_pstride = (size_t)n_out * nch; A.n_out = n_out; A.n_chunks = nch;
            A.count = nullptr;
            if (ctx->prof) {                       // profiling pass: count the lookups actually made
                if (!ctx->lut_count) {
                    CK(cudaMalloc((void**)&ctx->lut_count, 8));
                    CK(cudaMemsetAsync(ctx->lut_count, 0, 8, ctx->st));
                }
                A.count = ctx->lut_count;
            }
            g_work = work_per_proof * (double)nb;
            { ProfScope ps_(ctx, K_MSM_LUT, g_work);
            k_msm_lut<<<dim3(nch, n_out, (unsigned)nb), LUT_THREADS, 0, ctx->st>>>(A);
            }
            CK(cudaGetLastError());
        }
        if (nch > 1) {
            const size_t n_msm = batch * n_out;
            { ProfScope ps_(ctx, K_JAC_SUM, 0);
            if (n_msm <= 64) k_jac_sum_warp<<<(unsigned)n_msm, 32, 0, ctx->st>>>(parts, nch, d_out, n_msm);
            else k_jac_sum<<<(unsigned)((n_msm + 127) / 128), 128, 0, ctx->st>>>(parts, nch, nullptr, 0, d_out, n_msm);
            }
            CK(cudaGetLastError());
        }
        return BPPP_OK;
    }
    static bool attr = false;
    if (!attr) {
        CK(cudaFuncSetAttribute(k_msm_gens, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gt_smem_bytes(GT_MAX_CHUNK)));
        CK(cudaFuncSetAttribute(k_msm_gens, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr = true;
    }
    // terms per CTA: 2048 for batches; a lone proof is cut finer so that more SMs share its (latency-bound) work
    size_t chunk_terms = GT_MAX_CHUNK;
    const bool lone = batch * (size_t)n_out <= 8;
    if (lone) chunk_terms = std::min<size_t>(GT_MAX_CHUNK, std::max<size_t>(128, (n_terms + 31) / 32));
    int nch = (int)((n_terms + chunk_terms - 1) / chunk_terms);
    size_t ctas = batch * n_out * nch;
    DBuf<unsigned char> scratch;     // stream-ordered pool allocations: cheap, and safe across lanes
    DBuf<Jac> partsbuf;
    CK(scratch.alloc(ctas * GT_SCRATCH_BYTES(GT_THREADS)));
    Jac* parts = d_out;
    if (nch > 1) { CK(partsbuf.alloc(ctas)); parts = partsbuf.p; }
    int max_n = (int)std::min<size_t>(chunk_terms, n_terms);
    for (size_t b0 = 0; b0 < batch; b0 += 32768) {
        size_t nb = std::min<size_t>(32768, batch - b0);
        GtArgs A;
        A.tbl = g->tbl.p; A.sc = sc + b0 * sc_stride; A.sc_stride = sc_stride; A.sc_out_stride = sc_out_stride;
        A.n_total = (int)n_terms; A.term0 = 0; A.chunk_terms = (int)chunk_terms;
        A.scratch = scratch.p + b0 * n_out * nch * GT_SCRATCH_BYTES(GT_THREADS);
        A.out = parts + b0 * n_out * nch; A.out_pstride = (size_t)n_out * nch; A.n_out = n_out; A.n_chunks = nch;
        g_work = work_per_proof * (double)nb;
        { ProfScope ps_(ctx, K_MSM_GENS, g_work);
        k_msm_gens<<<dim3(nch, n_out, (unsigned)nb), GT_THREADS, gt_smem_bytes(max_n), ctx->st>>>(A);
        }
        CK(cudaGetLastError());
        const size_t n_cta = nb * n_out * nch;
        { ProfScope ps_(ctx, K_MSM_REDUCE, 0);
        if (lone)
            k_msm_gens_reduce_lat<<<(unsigned)n_cta, 32, 0, ctx->st>>>(A.scratch, GT_SCRATCH_BYTES(GT_THREADS), A.out, A.out_pstride, n_out, nch, n_cta);
        else
            k_msm_gens_reduce<<<(unsigned)((n_cta + 7) / 8), 256, 0, ctx->st>>>(A.scratch, GT_SCRATCH_BYTES(GT_THREADS), A.out, A.out_pstride,
                                                                               n_out, nch, n_cta);
        }
        CK(cudaGetLastError());
    }
    if (nch > 1) {
        size_t n_msm = batch * n_out;
        { ProfScope ps_(ctx, K_JAC_SUM, 0);
        if (lone) k_jac_sum_warp<<<(unsigned)n_msm, 32, 0, ctx->st>>>(parts, nch, d_out, n_msm);
        else k_jac_sum<<<(unsigned)((n_msm + 127) / 128), 128, 0, ctx->st>>>(parts, nch, nullptr, 0, d_out, n_msm);
        }
        CK(cudaGetLastError());
    }
    return BPPP_OK;
}
// One small fixed-base MSM per GROUP of `group` consecutive generators [term0 + j*group, ...) of every
// proof (scalars sc[p*sc_stride + term]): out[p*out_pstride + j], j < ceil(n_terms/group).  Used to
// materialise the folded generators of a tensor-mode argument (G'_j = sum over its block of
// coef_i * G_i) when it switches to folding.
int run_msm_groups(bppp_gens* g, const u256* sc, size_t sc_stride, size_t batch, size_t term0, size_t n_terms, size_t group,
                   Jac* out, size_t out_pstride) {
    bppp_ctx* ctx = g->ctx;
    if (n_terms == 0) return BPPP_OK;
    static bool attr = false;
    if (!attr) {
        CK(cudaFuncSetAttribute(k_msm_gens_small, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr = true;
    }
    if (group > 512) FAIL(BPPP_ERR_ARG, "run_msm_groups: group too large");
    const size_t ng = (n_terms + group - 1) / group;
    const size_t smem = gt_smem_bytes((int)group);
    // the per-CTA scratch is large (61 KB): launch in slices of proofs
    size_t per = std::max<size_t>(1, std::min<size_t>(batch, 4096 / std::max<size_t>(1, ng)));
    per = std::min<size_t>(per, 32768);
    DBuf<unsigned char> scratch;
    CK(scratch.alloc(per * ng * GT_SCRATCH_BYTES(GT_THREADS_SMALL)));
    for (size_t b0 = 0; b0 < batch; b0 += per) {
        const size_t nb = std::min(per, batch - b0);
        GtArgs A;
        A.tbl = g->tbl.p; A.sc = sc + b0 * sc_stride; A.sc_stride = sc_stride; A.sc_out_stride = 0;
        A.n_total = (int)n_terms; A.term0 = (int)term0; A.chunk_terms = (int)group;
        A.scratch = scratch.p; A.out = out + b0 * out_pstride; A.out_pstride = out_pstride; A.n_out = 1; A.n_chunks = (int)ng;
        { ProfScope ps_(ctx, K_MSM_GROUPS, 0);
        k_msm_gens_small<<<dim3((unsigned)ng, 1, (unsigned)nb), GT_THREADS_SMALL, smem, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
        { ProfScope ps_(ctx, K_MSM_REDUCE, 0);
        k_msm_gens_reduce<<<(unsigned)((nb * ng + 7) / 8), 256, 0, ctx->st>>>(A.scratch, GT_SCRATCH_BYTES(GT_THREADS_SMALL), A.out, A.out_pstride,
                                                                            1, (int)ng, nb * ng);
        }
        CK(cudaGetLastError());
    }
    return BPPP_OK;
}
}  // namespace

namespace {
// the resident list [g | G | H] (+ window table when it is short enough) from host bytes, or from a
// device array `d_src` of 1 + N + M affine points that is already in list order
int gens_create_impl(bppp_ctx* ctx, size_t N, size_t M, const uint8_t* g, const uint8_t* G, const uint8_t* H,
                     const Affine* d_src, bppp_gens** out) {
    bppp_gens* gg = new bppp_gens();
    gg->ctx = ctx; gg->N = N; gg->M = M; gg->P0 = 1 + N + M;
    if (!d_src) {
        gg->host.resize(gg->P0 * 64);
        memcpy(&gg->host[0], g, 64);
        if (N) memcpy(&gg->host[64], G, N * 64);
        if (M) memcpy(&gg->host[64 * (1 + N)], H, M * 64);
    }
    DBuf<Jac> tj;
    cudaError_t e;
    const bool with_table = gg->P0 <= GT_TABLE_MAX_TERMS;
    size_t total = with_table ? gg->P0 * GT_W : 0;
    if ((e = gg->base.alloc(gg->P0)) || (with_table && ((e = gg->tbl.alloc(total)) || (e = tj.alloc(total))))) {
        delete gg;
        ctx->err = cudaGetErrorString(e);
        return BPPP_ERR_CUDA;
    }
    if (d_src) {
        if ((e = cudaMemcpyAsync(gg->base.p, d_src, gg->P0 * 64, cudaMemcpyDeviceToDevice, ctx->st)) != cudaSuccess) {
            delete gg;
            ctx->err = cudaGetErrorString(e);
            return BPPP_ERR_CUDA;
        }
    } else {
        H2D(gg->base.p, gg->host.data(), gg->P0 * 64);
        int rc0 = check_points_sync(ctx, gg->base.p, gg->P0, "bppp_gens_create");
        if (rc0) { delete gg; return rc0; }
    }
    int rc = BPPP_OK;
    if (with_table) {
        { ProfScope ps_(ctx, K_GT_BUILD, 0);
        k_gt_build<<<(unsigned)((gg->P0 + 63) / 64), 64, 0, ctx->st>>>(gg->base.p, gg->P0, tj.p);
        }
        rc = to_affine(ctx, tj.p, total, gg->tbl.p, total, 0, (int)std::min<size_t>(total, 0x7fffffff), total);
    }
    if (rc == 0 && (e = ctx_sync(ctx)) != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = BPPP_ERR_CUDA; }
    if (rc) { delete gg; return rc; }
    *out = gg;
    return BPPP_OK;
}
}  // namespace
extern "C" int bppp_gens_create(bppp_ctx* ctx, size_t N, size_t M, const uint8_t* g, const uint8_t* G, const uint8_t* H,
                                bppp_gens** out) {
    if (!ctx) return BPPP_ERR_ARG;
    if (!out || !g || (N && !G) || (M && !H)) FAIL(BPPP_ERR_ARG, "bppp_gens_create: null argument");
    *out = nullptr;
    ENTER(ctx);
    if (!check_fq(g, 2) || !check_fq(G, 2 * N) || !check_fq(H, 2 * M)) FAIL(BPPP_ERR_RANGE, "coordinate >= field modulus");
    return gens_create_impl(ctx, N, M, g, G, H, nullptr, out);
}
// Give the generator list a full-multiples table (lut.cuh) of at most `budget_gb` gigabytes (and at most half of the
// free device memory): window width c = 10 .. 16 by what fits; nothing happens when not even c = 10 fits.  Tables are
// shared by all generator sets over the same list on the same device.  Returns the window width in *c_out (0 = none).
extern "C" int bppp_gens_enable_lut(bppp_gens* g, double budget_gb, int* c_out) {
    if (!g) return BPPP_ERR_ARG;
    bppp_ctx* ctx = g->ctx;
    if (c_out) *c_out = g->lut ? g->lut->D.c : 0;
    if (g->lut || budget_gb <= 0) return BPPP_OK;
    if (g->host.size() != g->P0 * 64 || !g->tbl.p) return BPPP_OK;       // long lists / device-made lists keep their paths
    ENTER(ctx);
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const double budget = std::min(budget_gb * 1e9, 0.5 * (double)free_b);
    std::lock_guard<std::mutex> lk(g_lut_mu);
    for (auto* t : g_luts)
        if (t->dev == ctx->dev && t->key == g->host) {                   // built already (another lane / setup)
            t->refs++;
            g->lut = t;
            if (c_out) *c_out = t->D.c;
            return BPPP_OK;
        }
    int c = 0;
    for (int cc = 16; cc >= 10; cc--) {
        const double bytes = (double)g->P0 * ((256 + cc - 1) / cc) * (double)(1u << (cc - 1)) * 64.0;
        if (bytes <= budget) { c = cc; break; }
    }
    if (!c) return BPPP_OK;
    LutTable* t = new LutTable();
    t->dev = ctx->dev; t->key = g->host; t->refs = 1;
    LutDesc& D = t->D;
    D.c = c; D.W = (256 + c - 1) / c; D.NB = 1 << (c - 1); D.P0 = g->P0; D.tbl = nullptr; D.carry = nullptr;
    const size_t n_ent = g->P0 * (size_t)D.W * D.NB, n_base = g->P0 * (size_t)(D.W + 1);
    Jac* bj = nullptr;
    Affine* ba = nullptr;
    auto fail = [&](cudaError_t e) {
        ctx->err = std::string("bppp_gens_enable_lut: ") + cudaGetErrorString(e);
        cudaFree(D.tbl); cudaFree(D.carry); cudaFree(bj); cudaFree(ba);
        delete t;
        return BPPP_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaMalloc((void**)&D.tbl, n_ent * 64)) || (e = cudaMalloc((void**)&D.carry, g->P0 * 64)) ||
        (e = cudaMalloc((void**)&bj, n_base * sizeof(Jac))) || (e = cudaMalloc((void**)&ba, n_base * 64)))
        return fail(e);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, ctx->st);
    { ProfScope ps_(ctx, K_LUT_BUILD, 0);
    k_lut_bases<<<(unsigned)((g->P0 + 63) / 64), 64, 0, ctx->st>>>(g->base.p, g->P0, c, D.W, bj);
    }
    if ((e = cudaGetLastError())) return fail(e);
    int rc = to_affine(ctx, bj, n_base, ba, n_base, 0, (int)n_base, n_base);
    if (rc) { fail(cudaErrorUnknown); return rc; }
    if ((e = cudaMemcpy2DAsync(D.carry, 64, ba + D.W, (size_t)(D.W + 1) * 64, 64, g->P0, cudaMemcpyDeviceToDevice, ctx->st))) return fail(e);
    const int run = std::min(LUT_RUN, D.NB);
    const size_t n_thr = g->P0 * (size_t)D.W * (D.NB / run);
    { ProfScope ps_(ctx, K_LUT_BUILD, 0);
    k_lut_fill<<<(unsigned)((n_thr + 127) / 128), 128, 0, ctx->st>>>(D, ba, 0, n_thr, run);
    }
    if ((e = cudaGetLastError())) return fail(e);
    cudaEventRecord(e1, ctx->st);
    if ((e = ctx_sync(ctx))) return fail(e);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    t->build_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(bj); cudaFree(ba);
    g_luts.push_back(t);
    g->lut = t;
    if (c_out) *c_out = c;
    return BPPP_OK;
}
extern "C" void bppp_gens_destroy(bppp_gens* g) {
    if (!g) return;
    cudaSetDevice(g->ctx->dev);
    cudaStreamSynchronize(g->ctx->st);
    if (g->lut) {
        std::lock_guard<std::mutex> lk(g_lut_mu);
        if (--g->lut->refs == 0) {
            g_luts.erase(std::remove(g_luts.begin(), g_luts.end(), g->lut), g_luts.end());
            cudaFree(g->lut->D.tbl);
            cudaFree(g->lut->D.carry);
            delete g->lut;
        }
    }
    delete g;
}
// `batch` MSMs over the first n generators of the list (commitRPW over [g | gs | hs])
extern "C" int bppp_gens_msm_batch(bppp_gens* g, size_t batch, size_t n, const uint8_t* scalars, uint8_t* out) {
    if (!g) return BPPP_ERR_ARG;
    bppp_ctx* ctx = g->ctx;
    if (!scalars || !out || batch == 0 || n == 0 || n > g->P0) FAIL(BPPP_ERR_ARG, "bppp_gens_msm_batch: bad argument");
    ENTER(ctx);
    if (!check_fr(scalars, batch * n)) FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    DBuf<u256> d_sc;
    DBuf<Jac> d_res;
    DBuf<Affine> d_aff;
    CK(d_sc.alloc(batch * n)); CK(d_res.alloc(batch)); CK(d_aff.alloc(batch));
    CK(H2D(d_sc.p, scalars, batch * n * 32));
    int rc = run_msm_gens(g, n, d_sc.p, n, 0, batch, 1, d_res.p, msm_alg_imads((double)n));
    if (rc) return rc;
    if ((rc = to_affine(ctx, d_res.p, 1, d_aff.p, 1, 0, 1, batch))) return rc;
    CK(D2H(out, d_aff.p, batch * 64));
    CK(ctx_sync(ctx));
    return BPPP_OK;
}

// =============================================================================== argument seam
using h64::Fr;
struct bppp_nl {
    bppp_ctx* ctx;
    bppp_gens* gens = nullptr;
    bool own_gens = false;
    int kind;
    size_t B, N, M;                 // batch, initial lengths
    size_t curN, curM;              // current lengths
    size_t N2, M2, P0, P2;          // half lengths, point strides (round 0 shared / later per proof)
    int round = 0;
    bool have_partials = false;
    int cur = 0;                    // which scalar buffer holds the current vectors
    int curp = -1;                  // which per-proof point buffer is current (-1: shared generators)
    DBuf<Affine> pts[2], aff;
    DBuf<u256> w[2], l[2], c[2];
    size_t wstride[2], lstride[2];
    DBuf<u256> sc;                  // [2][B][1+N+M] canonical MSM scalars (X then R)
    DBuf<u256> part_n, part_l, dots, consts, pw;   // pw: [B][32] powers rho^(2^j) for the dot weights
    DBuf<unsigned char> sgn;
    DBuf<Jac> jscratch, res;
    MsmPlan plan;
    int blocks_n = 1, blocks_l = 1;
    size_t shard_lo = 0;            // index of this handle's first norm element inside the whole (sharded) vector
    bppp_dtr* dtr = nullptr;        // not owned: device transcript the round commitments are absorbed into
    bool no_rebase = false;         // a shard of a larger argument keeps folding its own slice (bppp_nl_prove_sharded)
    // tensor mode: no generator folding; per-generator fold coefficients + folded opening scalars.
    // The "original" generators of tensor mode are the list `tgens` of tN + tM (+ g) points: the shared
    // list of the handle, or -- after a single large argument has folded down to a few thousand
    // generators -- a table built over ITS current generators (nl_rebase_tensor); tround counts the rounds
    // since then.
    bool tensor = false;
    bppp_gens* tgens = nullptr;
    bppp_gens* tail_gens = nullptr;     // owned: the re-based list
    size_t tN = 0, tM = 0, tP0 = 0;
    int tround = 0;
    DBuf<u256> coef, fsc;           // coef [B][N+M] (Montgomery); fsc [2][B][P0] (Montgomery)
    // IP argument (kind = BPPP_ARG_IP): w[] holds a (on G' = g1 + r g0), bv[] holds b (on H' = g1 - r g0);
    // Np = ceil(N/2) pairs; coef = [coefG (Np) | coefH (Np) | coefK (M)] per proof;
    // fsc = [fgl | fgr | fhl | fhr (Np each) | fkl | fkr (M each)] per proof
    size_t Np = 0;
    DBuf<u256> bv[2];
    std::vector<Fr> rr, ny;         // basis-change scalar r (q = r^4) and the second normalisation
    // host state (Montgomery, 4 x 64-bit limbs; bit-compatible with the device's u256)
    std::vector<Fr> q, qinv, nn, nl, s, sX, sR;
};

namespace {
enum { C_RHO = 0, C_K1, C_K2, C_AU, C_BU, C_AL, C_BL, C_AC, C_BC, C_COEF /* 8 */, C_KB = C_COEF + 8 /* 2 */,
       C_KA = C_KB + 2 /* 2 */, C_A0N = C_KA + 2, C_B0N, C_A0L, C_B0L, C_AV, C_BV, C_A0H, C_B0H, C_RR, C_NX, C_NY,
       C_Q, C_QINV, C_NN, C_NL, C_S, C_INV /* 2 */, C_COUNT = C_INV + 2 };   // C_Q..C_INV: the device-resident round state (rounds.cuh)
inline u256* cptr(bppp_nl* h, int which) { return h->consts.p + (size_t)which * h->B; }
static_assert(sizeof(Fr) == sizeof(u256), "host and device field elements share one layout");

// blocks of 256 threads per proof: a power of two (so rho^T is a table entry), about 4 pairs per
// thread for long vectors, at most 2^18 threads
int dots_blocks(size_t n_pairs) {
    size_t b = 1;
    while (b * 256 * 4 < n_pairs && b < 1024) b <<= 1;
    return (int)b;
}
int ilog2(size_t x) {
    int l = 0;
    while ((x >> l) > 1) l++;
    return l;
}
u256 fr_canon_u256(const Fr& a) {
    u256 r;
    uint64_t c[4];
    h64::to_canon(c, a);
    memcpy(r.v, c, 32);
    return r;
}
Fr fr_from_mag(const u256& mag, bool neg) {
    uint64_t c[4];
    memcpy(c, mag.v, 32);
    Fr m = h64::from_canon(c);
    return neg ? h64::neg(m) : m;
}

// launch k_fold_dots for norm and linear vectors.  fold = 0: dots of current vectors; fold = 1:
// fold current vectors into the other buffer and compute the dots of the result.
// NL: norm u = v = w (wL*wR, wR^2), linear u = c, v = l (cL*lR + cR*lL, cR*lR).
// IP: norm u = a, v = b (aL*bR, aR*bL), linear (cR*lL, cL*lR)   (InnerProductArgument.hs:70-81,149-152)
int launch_fold_dots(bppp_nl* h, int fold) {
    bppp_ctx* ctx = h->ctx;
    const bool ip = h->kind == BPPP_ARG_IP;
    int src = h->cur, dst = h->cur ^ 1;
    if (h->curN) {
        size_t ny = fold ? (h->curN + 1) / 2 : h->curN;
        h->blocks_n = dots_blocks((ny + 1) / 2);
        CK(h->part_n.ensure(h->B * h->blocks_n * 2));
        FoldDotsArgs A;
        A.u = h->w[src].p; A.uo = h->w[dst].p;
        A.v = ip ? h->bv[src].p : h->w[src].p; A.vo = ip ? h->bv[dst].p : h->w[dst].p;
        A.in_stride = h->wstride[src]; A.out_stride = h->wstride[dst];
        A.n_in = (int)h->curN; A.fold = fold;
        A.au = cptr(h, C_AU); A.bu = cptr(h, C_BU);
        A.av = ip ? cptr(h, C_AV) : cptr(h, C_AU); A.bv = ip ? cptr(h, C_BV) : cptr(h, C_BU);
        A.rho = cptr(h, C_RHO); A.pw = h->pw.p; A.log2T = ilog2((size_t)h->blocks_n * 256);
        A.m1 = 1; A.m2 = ip ? 2 : 4; A.partial = h->part_n.p;
        { ProfScope ps_(ctx, K_POW_TABLE, 0);
        k_pow_table<<<(unsigned)((h->B + 127) / 128), 128, 0, ctx->st>>>(cptr(h, C_RHO), h->pw.p, (int)h->B);
        }
        CK(cudaGetLastError());
        g_work = (ip ? 2 : 1) * 32.0 * (double)h->B * (fold ? (double)(h->curN + ny) : (double)h->curN);
        { ProfScope ps_(ctx, K_FOLD_DOTS, g_work);
        k_fold_dots<<<dim3(h->blocks_n, (unsigned)h->B), 256, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    }
    if (h->curM) {
        size_t ny = fold ? (h->curM + 1) / 2 : h->curM;
        h->blocks_l = dots_blocks((ny + 1) / 2);
        CK(h->part_l.ensure(h->B * h->blocks_l * 2));
        FoldDotsArgs A;
        A.u = h->c[src].p; A.v = h->l[src].p; A.uo = h->c[dst].p; A.vo = h->l[dst].p;
        A.in_stride = h->lstride[src]; A.out_stride = h->lstride[dst];
        A.n_in = (int)h->curM; A.fold = fold;
        A.au = cptr(h, C_AC); A.bu = cptr(h, C_BC); A.av = cptr(h, C_AL); A.bv = cptr(h, C_BL);
        A.rho = nullptr; A.pw = nullptr; A.log2T = 0; A.m1 = ip ? 2 : 3; A.m2 = ip ? 1 : 4; A.partial = h->part_l.p;
        g_work = 2 * 32.0 * (double)h->B * (fold ? (double)(h->curM + ny) : (double)h->curM);
        { ProfScope ps_(ctx, K_FOLD_DOTS, g_work);
        k_fold_dots<<<dim3(h->blocks_l, (unsigned)h->B), 256, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    }
    return BPPP_OK;
}
int upload_consts(bppp_nl* h, int which, const std::vector<Fr>& v) {
    bppp_ctx* ctx = h->ctx;
    CK(H2D(cptr(h, which), v.data(), v.size() * 32));
    return BPPP_OK;
}
}  // namespace
// =============================================================================== IP argument
// IP.NormLinear (src/Bulletproof/InnerProductArgument.hs) on the device, tensor style.
namespace {
inline u256* ip_coef(bppp_nl* h) { return h->coef.p; }
inline size_t ip_coef_stride(bppp_nl* h) { return 2 * h->Np + h->M; }
inline u256* ip_f(bppp_nl* h, int which) {          // 0 fgl, 1 fgr, 2 fhl, 3 fhr : [B][Np];  4 fkl, 5 fkr : [B][M]
    return which < 4 ? h->fsc.p + (size_t)which * h->B * h->Np : h->fsc.p + 4 * h->B * h->Np + (size_t)(which - 4) * h->B * h->M;
}

int ip_create(bppp_nl* h, const uint8_t* q, const uint8_t* s, const uint8_t* w, const uint8_t* l, const uint8_t* c) {
    bppp_ctx* ctx = h->ctx;
    const size_t B = h->B, N = h->N, M = h->M, Np = (N + 1) / 2;
    h->Np = Np;
    h->tensor = true;
    const size_t Np2 = (Np + 1) / 2;
    CK(h->aff.alloc(B * 2));
    CK(h->w[0].alloc(B * Np)); CK(h->w[1].alloc(B * Np2));
    CK(h->bv[0].alloc(B * Np)); CK(h->bv[1].alloc(B * Np2));
    CK(h->l[0].alloc(B * M)); CK(h->l[1].alloc(B * h->M2));
    CK(h->c[0].alloc(B * M)); CK(h->c[1].alloc(B * h->M2));
    h->wstride[0] = Np; h->wstride[1] = Np2; h->lstride[0] = M; h->lstride[1] = h->M2;
    CK(h->sc.alloc(2 * B * h->P0));
    CK(h->dots.alloc(B * 2));
    CK(h->pw.alloc(B * 32));
    CK(h->consts.alloc((size_t)C_COUNT * B));
    CK(h->res.alloc(B * 2));
    CK(h->coef.alloc(B * (2 * Np + M)));
    CK(h->fsc.alloc(B * (4 * Np + 2 * M)));
    // host state: the setup's `q` is the r with q = r^4 (makeNorm, :194-197)
    h->rr.resize(B); h->q.resize(B); h->qinv.resize(B); h->nn.assign(B, h64::one()); h->ny.assign(B, h64::one());
    h->nl.assign(B, h64::one()); h->s.resize(B); h->sX.resize(B); h->sR.resize(B);
    std::vector<Fr> r2inv(B), half(B);
    const Fr two = h64::from_u64(2);
    for (size_t b = 0; b < B; b++) {
        h->rr[b] = h64::from_bytes(q + 32 * b);
        Fr r2 = h64::sqr(h->rr[b]);
        h->q[b] = h64::sqr(r2);
        h->qinv[b] = h->q[b];
        h->s[b] = h64::from_bytes(s + 32 * b);
        r2inv[b] = h64::mul(two, h->rr[b]);
        half[b] = two;
    }
    h64::batch_inv(h->qinv.data(), B);
    h64::batch_inv(r2inv.data(), B);
    h64::batch_inv(half.data(), B);
    CK(H2D(cptr(h, C_AU), r2inv.data(), B * 32));
    CK(H2D(cptr(h, C_BU), half.data(), B * 32));
    CK(H2D(cptr(h, C_RR), h->rr.data(), B * 32));
    // witness -> Montgomery, then the basis change of the norm part
    DBuf<u256> wm;
    CK(wm.alloc(B * N));
    struct { const uint8_t* src; u256* dst; size_t n; } up[3] = {{w, wm.p, B * N}, {l, h->l[0].p, B * M}, {c, h->c[0].p, B * M}};
    for (auto& u : up) {
        if (!u.n) continue;
        CK(H2D(h->sc.p, u.src, u.n * 32));
        { ProfScope ps_(ctx, K_FR_CONVERT, 0);
        k_fr_convert<<<(unsigned)((u.n + 255) / 256), 256, 0, ctx->st>>>(h->sc.p, u.dst, u.n, 1);
        }
        CK(cudaGetLastError());
    }
    if (N) {
        { ProfScope ps_(ctx, K_IP_MISC, 0);
        k_ip_make_norm<<<dim3((unsigned)((Np + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(
            wm.p, N, (int)N, h->w[0].p, h->bv[0].p, Np, cptr(h, C_AU), cptr(h, C_BU));
        }
        CK(cudaGetLastError());
    }
    {
        size_t nc = B * (2 * Np + M);
        ProfScope ps_(ctx, K_IP_MISC, 0);
        k_fill_one<<<(unsigned)((nc + 255) / 256), 256, 0, ctx->st>>>(h->coef.p, nc);
        CK(cudaGetLastError());
    }
    CK(cudaMemsetAsync(h->sc.p, 0, 2 * B * h->P0 * 32, ctx->st));
    h->curN = Np;
    CK(ctx_sync(ctx));
    return BPPP_OK;
}

int ip_round_commit(bppp_nl* h, uint8_t* Lout, uint8_t* Rout) {
    bppp_ctx* ctx = h->ctx;
    const size_t B = h->B, P0 = h->P0, Np = h->Np, M = h->M;
    // sL = s q nx ny sum (q^2)^t aL bR ; sR = s q^2 nx ny sum (q^2)^t aR bL   (:70-81), s = 4 (:195)
    std::vector<Fr> rho(B), k1(B), k2(B), coef(B * 8, h64::zero());
    const Fr four = h64::from_u64(4);
    host_parallel_for(B, [&](size_t b) {
        Fr q2 = h64::sqr(h->q[b]);
        Fr k = h64::mul(four, h64::mul(h->nn[b], h->ny[b]));
        rho[b] = q2;
        k1[b] = h64::mul(k, h->q[b]);
        k2[b] = h64::mul(k, q2);
        coef[b * 8 + 2] = h->qinv[b];       // L on gR <- q^-1 * aL
        coef[b * 8 + 5] = h->q[b];          // R on gL <- q * aR
    });
    int rc;
    if (!h->have_partials) {
        if ((rc = upload_consts(h, C_RHO, rho))) return rc;
        if ((rc = launch_fold_dots(h, 0))) return rc;
        h->have_partials = true;
    }
    if ((rc = upload_consts(h, C_K1, k1)) || (rc = upload_consts(h, C_K2, k2))) return rc;
    CK(H2D(cptr(h, C_COEF), coef.data(), B * 8 * 32));
    u256* ls = h->sc.p;
    u256* rs = h->sc.p + B * P0;
    {
        DotsFinishArgs A;
        memset(&A, 0, sizeof A);
        if (h->curN) { A.partial[A.n_seg] = h->part_n.p; A.n_blocks[A.n_seg] = h->blocks_n; A.k1[A.n_seg] = cptr(h, C_K1); A.k2[A.n_seg] = cptr(h, C_K2); A.n_seg++; }
        if (h->curM) { A.partial[A.n_seg] = h->part_l.p; A.n_blocks[A.n_seg] = h->blocks_l; A.n_seg++; }
        A.res = h->dots.p; A.xs = ls; A.rs = rs; A.sc_stride = P0; A.batch = (int)B;
        { ProfScope ps_(ctx, K_DOTS_FINISH, 0);
        k_dots_finish<<<(unsigned)((B + 127) / 128), 128, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    }
    const int src = h->cur;
    struct { const u256* x; size_t stride; size_t n; u256* o0; u256* o1; size_t ostride; unsigned char kd[8]; } jobs[3] = {
        // a on G':  L: gR <- qInv*aL ; R: gL <- q*aR          (:78-80)
        {h->w[src].p, h->wstride[src], h->curN, ip_f(h, 0), ip_f(h, 1), Np, {0, 0, 2, 0, 0, 2, 0, 0}},
        // b on H':  L: hL <- bR ; R: hR <- bL
        {h->bv[src].p, h->wstride[src], h->curN, ip_f(h, 2), ip_f(h, 3), Np, {0, 1, 0, 0, 0, 0, 1, 0}},
        // l on K:   L: kR <- lL ; R: kL <- lR                  (:149-152)
        {h->l[src].p, h->lstride[src], h->curM, ip_f(h, 4), ip_f(h, 5), M, {0, 0, 1, 0, 0, 1, 0, 0}}};
    for (auto& j : jobs) {
        if (!j.n) continue;
        MsmScalarsArgs A;
        A.x = j.x; A.in_stride = j.stride; A.n_in = (int)j.n;
        A.xs = j.o0; A.rs = j.o1; A.mont_out = 1; A.sc_stride = j.ostride; A.off = 0; A.coef = cptr(h, C_COEF);
        memcpy(A.kind, j.kd, 8);
        { ProfScope ps_(ctx, K_MSM_SCALARS, 0);
        k_msm_scalars<<<dim3((unsigned)(((j.n + 1) / 2 + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    }
    if (h->N) {
        IpExpandArgs A;
        A.fgl = ip_f(h, 0); A.fgr = ip_f(h, 1); A.fhl = ip_f(h, 2); A.fhr = ip_f(h, 3); A.f_stride = Np;
        A.cg = ip_coef(h); A.ch = ip_coef(h) + Np; A.c_stride = ip_coef_stride(h);
        A.r = cptr(h, C_RR); A.ls = ls; A.rs = rs; A.sc_stride = P0; A.off = 1; A.n = (int)h->N; A.shift = h->round;
        { ProfScope ps_(ctx, K_EXPAND, 0);
        k_ip_expand<<<dim3((unsigned)((Np + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    }
    if (M) {
        ExpandArgs A;
        A.fx = ip_f(h, 4); A.fr_ = ip_f(h, 5); A.f_stride = M; A.f_off = 0;
        A.coef = ip_coef(h); A.coef_stride = ip_coef_stride(h); A.coef_off = (int)(2 * Np);
        A.xs = ls; A.rs = rs; A.sc_stride = P0; A.off = 1 + (int)h->N; A.n = (int)M; A.shift = h->round;
        { ProfScope ps_(ctx, K_EXPAND, 0);
        k_expand_scalars<<<dim3((unsigned)((M + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    }
    if ((rc = run_msm_gens(h->gens, P0, ls, P0, B * P0, B, 2, h->res.p, 2 * msm_alg_imads((double)P0)))) return rc;
    if ((rc = to_affine(ctx, h->res.p, 1, h->aff.p, 1, 0, 1, B * 2))) return rc;
    std::vector<Affine> xr(B * 2);
    std::vector<Fr> dots(B * 2);
    CK(D2H(xr.data(), h->aff.p, B * 2 * 64));
    CK(D2H(dots.data(), h->dots.p, B * 2 * 32));
    CK(ctx_sync(ctx));
    for (size_t b = 0; b < B; b++) {
        memcpy(Lout + 64 * b, &xr[2 * b], 64);
        memcpy(Rout + 64 * b, &xr[2 * b + 1], 64);
        h->sX[b] = dots[2 * b];
        h->sR[b] = dots[2 * b + 1];
    }
    return BPPP_OK;
}

int ip_round_fold(bppp_nl* h, const uint8_t* e) {
    bppp_ctx* ctx = h->ctx;
    const size_t B = h->B, Np = h->Np, M = h->M;
    std::vector<Fr> em(B), ei(B);
    for (size_t b = 0; b < B; b++) { em[b] = h64::from_bytes(e + 32 * b); ei[b] = em[b]; }
    h64::batch_inv(ei.data(), B);
    std::vector<Fr> aG(B), bG(B), cH(B), dH(B), aK(B), bK(B), inv(3 * B), rho(B);
    host_parallel_for(B, [&](size_t b) {
        // :89 (a', b') = rationalReduceScalar (qInv * eInv) ; :93 (c', d') = rationalReduceScalar e ;
        // linear :158 rationalReduceScalar (recip e)
        host::Ratio rg = host::rational_reduce(fr_canon_u256(h64::mul(h->qinv[b], ei[b])));
        host::Ratio rh = host::rational_reduce(host::from_bytes(e + 32 * b));
        host::Ratio rk = host::rational_reduce(fr_canon_u256(ei[b]));
        aG[b] = fr_from_mag(rg.a, rg.a_neg); bG[b] = fr_from_mag(rg.b, rg.b_neg);
        cH[b] = fr_from_mag(rh.a, rh.a_neg); dH[b] = fr_from_mag(rh.b, rh.b_neg);
        aK[b] = fr_from_mag(rk.a, rk.a_neg); bK[b] = fr_from_mag(rk.b, rk.b_neg);
        inv[b] = bG[b]; inv[B + b] = dH[b]; inv[2 * B + b] = bK[b];
    });
    h64::batch_inv(inv.data(), 3 * B);
    std::vector<Fr> au(B), bu(B), av(B), bvv(B), al(B), bl(B);
    host_parallel_for(B, [&](size_t b) {
        au[b] = inv[b];                                            // x' = b0Inv (xL + e q xR)      (:97)
        bu[b] = h64::mul(h64::mul(em[b], h->q[b]), inv[b]);
        av[b] = inv[B + b];                                        // y' = d0Inv (yL + eInv yR)
        bvv[b] = h64::mul(ei[b], inv[B + b]);
        al[b] = inv[2 * B + b];                                    // l' = b0Inv xL + e b0Inv xR    (:165)
        bl[b] = h64::mul(em[b], inv[2 * B + b]);
        // s' = s + eInv sL + e sR   (makeEs = (recip e, e), :68)
        h->s[b] = h64::add(h->s[b], h64::add(h64::mul(ei[b], h->sX[b]), h64::mul(em[b], h->sR[b])));
        h->nn[b] = h64::mul(h64::mul(h->nn[b], bG[b]), h->qinv[b]);   // nx * b0 * qInv
        h->ny[b] = h64::mul(h->ny[b], dH[b]);
        h->nl[b] = h64::mul(h->nl[b], bK[b]);
        h->q[b] = h64::sqr(h->q[b]);
        h->qinv[b] = h64::sqr(h->qinv[b]);
        rho[b] = h64::sqr(h->q[b]);
    });
    int rc;
    if ((rc = upload_consts(h, C_AU, au)) || (rc = upload_consts(h, C_BU, bu)) || (rc = upload_consts(h, C_AV, av)) ||
        (rc = upload_consts(h, C_BV, bvv)) || (rc = upload_consts(h, C_AL, al)) || (rc = upload_consts(h, C_BL, bl)) ||
        (rc = upload_consts(h, C_AC, bK)) || (rc = upload_consts(h, C_BC, aK)) || (rc = upload_consts(h, C_RHO, rho)) ||
        (rc = upload_consts(h, C_A0N, aG)) || (rc = upload_consts(h, C_B0N, bG)) || (rc = upload_consts(h, C_A0H, cH)) ||
        (rc = upload_consts(h, C_B0H, dH)) || (rc = upload_consts(h, C_A0L, aK)) || (rc = upload_consts(h, C_B0L, bK)))
        return rc;
    if ((rc = launch_fold_dots(h, 1))) return rc;
    struct { int off; size_t n; int ca, cb; } segs[3] = {{0, Np, C_A0N, C_B0N}, {(int)Np, Np, C_A0H, C_B0H}, {(int)(2 * Np), M, C_A0L, C_B0L}};
    for (auto& sg : segs) {
        if (!sg.n) continue;
        { ProfScope ps_(ctx, K_COEF, 0);
        k_coef_update<<<dim3((unsigned)((sg.n + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(
            ip_coef(h), ip_coef_stride(h), sg.off, (int)sg.n, h->round, cptr(h, sg.ca), cptr(h, sg.cb), 1, 0);
        }
        CK(cudaGetLastError());
    }
    h->cur ^= 1;
    h->curN = (h->curN + 1) / 2;
    h->curM = (h->curM + 1) / 2;
    h->round++;
    CK(ctx_sync(ctx));
    return BPPP_OK;
}

int ip_final(bppp_nl* h, uint8_t* s, uint8_t* w, uint8_t* l) {
    bppp_ctx* ctx = h->ctx;
    const size_t B = h->B, cn = h->curN, cl = h->curM;
    if (s) for (size_t b = 0; b < B; b++) h64::to_bytes(s + 32 * b, h->s[b]);
    if (w && cn) {
        DBuf<u256> out;
        CK(out.alloc(B * cn * 2));
        CK(H2D(cptr(h, C_NX), h->nn.data(), B * 32));
        CK(H2D(cptr(h, C_NY), h->ny.data(), B * 32));
        { ProfScope ps_(ctx, K_IP_MISC, 0);
        k_ip_final<<<dim3((unsigned)((cn + 127) / 128), (unsigned)B), 128, 0, ctx->st>>>(
            h->w[h->cur].p, h->bv[h->cur].p, h->wstride[h->cur], (int)cn, cptr(h, C_NX), cptr(h, C_NY), out.p);
        }
        CK(cudaGetLastError());
        CK(D2H(w, out.p, B * cn * 2 * 32));
        CK(ctx_sync(ctx));
    }
    if (l && cl) {
        std::vector<Fr> hl(B * cl);
        ctx->d2h += B * cl * 32;
        CK(cudaMemcpy2DAsync(hl.data(), cl * 32, h->l[h->cur].p, h->lstride[h->cur] * 32, cl * 32, B, cudaMemcpyDeviceToHost, ctx->st));
        CK(ctx_sync(ctx));
        for (size_t b = 0; b < B; b++)
            for (size_t i = 0; i < cl; i++) h64::to_bytes(l + 32 * (b * cl + i), h64::mul(h->nl[b], hl[b * cl + i]));
    }
    return BPPP_OK;
}
}  // namespace

namespace {
int nl_create_impl(bppp_gens* gens, bool own, int kind, size_t batch, const uint8_t* q, const uint8_t* s, const uint8_t* w,
                   const uint8_t* l, const uint8_t* c, bppp_nl** out, const u256* w_dev = nullptr, const u256* l_dev = nullptr,
                   const u256* c_dev = nullptr) {
    bppp_ctx* ctx = gens->ctx;
    const size_t N = gens->N, M = gens->M;
    if (N + M + 1 > 0x7fffffff) FAIL(BPPP_ERR_ARG, "vector too long");
    if (!check_fr(q, batch) || !check_fr(s, batch) || (!w_dev && !check_fr(w, batch * N)) || (!l_dev && !check_fr(l, batch * M)) ||
        (!c_dev && !check_fr(c, batch * M)))
        FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    bppp_nl* h = new bppp_nl();
    h->ctx = ctx; h->gens = gens; h->own_gens = own; h->kind = kind;
    h->tgens = gens; h->tN = N; h->tM = M; h->tP0 = 1 + N + M; h->tround = 0;
    h->B = batch; h->N = N; h->M = M; h->curN = N; h->curM = M;
    h->N2 = (N + 1) / 2; h->M2 = (M + 1) / 2; h->P0 = 1 + N + M; h->P2 = 1 + h->N2 + h->M2;
    auto fail = [&](int rc) { h->own_gens = false; bppp_nl_destroy(h); return rc; };
    if (kind == BPPP_ARG_IP) {
        if (!gens->tbl.p) { ctx->err = "IP argument: the generator list is too long for the window table"; return fail(BPPP_ERR_ARG); }
        int rc = ip_create(h, q, s, w, l, c);
        if (rc) return fail(rc);
        *out = h;
        return BPPP_OK;
    }
#define CKH(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->err = std::string(#x ": ") + cudaGetErrorString(e_); return fail(BPPP_ERR_CUDA); } } while (0)
    {
        const char* ev = getenv("BPPP_ROUND_MODE");          // fold | tensor | auto (default)
        if (ev && !strcmp(ev, "fold")) h->tensor = false;
        else if (ev && !strcmp(ev, "tensor")) h->tensor = true;
        else h->tensor = (h->P0 <= GT_TABLE_MAX_TERMS);       // few rounds: k fixed-base MSMs beat fold + bucket MSMs
        if (!gens->tbl.p) h->tensor = false;                  // tensor mode needs the window table
    }
    if (h->tensor) {
        CKH(h->coef.alloc(batch * (N + M)));
        CKH(h->fsc.alloc(2 * batch * h->P0));
    } else {
        CKH(h->pts[0].alloc(batch * h->P2)); CKH(h->pts[1].alloc(batch * h->P2));
        CKH(h->jscratch.alloc(batch * (h->N2 + h->M2)));
    }
    CKH(h->aff.alloc(batch * 2));
    CKH(h->w[0].alloc(batch * N)); CKH(h->w[1].alloc(batch * h->N2));
    CKH(h->l[0].alloc(batch * M)); CKH(h->l[1].alloc(batch * h->M2));
    CKH(h->c[0].alloc(batch * M)); CKH(h->c[1].alloc(batch * h->M2));
    h->wstride[0] = N; h->wstride[1] = h->N2; h->lstride[0] = M; h->lstride[1] = h->M2;
    CKH(h->sc.alloc(2 * batch * h->P0));
    CKH(h->dots.alloc(batch * 2));
    CKH(h->pw.alloc(batch * 32));
    CKH(h->consts.alloc((size_t)C_COUNT * batch));
    CKH(h->sgn.alloc(batch * 2));
    CKH(h->res.alloc(batch * 2));
    for (int k = 0; k < 2 && !h->tensor; k++) {
        { ProfScope ps_(ctx, K_BCAST, 0);
        k_bcast_point<<<(unsigned)((batch + 127) / 128), 128, 0, ctx->st>>>(gens->base.p, h->pts[k].p, h->P2, batch);
        }
        CKH(cudaGetLastError());
    }
    // scalars -> Montgomery on the device (staged through the scalar scratch buffer)
    struct { const uint8_t* src; u256* dst; size_t n; } up[3] = {
        {w, h->w[0].p, batch * N}, {l, h->l[0].p, batch * M}, {c, h->c[0].p, batch * M}};
    if (w_dev) {                    // the witness was produced on the device (Montgomery form already)
        CKH(cudaMemcpyAsync(h->w[0].p, w_dev, batch * N * 32, cudaMemcpyDeviceToDevice, ctx->st));
        up[0].n = 0;
    }
    if (l_dev && M) { CKH(cudaMemcpyAsync(h->l[0].p, l_dev, batch * M * 32, cudaMemcpyDeviceToDevice, ctx->st)); up[1].n = 0; }
    if (c_dev && M) { CKH(cudaMemcpyAsync(h->c[0].p, c_dev, batch * M * 32, cudaMemcpyDeviceToDevice, ctx->st)); up[2].n = 0; }
    for (auto& u : up) {
        if (!u.n) continue;
        CKH(H2D(h->sc.p, u.src, u.n * 32));
        { ProfScope ps_(ctx, K_FR_CONVERT, 0);
        k_fr_convert<<<(unsigned)((u.n + 255) / 256), 256, 0, ctx->st>>>(h->sc.p, u.dst, u.n, 1);
        }
        CKH(cudaGetLastError());
    }
    CKH(cudaMemsetAsync(h->sc.p, 0, 2 * batch * h->P0 * 32, ctx->st));
    h->q.resize(batch); h->qinv.resize(batch); h->nn.assign(batch, h64::one()); h->nl.assign(batch, h64::one());
    h->s.resize(batch); h->sX.resize(batch); h->sR.resize(batch);
    host_parallel_for(batch, [&](size_t b) {
        h->q[b] = h64::from_bytes(q + 32 * b);
        h->qinv[b] = h->q[b];
        h->s[b] = h64::from_bytes(s + 32 * b);
    });
    h64::batch_inv(h->qinv.data(), batch);
    CKH(ctx_sync(ctx));
#undef CKH
    *out = h;
    return BPPP_OK;
}
}  // namespace

extern "C" int bppp_nl_create_gens(bppp_gens* gens, int kind, size_t batch, const uint8_t* q, const uint8_t* s,
                                   const uint8_t* w, const uint8_t* l, const uint8_t* c, bppp_nl** out) {
    if (!gens) return BPPP_ERR_ARG;
    bppp_ctx* ctx = gens->ctx;
    if (!out || !q || !s || batch == 0 || (gens->N && !w) || (gens->M && (!l || !c)))
        FAIL(BPPP_ERR_ARG, "bppp_nl_create: null/empty argument");
    *out = nullptr;
    if (kind != BPPP_ARG_NL && kind != BPPP_ARG_IP) FAIL(BPPP_ERR_ARG, "bppp_nl_create: unknown argument kind");
    ENTER(ctx);
    return nl_create_impl(gens, false, kind, batch, q, s, w, l, c, out);
}
extern "C" int bppp_nl_create(bppp_ctx* ctx, int kind, size_t batch, size_t N, size_t M, const uint8_t* g,
                              const uint8_t* G, const uint8_t* H, const uint8_t* q, const uint8_t* s, const uint8_t* w,
                              const uint8_t* l, const uint8_t* c, bppp_nl** out) {
    if (!ctx) return BPPP_ERR_ARG;
    if (!out || !g || !q || !s || batch == 0 || (N && (!G || !w)) || (M && (!H || !l || !c)))
        FAIL(BPPP_ERR_ARG, "bppp_nl_create: null/empty argument");
    *out = nullptr;
    if (kind != BPPP_ARG_NL && kind != BPPP_ARG_IP) FAIL(BPPP_ERR_ARG, "bppp_nl_create: unknown argument kind");
    bppp_gens* gens = nullptr;
    int rc = bppp_gens_create(ctx, N, M, g, G, H, &gens);
    if (rc) return rc;
    rc = nl_create_impl(gens, true, kind, batch, q, s, w, l, c, out);
    if (rc) bppp_gens_destroy(gens);
    return rc;
}

// =============================================================================== TypedReciprocal scalar phases
// Device side of proveTRRPM's phases 1-4 for a batch of proofs on one lane (see kernels.cuh).
struct bppp_trrp {
    bppp_gens* gens;
    size_t n_ent, n_ranges, n_bases;
    DBuf<TrrpEnt> ent;
    DBuf<u256> eb, es;                 // static per-entry b, s (Montgomery)
    size_t B = 0;
    DBuf<u256> scA, scR, scBL, amounts, chal2, chal3, chal4, chalv, xp, vt, r, c, bl, w, small;
    DBuf<Jac> res;
    DBuf<Affine> aff;
    int phase = 0;
    // device transcript (bppp_trrp_set_transcript): the commitments of every phase are absorbed where they are
    // produced and the challenges come back with them
    int tr_fmt = -1;
    bppp_dtr* tr = nullptr;
    uint8_t* chal_inv_out = nullptr;   // host buffer the next _tr call also fills with the inverted challenges (one-shot)
    size_t n_shared = 0;               // shared-multiplicity coefficient slots (bppp_trrp_set_shared)
    DBuf<int> sh_bidx;
    DBuf<u256> sh_sym, sh_out;
    DBuf<unsigned char> seeds, seed_len;
    DBuf<unsigned long long> n0s;
};
namespace {
TrrpStatic trrp_static(bppp_trrp* h) {
    TrrpStatic st;
    st.ent = h->ent.p; st.b = h->eb.p; st.s = h->es.p;
    st.n_ent = (int)h->n_ent; st.n_ranges = (int)h->n_ranges; st.n_bases = (int)h->n_bases;
    return st;
}
// MSM of `n_msm` scalar rows of length P0 (device resident, canonical) over the lane's generators -> affine, host.
// With chal_out: the batch's `per` commitments per proof (followed by `n_extra` host-resident ones per proof) are
// absorbed by the device transcript as ONE oracle call and the first `count` challenges returned, [batch][count].
int trrp_commit(bppp_trrp* h, const u256* sc, size_t n_msm, uint8_t* out, double scalar_bits = 256.0, uint8_t* chal_out = nullptr,
                int count = 0, size_t per = 1, const uint8_t* extra = nullptr, size_t n_extra = 0) {
    bppp_ctx* ctx = h->gens->ctx;
    uint8_t* inv_out = h->chal_inv_out;          // one-shot request: consumed by this call whatever its outcome
    h->chal_inv_out = nullptr;
    const size_t P0 = h->gens->P0;
    CK(h->res.ensure(n_msm)); CK(h->aff.ensure(n_msm));
    int rc = run_msm_gens(h->gens, P0, sc, P0, 0, n_msm, 1, h->res.p, msm_alg_imads((double)P0, scalar_bits));
    if (rc) return rc;
    if ((rc = to_affine(ctx, h->res.p, 1, h->aff.p, 1, 0, 1, n_msm))) return rc;
    CK(D2H(out, h->aff.p, n_msm * 64));
    if (chal_out) {
        bppp_dtr* t = h->tr;
        if (!t) FAIL(BPPP_ERR_STATE, "device transcript not enabled (bppp_trrp_set_transcript)");
        const Affine* src = h->aff.p;
        if (n_extra) {                  // rows [per device points | n_extra host points] staged side by side
            const size_t row = per + n_extra, off = t->B * (size_t)t->ncoms[t->n_calls];
            if (t->ncoms[t->n_calls] + row > t->cap) FAIL(BPPP_ERR_STATE, "device transcript: capacity exceeded");
            CK(t->stage.ensure(t->B * t->cap));
            Affine* st = t->stage.p + off;
            CK(cudaMemcpy2DAsync(st, row * 64, h->aff.p, per * 64, per * 64, t->B, cudaMemcpyDeviceToDevice, ctx->st));
            ctx->h2d += t->B * n_extra * 64;
            CK(cudaMemcpy2DAsync(st + per, row * 64, extra, n_extra * 64, n_extra * 64, t->B, cudaMemcpyHostToDevice, ctx->st));
            src = st;
            per = row;
        }
        if ((rc = dtr_absorb_dev(t, src, per, per))) return rc;
        if ((rc = dtr_squeeze_first(t, count))) return rc;
        CK(D2H(chal_out, t->chal.p, t->B * (size_t)count * 32));
        if (inv_out) {
            if ((rc = dtr_invert_dev(t, count))) return rc;
            CK(D2H(inv_out, t->chal_inv.p, t->B * (size_t)count * 32));
        }
    }
    CK(ctx_sync(ctx));
    return BPPP_OK;
}
}  // namespace
// one-shot: the next bppp_trrp_*_tr call also writes 1 / challenge for each of its challenges to `out` (same layout)
extern "C" int bppp_trrp_want_inverses(bppp_trrp* h, uint8_t* out) {
    if (!h) return BPPP_ERR_ARG;
    h->chal_inv_out = out;
    return BPPP_OK;
}
// makeSharedCoeffs on the device (TypedReciprocal.hs:204-206).  Static part, once per setup: slot i belongs to shared
// base number base_idx[i] (index into the setup's sorted base list) and symbol sym[i] (canonical scalar).
extern "C" int bppp_trrp_set_shared(bppp_trrp* h, size_t n_slots, const int32_t* base_idx, const uint8_t* sym) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->gens->ctx;
    if (n_slots && (!base_idx || !sym)) FAIL(BPPP_ERR_ARG, "bppp_trrp_set_shared: null argument");
    ENTER(ctx);
    for (size_t i = 0; i < n_slots; i++)
        if (base_idx[i] < 0 || (size_t)base_idx[i] >= h->n_bases) FAIL(BPPP_ERR_ARG, "bppp_trrp_set_shared: base index out of range");
    if (!check_fr(sym, n_slots)) FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    h->n_shared = n_slots;
    if (!n_slots) return BPPP_OK;
    CK(h->sh_bidx.alloc(n_slots)); CK(h->sh_sym.alloc(n_slots));
    CK(cudaMemcpyAsync(h->sh_bidx.p, base_idx, n_slots * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(h->sh_sym.p, sym, n_slots * 32, cudaMemcpyHostToDevice, ctx->st));
    k_fr_convert<<<(unsigned)((n_slots + 255) / 256), 256, 0, ctx->st>>>(h->sh_sym.p, h->sh_sym.p, n_slots, 1);
    CK(cudaGetLastError());
    CK(ctx_sync(ctx));
    return BPPP_OK;
}
// Per batch: out[b][i] = base power of slot i * (1/e - 1/(e + sym_i)) for the challenges e of the last bppp_trrp_phase2*
// (prover) or bppp_trrp_verify_pub (verifier) call.  montgomery != 0: residues times 2^256 mod r (the library's
// internal form, what the host layer multiplies with); else canonical scalars.
extern "C" int bppp_trrp_shared_coeffs(bppp_trrp* h, int montgomery, uint8_t* out) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->gens->ctx;
    if (!out) FAIL(BPPP_ERR_ARG, "bppp_trrp_shared_coeffs: null output");
    if (!h->n_shared) FAIL(BPPP_ERR_STATE, "bppp_trrp_shared_coeffs: bppp_trrp_set_shared first");
    if (h->phase != 2 && h->phase != 10) FAIL(BPPP_ERR_STATE, "bppp_trrp_shared_coeffs: after bppp_trrp_phase2 or bppp_trrp_verify_pub");
    ENTER(ctx);
    const size_t B = h->B, S = h->n_shared;
    const u256* chal = h->phase == 2 ? h->chal2.p : h->chalv.p;
    const int stride = h->phase == 2 ? 4 : 8;
    CK(h->sh_out.ensure(B * S));
    const int per = (int)((S + 255) / 256);
    { ProfScope ps_(ctx, K_TRRP, 0);
    k_trrp_shared<<<(unsigned)(per * B), 256, 0, ctx->st>>>(chal, stride, h->vt.p, (int)h->n_bases, h->sh_bidx.p, h->sh_sym.p, (int)S, (int)B, h->sh_out.p);
    }
    CK(cudaGetLastError());
    if (!montgomery) {
        { ProfScope ps_(ctx, K_FR_CONVERT, 0);
        k_fr_convert<<<(unsigned)((B * S + 255) / 256), 256, 0, ctx->st>>>(h->sh_out.p, h->sh_out.p, B * S, 0);
        }
        CK(cudaGetLastError());
    }
    CK(D2H(out, h->sh_out.p, B * S * 32));
    CK(ctx_sync(ctx));
    return BPPP_OK;
}

extern "C" int bppp_trrp_create(bppp_gens* gens, size_t n_entries, const uint8_t* ent_desc, const uint8_t* ent_b,
                                const uint8_t* ent_s, size_t n_ranges, size_t n_bases, bppp_trrp** out) {
    if (!gens) return BPPP_ERR_ARG;
    bppp_ctx* ctx = gens->ctx;
    if (!out || !ent_desc || !ent_b || !ent_s || n_entries == 0 || n_entries != gens->N || n_ranges == 0)
        FAIL(BPPP_ERR_ARG, "bppp_trrp_create: bad argument");
    *out = nullptr;
    ENTER(ctx);
    if (!check_fr(ent_b, n_entries) || !check_fr(ent_s, n_entries)) FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    for (size_t i = 0; i < n_entries; i++) {
        TrrpEnt e;
        memcpy(&e, ent_desc + 16 * i, 16);
        if (e.ind < 0 || (size_t)e.ind >= n_ranges || (!(e.flags & TE_T) && (e.base_idx < 0 || (size_t)e.base_idx >= n_bases)))
            FAIL(BPPP_ERR_ARG, "bppp_trrp_create: entry descriptor out of range");
    }
    bppp_trrp* h = new bppp_trrp();
    h->gens = gens; h->n_ent = n_entries; h->n_ranges = n_ranges; h->n_bases = n_bases ? n_bases : 1;
    auto fail = [&](cudaError_t e) { ctx->err = std::string("bppp_trrp_create: ") + cudaGetErrorString(e); delete h; return BPPP_ERR_CUDA; };
    cudaError_t e;
    if ((e = h->ent.alloc(n_entries)) || (e = h->eb.alloc(n_entries)) || (e = h->es.alloc(n_entries)) || (e = h->small.alloc(2 * n_entries)))
        return fail(e);
    if ((e = cudaMemcpyAsync(h->ent.p, ent_desc, n_entries * 16, cudaMemcpyHostToDevice, ctx->st))) return fail(e);
    if ((e = cudaMemcpyAsync(h->small.p, ent_b, n_entries * 32, cudaMemcpyHostToDevice, ctx->st))) return fail(e);
    if ((e = cudaMemcpyAsync(h->small.p + n_entries, ent_s, n_entries * 32, cudaMemcpyHostToDevice, ctx->st))) return fail(e);
    k_fr_convert<<<(unsigned)((n_entries + 255) / 256), 256, 0, ctx->st>>>(h->small.p, h->eb.p, n_entries, 1);
    k_fr_convert<<<(unsigned)((n_entries + 255) / 256), 256, 0, ctx->st>>>(h->small.p + n_entries, h->es.p, n_entries, 1);
    if ((e = ctx_sync(ctx))) return fail(e);
    *out = h;
    return BPPP_OK;
}
extern "C" void bppp_trrp_destroy(bppp_trrp* h) {
    if (!h) return;
    cudaSetDevice(h->gens->ctx->dev);
    if (h->tr) bppp_dtr_destroy(h->tr);
    g_alloc_stream = h->gens->ctx->st;
    delete h;
}
// phase 1 (TypedReciprocal.hs:399-410): scalars of the digit/multiplicity commitments, [batch][2][P0]
// canonical (dm row, m row), and the committed values [batch][n_ranges]; coms = [batch][2] points
static int trrp_phase1_impl(bppp_trrp* h, size_t batch, const uint8_t* sc_dm_m, const uint8_t* amounts, uint8_t* coms,
                           size_t n_inputs, const uint8_t* n_coms, uint8_t* chal_out) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->gens->ctx;
    if (!sc_dm_m || !amounts || !coms || batch == 0) FAIL(BPPP_ERR_ARG, "bppp_trrp_phase1: null/empty argument");
    ENTER(ctx);
    if (chal_out) {
        if (h->tr_fmt < 0) FAIL(BPPP_ERR_STATE, "bppp_trrp_phase1_tr: call bppp_trrp_set_transcript first");
        if (n_inputs && !n_coms) FAIL(BPPP_ERR_ARG, "bppp_trrp_phase1_tr: null input commitments");
        for (size_t i = 0; i < 2 * batch * n_inputs; i++)
            if (!host::fq_is_canonical(host::from_bytes(n_coms + 32 * i))) FAIL(BPPP_ERR_RANGE, "coordinate >= field modulus");
        const size_t cap = 4 + n_inputs + 2 * 64;              // bl, r, dm, m, inputs, (X, R) of up to 64 rounds
        if (h->tr && (h->tr->B != batch || h->tr->cap < cap || h->tr->fmt != h->tr_fmt)) { bppp_dtr_destroy(h->tr); h->tr = nullptr; }
        if (!h->tr) {
            int rc = bppp_dtr_create(ctx, batch, cap, h->tr_fmt, &h->tr);
            if (rc) return rc;
        }
        bppp_dtr_reset(h->tr);
    }
    const size_t P0 = h->gens->P0;
    if (!check_fr(sc_dm_m, batch * 2 * P0) || !check_fr(amounts, batch * h->n_ranges)) FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    h->B = batch;
    CK(h->scA.ensure(batch * 2 * P0)); CK(h->scR.ensure(batch * P0)); CK(h->scBL.ensure(batch * P0));
    CK(h->amounts.ensure(batch * h->n_ranges));
    CK(h->chal2.ensure(batch * 4)); CK(h->chal3.ensure(batch * 2)); CK(h->chal4.ensure(batch * 2));
    CK(h->xp.ensure(batch * h->n_ranges)); CK(h->vt.ensure(batch * h->n_bases));
    CK(h->r.ensure(batch * h->n_ent)); CK(h->c.ensure(batch * h->n_ent)); CK(h->bl.ensure(batch * h->n_ent));
    CK(h->w.ensure(batch * h->n_ent)); CK(h->small.ensure(batch * 8));
    CK(H2D(h->scA.p, sc_dm_m, batch * 2 * P0 * 32));
    CK(H2D(h->amounts.p, amounts, batch * h->n_ranges * 32));
    h->phase = 1;
    // digits and multiplicities: short scalars.  T3 e x r0 <- oracle' (dmCom : mCom : nComs)  (TypedReciprocal.hs:411)
    return trrp_commit(h, h->scA.p, 2 * batch, coms, 16.0, chal_out, 3, 2, n_coms, n_inputs);
}
extern "C" int bppp_trrp_phase1(bppp_trrp* h, size_t batch, const uint8_t* sc_dm_m, const uint8_t* amounts, uint8_t* coms) {
    return trrp_phase1_impl(h, batch, sc_dm_m, amounts, coms, 0, nullptr, nullptr);
}
// The same with the transcript on the device: n_coms = [batch][n_inputs] input commitments (host);
// chal = [batch][3] (e, x, r0), the first oracle call of proveTRRPM.
extern "C" int bppp_trrp_phase1_tr(bppp_trrp* h, size_t batch, const uint8_t* sc_dm_m, const uint8_t* amounts, size_t n_inputs,
                                   const uint8_t* n_coms, uint8_t* coms, uint8_t* chal) {
    if (!chal) return BPPP_ERR_ARG;
    return trrp_phase1_impl(h, batch, sc_dm_m, amounts, coms, n_inputs, n_coms, chal);
}
// show_format >= 0 enables the device transcript of this handle (TR_PREFIXED_P / TR_BARE_DECIMAL), -1 disables it
extern "C" int bppp_trrp_set_transcript(bppp_trrp* h, int show_format) {
    if (!h || show_format > 1) return BPPP_ERR_ARG;
    h->tr_fmt = show_format < 0 ? -1 : show_format;
    return BPPP_OK;
}
// phase 2 (:412-419): chal = [batch][4] (e, 1/e, x, 1/r0); r_sclin = [batch][1 + M] scalar and linear
// slots of the reciprocal witness with the err7 slot zero; err7_slot indexes the linear part.
// Out: rcom [batch] points, err7 [batch] scalars.
static int trrp_phase2_impl(bppp_trrp* h, const uint8_t* chal, const uint8_t* r_sclin, size_t err7_slot, uint8_t* rcom,
                           uint8_t* err7, uint8_t* chal_out) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->gens->ctx;
    if (!chal || !r_sclin || !rcom || !err7 || err7_slot >= h->gens->M) FAIL(BPPP_ERR_ARG, "bppp_trrp_phase2: bad argument");
    if (h->phase != 1) FAIL(BPPP_ERR_STATE, "bppp_trrp_phase2: call phase1 first");
    ENTER(ctx);
    const size_t B = h->B, P0 = h->gens->P0, N = h->gens->N, M = h->gens->M;
    if (!check_fr(chal, B * 4) || !check_fr(r_sclin, B * (1 + M))) FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    CK(H2D(h->chal2.p, chal, B * 4 * 32));
    // sc -> slot 0, lin -> slots 1+N.. of each scalar row
    CK(cudaMemcpy2DAsync(h->scR.p, P0 * 32, r_sclin, (1 + M) * 32, 32, B, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpy2DAsync(h->scR.p + 1 + N, P0 * 32, r_sclin + 32, (1 + M) * 32, M * 32, B, cudaMemcpyHostToDevice, ctx->st));
    ctx->h2d += B * (1 + M) * 32;
    { ProfScope ps_(ctx, K_TRRP, 0);
    k_trrp_tables<<<(unsigned)((B + 63) / 64), 64, 0, ctx->st>>>(h->chal2.p, 4, 2, (int)h->n_ranges, (int)h->n_bases, h->xp.p, h->vt.p, (int)B);
    }
    CK(cudaGetLastError());
    TrrpP2Args A;
    A.st = trrp_static(h); A.chal = h->chal2.p; A.scA = h->scA.p; A.P0 = P0; A.amounts = h->amounts.p;
    A.xp = h->xp.p; A.vt = h->vt.p; A.r = h->r.p; A.c = h->c.p; A.scR = h->scR.p;
    A.err7_slot = (int)(1 + N + err7_slot); A.err7 = h->small.p;
    { ProfScope ps_(ctx, K_TRRP, 0);
    k_trrp_phase2<<<(unsigned)B, TRRP_THREADS, 0, ctx->st>>>(A);
    }
    CK(cudaGetLastError());
    CK(D2H(err7, h->small.p, B * 32));
    h->phase = 2;
    return trrp_commit(h, h->scR.p, B, rcom, 256.0, chal_out, 3);     // T3 q x' r1 <- oracle' [rCom]  (:420)
}
extern "C" int bppp_trrp_phase2(bppp_trrp* h, const uint8_t* chal, const uint8_t* r_sclin, size_t err7_slot, uint8_t* rcom,
                                uint8_t* err7) {
    return trrp_phase2_impl(h, chal, r_sclin, err7_slot, rcom, err7, nullptr);
}
// with the device transcript: chal_out = [batch][3] (q, x', r1)
extern "C" int bppp_trrp_phase2_tr(bppp_trrp* h, const uint8_t* chal, const uint8_t* r_sclin, size_t err7_slot, uint8_t* rcom,
                                   uint8_t* err7, uint8_t* chal_out) {
    if (!chal_out) return BPPP_ERR_ARG;
    return trrp_phase2_impl(h, chal, r_sclin, err7_slot, rcom, err7, chal_out);
}
// phase 3, first half (:421-433): chal = [batch][2] (q-power base q0, x'); bls_nrm = [batch][N] blinders
// of the norm part.  Out: errs [batch][6] error terms over the norm entries.
static int trrp_phase3_impl(bppp_trrp* h, const uint8_t* chal, const uint8_t* bls_nrm, const char* const* seeds, const uint64_t* n0,
                           uint8_t* errs) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->gens->ctx;
    if (!chal || (!bls_nrm && (!seeds || !n0)) || !errs) FAIL(BPPP_ERR_ARG, "bppp_trrp_phase3: null argument");
    if (h->phase != 2) FAIL(BPPP_ERR_STATE, "bppp_trrp_phase3: call phase2 first");
    ENTER(ctx);
    const size_t B = h->B, P0 = h->gens->P0, N = h->gens->N;
    if (!check_fr(chal, B * 2) || (bls_nrm && !check_fr(bls_nrm, B * N))) FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    CK(H2D(h->chal3.p, chal, B * 2 * 32));
    if (bls_nrm) CK(H2D(h->bl.p, bls_nrm, B * N * 32));
    else {
        // blsNrm <- replicateM N random (TypedReciprocal.hs:425, src/ZKP.hs:90-93): hash(seed_b <> show (n0_b + i)) on the device
        std::vector<unsigned char> hs(B * 64, 0), hl(B);
        for (size_t b = 0; b < B; b++) {
            const size_t l = strlen(seeds[b]);
            if (l > 40) FAIL(BPPP_ERR_ARG, "bppp_trrp_phase3_rnd: seed longer than 40 bytes");
            memcpy(&hs[64 * b], seeds[b], l);
            hl[b] = (unsigned char)l;
        }
        CK(h->seeds.ensure(B * 64)); CK(h->seed_len.ensure(B)); CK(h->n0s.ensure(B));
        CK(H2D(h->seeds.p, hs.data(), B * 64));
        CK(H2D(h->seed_len.p, hl.data(), B));
        CK(H2D(h->n0s.p, n0, B * 8));
        { ProfScope ps_(ctx, K_TR_RANDOM, 0);
        k_tr_random<<<(unsigned)((B * N + TR_THREADS - 1) / TR_THREADS), TR_THREADS, 0, ctx->st>>>(
            h->seeds.p, h->seed_len.p, 0, h->n0s.p, B, N, h->bl.p, N);
        }
        CK(cudaGetLastError());
        CK(ctx_sync(ctx));              // the staging vectors above go out of scope
    }
    TrrpP3Args A;
    A.st = trrp_static(h); A.chal2 = h->chal2.p; A.chal3 = h->chal3.p; A.scA = h->scA.p; A.P0 = P0;
    A.r = h->r.p; A.c = h->c.p; A.bl = h->bl.p; A.xp = h->xp.p; A.vt = h->vt.p; A.errs = h->small.p;
    { ProfScope ps_(ctx, K_TRRP, 0);
    k_trrp_errterms<<<(unsigned)B, TRRP_THREADS, 0, ctx->st>>>(A);
    }
    CK(cudaGetLastError());
    CK(D2H(errs, h->small.p, B * 6 * 32));
    CK(ctx_sync(ctx));
    h->phase = 3;
    return BPPP_OK;
}
extern "C" int bppp_trrp_phase3(bppp_trrp* h, const uint8_t* chal, const uint8_t* bls_nrm, uint8_t* errs) {
    if (!bls_nrm) return BPPP_ERR_ARG;
    return trrp_phase3_impl(h, chal, bls_nrm, nullptr, nullptr, errs);
}
// the same with the N norm blinders drawn on the device: proof b takes `random` values number n0[b] .. n0[b] + N - 1
// of its randomSeed seeds[b] (at most 40 bytes); the caller advances its random counter by N
extern "C" int bppp_trrp_phase3_rnd(bppp_trrp* h, const uint8_t* chal, const char* const* seeds, const uint64_t* n0, uint8_t* errs) {
    if (!seeds || !n0) return BPPP_ERR_ARG;
    return trrp_phase3_impl(h, chal, nullptr, seeds, n0, errs);
}
// phase 3, second half (:434): bl_sclin = [batch][1 + M] scalar and linear slots of the blinding
// witness (its norm part is the bls_nrm of phase 3).  Out: blcom [batch] points.
static int trrp_commit_bl_impl(bppp_trrp* h, const uint8_t* bl_sclin, uint8_t* blcom, uint8_t* chal_out) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->gens->ctx;
    if (!bl_sclin || !blcom) FAIL(BPPP_ERR_ARG, "bppp_trrp_commit_bl: null argument");
    if (h->phase != 3) FAIL(BPPP_ERR_STATE, "bppp_trrp_commit_bl: call phase3 first");
    ENTER(ctx);
    const size_t B = h->B, P0 = h->gens->P0, N = h->gens->N, M = h->gens->M;
    if (!check_fr(bl_sclin, B * (1 + M))) FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    CK(cudaMemcpy2DAsync(h->scBL.p, P0 * 32, bl_sclin, (1 + M) * 32, 32, B, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpy2DAsync(h->scBL.p + 1 + N, P0 * 32, bl_sclin + 32, (1 + M) * 32, M * 32, B, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpy2DAsync(h->scBL.p + 1, P0 * 32, h->bl.p, N * 32, N * 32, B, cudaMemcpyDeviceToDevice, ctx->st));
    ctx->h2d += B * (1 + M) * 32;
    h->phase = 4;
    return trrp_commit(h, h->scBL.p, B, blcom, 256.0, chal_out, 1);   // t <- oracle [blCom]  (:435)
}
extern "C" int bppp_trrp_commit_bl(bppp_trrp* h, const uint8_t* bl_sclin, uint8_t* blcom) {
    return trrp_commit_bl_impl(h, bl_sclin, blcom, nullptr);
}
// with the device transcript: chal_out = [batch] (t)
extern "C" int bppp_trrp_commit_bl_tr(bppp_trrp* h, const uint8_t* bl_sclin, uint8_t* blcom, uint8_t* chal_out) {
    if (!chal_out) return BPPP_ERR_ARG;
    return trrp_commit_bl_impl(h, bl_sclin, blcom, chal_out);
}
// phase 4 (:435-444): chal = [batch][2] (t, 1/q0).  The norm part of the argument witness stays on
// the device; out: sums [batch][3] = (sum q2_i p_i^2, sum q2_i over digit entries, sum v_i over digit entries)
extern "C" int bppp_trrp_phase4(bppp_trrp* h, const uint8_t* chal, uint8_t* sums) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->gens->ctx;
    if (!chal || !sums) FAIL(BPPP_ERR_ARG, "bppp_trrp_phase4: null argument");
    if (h->phase != 4) FAIL(BPPP_ERR_STATE, "bppp_trrp_phase4: call commit_bl first");
    ENTER(ctx);
    const size_t B = h->B, P0 = h->gens->P0;
    if (!check_fr(chal, B * 2)) FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    CK(H2D(h->chal4.p, chal, B * 2 * 32));
    TrrpP4Args A;
    A.st = trrp_static(h); A.chal2 = h->chal2.p; A.chal3 = h->chal3.p; A.chal4 = h->chal4.p; A.scA = h->scA.p; A.P0 = P0;
    A.r = h->r.p; A.c = h->c.p; A.bl = h->bl.p; A.xp = h->xp.p; A.vt = h->vt.p; A.w = h->w.p; A.sums = h->small.p;
    { ProfScope ps_(ctx, K_TRRP, 0);
    k_trrp_phase4<<<(unsigned)B, TRRP_THREADS, 0, ctx->st>>>(A);
    }
    CK(cudaGetLastError());
    CK(D2H(sums, h->small.p, B * 3 * 32));
    CK(ctx_sync(ctx));
    h->phase = 5;
    return BPPP_OK;
}
// the norm-linear argument over the device-resident witness of phase 4 (q, s, l, c as in bppp_nl_create)
extern "C" int bppp_nl_create_trrp(bppp_trrp* h, const uint8_t* q, const uint8_t* s, const uint8_t* l, const uint8_t* c,
                                   bppp_nl** out) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->gens->ctx;
    if (!out || !q || !s || (h->gens->M && (!l || !c))) FAIL(BPPP_ERR_ARG, "bppp_nl_create_trrp: null argument");
    if (h->phase != 5) FAIL(BPPP_ERR_STATE, "bppp_nl_create_trrp: call phase4 first");
    *out = nullptr;
    ENTER(ctx);
    h->phase = 0;
    int rc = nl_create_impl(h->gens, false, BPPP_ARG_NL, h->B, q, s, nullptr, l, c, out, h->w.p);
    if (rc == BPPP_OK && h->tr_fmt >= 0 && h->tr && h->tr->n_calls) (*out)->dtr = h->tr;   // the rounds continue the proof's transcript
    return rc;
}

extern "C" void bppp_nl_destroy(bppp_nl* h) {
    if (!h) return;
    cudaSetDevice(h->ctx->dev);
    cudaStreamSynchronize(h->ctx->st);
    if (h->own_gens) bppp_gens_destroy(h->gens);
    if (h->tail_gens) bppp_gens_destroy(h->tail_gens);
    delete h;
}

// Sharding one large argument over several handles / GPUs: this handle holds the
// contiguous slice of the norm vector that starts at element `first_element` (a multiple of
// 2^rounds); only the weights q^(4i) of the scalar sums depend on the position.
extern "C" int bppp_nl_set_shard(bppp_nl* h, size_t first_element) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->ctx;
    if (h->kind != BPPP_ARG_NL || h->round != 0) FAIL(BPPP_ERR_STATE, "bppp_nl_set_shard: NL argument before the first round only");
    h->shard_lo = first_element;
    return BPPP_OK;
}
// Current state in stored form (fold mode): normalisations nn[b], nl[b], the generator vectors
// [G (n_norm) | H (n_lin)] and the public coefficients c[b][n_lin] as the handle holds them
// (G_true = G / nn, c_true = c / nl; the witness in true terms is bppp_nl_final's).
extern "C" int bppp_nl_export(bppp_nl* h, uint8_t* nn, uint8_t* nl, uint8_t* points, uint8_t* c) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->ctx;
    if (h->kind != BPPP_ARG_NL) FAIL(BPPP_ERR_STATE, "bppp_nl_export: NL argument only");
    ENTER(ctx);
    const size_t B = h->B, cn = h->curN, cl = h->curM;
    if (h->tensor && points && cn + cl) {
        // tensor mode keeps fold coefficients instead of folded generators: materialise
        // G^(r)_i = sum_{idx >> r == i} coef_idx * G_idx with fixed-base MSMs (meant for the few
        // elements left when a sharded argument is gathered)
        const size_t P0 = h->tP0, NO = cn + cl, NM = h->tN + h->tM;
        if (B * NO * P0 > ((size_t)1 << 26)) FAIL(BPPP_ERR_ARG, "bppp_nl_export: state too large to materialise in tensor mode");
        std::vector<Fr> hc(B * NM);
        if (h->tround > 0) {
            CK(D2H(hc.data(), h->coef.p, B * NM * 32));
            CK(ctx_sync(ctx));
        } else {
            for (auto& v : hc) v = h64::one();
        }
        std::vector<u256> sc(B * NO * P0, u256_zero());
        for (size_t b = 0; b < B; b++)
            for (size_t idx = 0; idx < NM; idx++) {
                const bool lin = idx >= h->tN;
                const size_t local = lin ? idx - h->tN : idx;
                const size_t j = (local >> h->tround) + (lin ? cn : 0);
                sc[(b * NO + j) * P0 + 1 + idx] = fr_canon_u256(hc[b * NM + idx]);
            }
        DBuf<u256> d_sc;
        DBuf<Jac> d_res;
        DBuf<Affine> d_aff;
        CK(d_sc.alloc(sc.size())); CK(d_res.alloc(B * NO)); CK(d_aff.alloc(B * NO));
        CK(H2D(d_sc.p, sc.data(), sc.size() * 32));
        int rc = run_msm_gens(h->tgens, P0, d_sc.p, P0, 0, B * NO, 1, d_res.p, 0);
        if (rc) return rc;
        if ((rc = to_affine(ctx, d_res.p, 1, d_aff.p, 1, 0, 1, B * NO))) return rc;
        CK(D2H(points, d_aff.p, B * NO * 64));
        CK(ctx_sync(ctx));
        points = nullptr;                                   // done
    }
    for (size_t b = 0; b < B; b++) {
        if (nn) h64::to_bytes(nn + 32 * b, h->nn[b]);
        if (nl) h64::to_bytes(nl + 32 * b, h->nl[b]);
    }
    if (points && cn + cl) {
        const Affine* src = h->curp < 0 ? h->gens->base.p + 1 : h->pts[h->curp].p + 1;
        size_t stride = h->curp < 0 ? 0 : h->P2;
        ctx->d2h += B * (cn + cl) * 64;
        for (size_t b = 0; b < B; b++)
            CK(cudaMemcpyAsync(points + 64 * b * (cn + cl), src + b * stride, (cn + cl) * 64, cudaMemcpyDeviceToHost, ctx->st));
    }
    if (c && cl) {
        DBuf<u256> tmp;
        CK(tmp.alloc(B * cl));
        for (size_t b = 0; b < B; b++) {
            { ProfScope ps_(ctx, K_FR_CONVERT, 0);
            k_fr_convert<<<(unsigned)((cl + 255) / 256), 256, 0, ctx->st>>>(h->c[h->cur].p + b * h->lstride[h->cur], tmp.p + b * cl, cl, 0);
            }
            CK(cudaGetLastError());
        }
        CK(D2H(c, tmp.p, B * cl * 32));
    }
    CK(ctx_sync(ctx));
    return BPPP_OK;
}

extern "C" int bppp_nl_lengths(bppp_nl* h, size_t* n_norm, size_t* n_lin) {
    if (!h) return BPPP_ERR_ARG;
    if (n_norm) *n_norm = h->kind == BPPP_ARG_IP ? 2 * h->curN : h->curN;   // IP.Norm.getWitness: two scalars per element
    if (n_lin) *n_lin = h->curM;
    return BPPP_OK;
}

static int nl_enqueue_commit(bppp_nl* h, bool affine = true);
static int nl_round_commit_impl(bppp_nl* h, uint8_t* X, uint8_t* R, uint8_t* E) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->ctx;
    if (!X || !R) FAIL(BPPP_ERR_ARG, "bppp_nl_round_commit: null output");
    ENTER(ctx);
    if (E && (!h->dtr || h->kind == BPPP_ARG_IP)) FAIL(BPPP_ERR_STATE, "bppp_nl_round_challenge: no device transcript attached to this argument");
    if (h->kind == BPPP_ARG_IP) return ip_round_commit(h, X, R);
    const size_t B = h->B;
    // per-proof constants of this round (NormArgument.hs:113): rho = q^4, k1 = 2 n^2 q^3, k2 = n^2 q^4
    std::vector<Fr> rho(B), k1(B), k2(B), coef(B * 8, h64::zero());
    host_parallel_for(B, [&](size_t b) {
        Fr q2 = h64::sqr(h->q[b]), q3 = h64::mul(q2, h->q[b]), q4 = h64::sqr(q2), n2 = h64::sqr(h->nn[b]);
        rho[b] = q4;
        k1[b] = h64::dbl(h64::mul(n2, q3));
        k2[b] = h64::mul(n2, q4);
        if (h->shard_lo) {              // weights of a shard start at (q^4)^(first pair index)
            Fr off = h64::pow_u64(q4, (uint64_t)(h->shard_lo >> (h->round + 1)));
            k1[b] = h64::mul(k1[b], off);
            k2[b] = h64::mul(k2[b], off);
        }
        coef[b * 8 + 1] = h->q[b];          // X on gL <- q * xR
        coef[b * 8 + 2] = h->qinv[b];       // X on gR <- q^-1 * xL
    });
    int rc;
    if (!h->have_partials) {
        if ((rc = upload_consts(h, C_RHO, rho))) return rc;
        if ((rc = launch_fold_dots(h, 0))) return rc;
        h->have_partials = true;
    }
    if ((rc = upload_consts(h, C_K1, k1))) return rc;
    if ((rc = upload_consts(h, C_K2, k2))) return rc;
    CK(H2D(cptr(h, C_COEF), coef.data(), B * 8 * 32));
    if ((rc = nl_enqueue_commit(h))) return rc;
    std::vector<Affine> xr(B * 2);
    std::vector<Fr> dots(B * 2);
    CK(D2H(xr.data(), h->aff.p, B * 2 * 64));
    CK(D2H(dots.data(), h->dots.p, B * 2 * 32));
    if (E) {                            // e <- head <$> oracle [X, R]  (src/Bulletproof.hs:351)
        if ((rc = dtr_absorb_dev(h->dtr, h->aff.p, 2, 2))) return rc;
        if ((rc = dtr_squeeze_first(h->dtr, 1))) return rc;
        CK(D2H(E, h->dtr->chal.p, B * 32));
    }
    CK(ctx_sync(ctx));
    for (size_t b = 0; b < B; b++) {
        memcpy(X + 64 * b, &xr[2 * b], 64);
        memcpy(R + 64 * b, &xr[2 * b + 1], 64);
        h->sX[b] = dots[2 * b];
        h->sR[b] = dots[2 * b + 1];
    }
    return BPPP_OK;
}
// The device work of a round's two commitments once the round constants (C_K1, C_K2, C_COEF; C_RHO and the dot
// partials) are in place: opening scalars in MSM order, the two MSMs, affine results in h->aff ([B][2] = X, R),
// the scalar parts sX, sR in h->dots.  No host synchronisation.
static int nl_enqueue_commit(bppp_nl* h, bool affine) {
    bppp_ctx* ctx = h->ctx;
    const size_t B = h->B;
    int rc;
    const size_t P0 = h->tensor ? h->tP0 : h->P0;               // row length of the MSM scalar vectors
    u256* xs = h->sc.p;
    u256* rs = h->sc.p + B * P0;
    {
        DotsFinishArgs A;
        memset(&A, 0, sizeof A);
        A.n_seg = 0;
        if (h->curN) { A.partial[A.n_seg] = h->part_n.p; A.n_blocks[A.n_seg] = h->blocks_n; A.k1[A.n_seg] = cptr(h, C_K1); A.k2[A.n_seg] = cptr(h, C_K2); A.n_seg++; }
        if (h->curM) { A.partial[A.n_seg] = h->part_l.p; A.n_blocks[A.n_seg] = h->blocks_l; A.k1[A.n_seg] = nullptr; A.k2[A.n_seg] = nullptr; A.n_seg++; }
        A.res = h->dots.p; A.xs = xs; A.rs = rs; A.sc_stride = P0; A.batch = (int)B;
        { ProfScope ps_(ctx, K_DOTS_FINISH, 0);
        k_dots_finish<<<(unsigned)((B + 127) / 128), 128, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    }
    const int src = h->cur;
    const bool expand = h->tensor && h->tround > 0;             // folded scalars go through the coefficient vector
    u256* fxs = expand ? h->fsc.p : xs;
    u256* frs = expand ? h->fsc.p + B * P0 : rs;
    if (h->curN) {
        MsmScalarsArgs A;
        A.x = h->w[src].p; A.in_stride = h->wstride[src]; A.n_in = (int)h->curN;
        A.xs = fxs; A.rs = frs; A.mont_out = expand; A.sc_stride = P0; A.off = 1; A.coef = cptr(h, C_COEF);
        const unsigned char kd[8] = {0, 2, 2, 0, 0, 0, 0, 1};
        memcpy(A.kind, kd, 8);
        { ProfScope ps_(ctx, K_MSM_SCALARS, 0);
        k_msm_scalars<<<dim3((unsigned)(((h->curN + 1) / 2 + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    }
    if (h->curM) {
        MsmScalarsArgs A;
        A.x = h->l[src].p; A.in_stride = h->lstride[src]; A.n_in = (int)h->curM;
        A.xs = fxs; A.rs = frs; A.mont_out = expand; A.sc_stride = P0; A.off = 1 + (int)h->curN; A.coef = cptr(h, C_COEF);
        const unsigned char kd[8] = {0, 1, 1, 0, 0, 0, 0, 1};
        memcpy(A.kind, kd, 8);
        { ProfScope ps_(ctx, K_MSM_SCALARS, 0);
        k_msm_scalars<<<dim3((unsigned)(((h->curM + 1) / 2 + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    }
    if (expand) {
        for (int seg = 0; seg < 2; seg++) {
            const size_t n0 = seg ? h->tM : h->tN;
            if (!n0) continue;
            ExpandArgs A;
            A.fx = fxs; A.fr_ = frs; A.f_stride = P0; A.f_off = seg ? 1 + (int)h->curN : 1;
            A.coef = h->coef.p; A.coef_stride = h->tN + h->tM; A.coef_off = seg ? (int)h->tN : 0;
            A.xs = xs; A.rs = rs; A.sc_stride = P0; A.off = seg ? 1 + (int)h->tN : 1;
            A.n = (int)n0; A.shift = h->tround;
            { ProfScope ps_(ctx, K_EXPAND, 0);
            k_expand_scalars<<<dim3((unsigned)((n0 + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(A);
            }
            CK(cudaGetLastError());
        }
    }
    // the two commitments: MSMs over [g | G | H] with X scalars (output 0) and R scalars (output 1)
    const size_t nterms = h->tensor ? P0 : 1 + h->curN + h->curM;
    // algorithmic work of this round in the reference's units: the X and R MSMs over the
    // CURRENT lengths, plus -- in tensor mode, where the launch below also does the job of the
    // previous round's generator fold (the folded generators are never materialised) -- that fold
    const double nX = 1.0 + (double)h->curN + (double)h->curM, nR = 1.0 + (double)((h->curN + 1) / 2) + (double)((h->curM + 1) / 2);
    double work = msm_alg_imads(nX) + msm_alg_imads(nR);
    if (h->tensor && h->round > 0) work += fold_alg_imads((double)(h->curN + h->curM));
    if (h->curp < 0 || h->tensor) {
        // the generators are the shared list (or the re-based one) -> fixed-base tables
        if ((rc = run_msm_gens(h->tensor ? h->tgens : h->gens, nterms, xs, P0, B * P0, B, 2, h->res.p, work))) return rc;
    } else {
        h->plan.slices.clear();
        h->plan.add(h->pts[h->curp].p, h->P2, xs, P0, B * P0, nterms);
        if ((rc = run_msm(ctx, h->plan, B, 2, h->res.p, work))) return rc;
    }
    return affine ? to_affine(ctx, h->res.p, 1, h->aff.p, 1, 0, 1, B * 2) : BPPP_OK;     // a shard's partial sums stay Jacobian
}
extern "C" int bppp_nl_round_commit(bppp_nl* h, uint8_t* X, uint8_t* R) { return nl_round_commit_impl(h, X, R, nullptr); }
// The commitments of the round and, from the device transcript the argument continues (bppp_nl_create_trrp after
// bppp_trrp_*_tr), its challenge E = [batch] scalars: proveRoundM's `oracle [X, R]` without a host hash
extern "C" int bppp_nl_round_challenge(bppp_nl* h, uint8_t* X, uint8_t* R, uint8_t* E) {
    if (!E) return BPPP_ERR_ARG;
    return nl_round_commit_impl(h, X, R, E);
}

namespace {
// BPPP_HYBRID_MAX=n: a tensor-mode argument switches to folding once at most n generators are left
// (default 0 = never).  From there on a round needs a handful of pair folds and two tiny MSMs
// instead of two full-length fixed-base MSMs -- about a quarter less arithmetic per 128by64 proof --
// but the small-size kernels (k_msm_gens_small, k_pair_fold, k_msm_bucket on <= 96 points) are
// latency-bound, and measured end to end the switch loses (6380 -> 4224..5592 proofs/s for n = 96..24),
// so it stays opt-in until those kernels are reworked.  Results are bit-identical either way (tested).
size_t hybrid_limit() {
    const char* ev = getenv("BPPP_HYBRID_MAX");
    return ev ? (size_t)atoll(ev) : (size_t)0;
}
// Materialise the folded generators G^(r)_j = sum_{idx >> r == j} coef_idx * G_idx of every proof (one
// small fixed-base MSM per block of 2^r original generators) and continue in fold mode.
int nl_switch_to_fold(bppp_nl* h) {
    bppp_ctx* ctx = h->ctx;
    const size_t B = h->B, N = h->tN, M = h->tM, P0 = h->tP0, cn = h->curN, cl = h->curM, group = (size_t)1 << h->tround;
    CK(h->pts[0].ensure(B * h->P2)); CK(h->pts[1].ensure(B * h->P2));
    CK(h->jscratch.ensure(B * (h->N2 + h->M2)));
    u256* sc = h->sc.p;                                   // free between rounds: canonical coefficients in scalar-row layout
    for (int seg = 0; seg < 2; seg++) {
        const size_t n0 = seg ? M : N;
        if (!n0) continue;
        { ProfScope ps_(ctx, K_FR_CONVERT, 0);
        k_fr_from_mont_rows<<<dim3((unsigned)((n0 + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(
            h->coef.p + (seg ? N : 0), N + M, sc, P0, 1 + (int)(seg ? N : 0), (int)n0);
        }
        CK(cudaGetLastError());
    }
    int rc;
    if ((rc = run_msm_groups(h->tgens, sc, P0, B, 1, N, group, h->jscratch.p, cn + cl))) return rc;
    if ((rc = run_msm_groups(h->tgens, sc, P0, B, 1 + N, M, group, h->jscratch.p + cn, cn + cl))) return rc;
    if ((rc = to_affine(ctx, h->jscratch.p, cn + cl, h->pts[0].p, h->P2, 1, (int)(cn + cl), B * (cn + cl)))) return rc;
    for (int k = 0; k < 2; k++) {
        { ProfScope ps_(ctx, K_BCAST, 0);
        k_bcast_point<<<(unsigned)((B + 127) / 128), 128, 0, ctx->st>>>(h->gens->base.p, h->pts[k].p, h->P2, B);
        }
        CK(cudaGetLastError());
    }
    h->tensor = false;
    h->curp = 0;
    return BPPP_OK;
}
}  // namespace

namespace {
// A single large argument is latency-bound once a few thousand generators are left: a fold-mode round
// then costs a 129-doubling chain per folded generator (k_pair_fold) plus a 128-doubling Horner per
// commitment, whatever the length.  At that point the CURRENT generators become the base list of a
// tensor-mode argument: one window table is built over them (252 doublings per point, all points in
// parallel, once), and every remaining round is two fixed-base MSMs with no doublings and no generator
// folds at all.  BPPP_REBASE_MAX=n moves the switch (default GT_TABLE_MAX_TERMS, 0 = never); results
// are bit-identical (tests).
size_t rebase_limit() {
    const char* ev = getenv("BPPP_REBASE_MAX");
    size_t v = ev ? (size_t)atoll(ev) : (size_t)GT_TABLE_MAX_TERMS;
    return std::min<size_t>(v, GT_TABLE_MAX_TERMS);
}
int nl_rebase_tensor(bppp_nl* h) {
    bppp_ctx* ctx = h->ctx;
    const size_t B = h->B, cn = h->curN, cl = h->curM;
    bppp_gens* tg = nullptr;
    // pts[curp] = [g | G^(k) (cn) | H^(k) (cl)] of the one proof: exactly a generator list
    int rc = gens_create_impl(ctx, cn, cl, nullptr, nullptr, nullptr, h->pts[h->curp].p, &tg);
    if (rc) return rc;
    h->tail_gens = tg; h->tgens = tg;
    h->tN = cn; h->tM = cl; h->tP0 = 1 + cn + cl; h->tround = 0;
    CK(h->coef.ensure(B * (cn + cl)));
    CK(h->fsc.ensure(2 * B * h->tP0));
    h->tensor = true;
    return BPPP_OK;
}
}  // namespace

static int nl_enqueue_fold(bppp_nl* h);
extern "C" int bppp_nl_round_fold(bppp_nl* h, const uint8_t* e) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->ctx;
    if (!e) FAIL(BPPP_ERR_ARG, "bppp_nl_round_fold: null challenge");
    if (!h->have_partials) FAIL(BPPP_ERR_STATE, "bppp_nl_round_fold before bppp_nl_round_commit");
    ENTER(ctx);
    const size_t B = h->B;
    if (!check_fr(e, B)) FAIL(BPPP_ERR_RANGE, "challenge >= group order");
    if (h->kind == BPPP_ARG_IP) return ip_round_fold(h, e);
    std::vector<Fr> au(B), bu(B), al(B), bl(B), ac(B), bc(B), rho(B), b0n(B), b0l(B), em(B), inv(2 * B), a0n(B);
    std::vector<u256> kk(B * 4);
    std::vector<unsigned char> sg(B * 2);
    host_parallel_for(B, [&](size_t b) {
        em[b] = h64::from_bytes(e + 32 * b);
        // NormArgument.hs:125  (a', b') = rationalReduceScalar (e * qInv)
        host::Ratio rn = host::rational_reduce(fr_canon_u256(h64::mul(em[b], h->qinv[b])));
        // NormArgument.hs:66   (a', b') = rationalReduceScalar e
        host::Ratio rl = host::rational_reduce(host::from_bytes(e + 32 * b));
        kk[b * 2 + 0] = rn.b; kk[b * 2 + 1] = rl.b;                  // kb[b][seg]
        kk[B * 2 + b * 2 + 0] = rn.a; kk[B * 2 + b * 2 + 1] = rl.a;  // ka[b][seg]
        sg[b * 2 + 0] = (unsigned char)((rn.b_neg ? 1 : 0) | (rn.a_neg ? 2 : 0));
        sg[b * 2 + 1] = (unsigned char)((rl.b_neg ? 1 : 0) | (rl.a_neg ? 2 : 0));
        b0n[b] = fr_from_mag(rn.b, rn.b_neg);
        a0n[b] = fr_from_mag(rn.a, rn.a_neg);
        b0l[b] = fr_from_mag(rl.b, rl.b_neg);
        ac[b] = b0l[b];                                              // c' = b0*cL + a0*cR
        bc[b] = fr_from_mag(rl.a, rl.a_neg);
        inv[b] = b0n[b];
        inv[B + b] = b0l[b];
    });
    h64::batch_inv(inv.data(), 2 * B);
    host_parallel_for(B, [&](size_t b) {
        // x' = b0Inv*xL + e*q*b0Inv*xR   (NormArgument.hs:129);  l' = b0Inv*xL + e*b0Inv*xR  (:71)
        au[b] = inv[b];
        bu[b] = h64::mul(h64::mul(em[b], h->q[b]), inv[b]);
        al[b] = inv[B + b];
        bl[b] = h64::mul(em[b], inv[B + b]);
        // s' = s + e*sX + (e^2 - 1)*sR   (Bulletproof.hs:352-353, makeEs NormArgument.hs:109)
        Fr e1 = h64::sub(h64::sqr(em[b]), h64::one());
        h->s[b] = h64::add(h->s[b], h64::add(h64::mul(em[b], h->sX[b]), h64::mul(e1, h->sR[b])));
        // n <- n*b0*qInv ; q <- q^2 ; linear n <- n*b0
        h->nn[b] = h64::mul(h64::mul(h->nn[b], b0n[b]), h->qinv[b]);
        h->nl[b] = h64::mul(h->nl[b], b0l[b]);
        h->q[b] = h64::sqr(h->q[b]);
        h->qinv[b] = h64::sqr(h->qinv[b]);
        rho[b] = h64::sqr(h64::sqr(h->q[b]));
    });
    int rc;
    if ((rc = upload_consts(h, C_AU, au)) || (rc = upload_consts(h, C_BU, bu)) || (rc = upload_consts(h, C_AL, al)) ||
        (rc = upload_consts(h, C_BL, bl)) || (rc = upload_consts(h, C_AC, ac)) || (rc = upload_consts(h, C_BC, bc)) ||
        (rc = upload_consts(h, C_RHO, rho)))
        return rc;
    CK(H2D(cptr(h, C_KB), kk.data(), B * 4 * 32));
    CK(H2D(h->sgn.p, sg.data(), B * 2));
    if (h->tensor && ((rc = upload_consts(h, C_A0N, a0n)) || (rc = upload_consts(h, C_B0N, b0n)) || (rc = upload_consts(h, C_A0L, bc)) ||
                      (rc = upload_consts(h, C_B0L, b0l))))
        return rc;
    if ((rc = nl_enqueue_fold(h))) return rc;
    CK(ctx_sync(ctx));   // host vectors above go out of scope
    return BPPP_OK;
}
// The device work of a fold once the fold factors (C_AU .. C_BC, C_RHO, C_KB / C_KA / sgn, C_A0N .. C_B0L) are in
// place: scalar vectors folded (with the next round's dot partials), fold coefficients updated (tensor mode) or
// generators folded (fold mode), lengths halved.  No host synchronisation except where a large single argument
// re-bases to tensor mode.
static int nl_enqueue_fold(bppp_nl* h) {
    bppp_ctx* ctx = h->ctx;
    const size_t B = h->B;
    int rc;
    // scalar vectors (and the next round's dots)
    if ((rc = launch_fold_dots(h, 1))) return rc;
    const size_t nN = (h->curN + 1) / 2, nM = (h->curM + 1) / 2;
    if (h->tensor) {
        // generators are not folded: update the per-generator fold coefficients instead
        for (int seg = 0; seg < 2; seg++) {
            const size_t n0 = seg ? h->tM : h->tN;
            if (!n0) continue;
            { ProfScope ps_(ctx, K_COEF, 0);
            k_coef_update<<<dim3((unsigned)((n0 + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(
                h->coef.p, h->tN + h->tM, seg ? (int)h->tN : 0, (int)n0, h->tround, cptr(h, seg ? C_A0L : C_A0N),
                cptr(h, seg ? C_B0L : C_B0N), 1, h->tround == 0);
            }
            CK(cudaGetLastError());
        }
        h->cur ^= 1;
        h->curN = nN;
        h->curM = nM;
        h->round++;
        h->tround++;
        if (hybrid_limit() && !h->tail_gens && !h->shard_lo && nN + nM <= hybrid_limit() && nN + nM >= 12 && h->round <= 9 && (rc = nl_switch_to_fold(h)))
            return rc;
        return BPPP_OK;
    }
    // generators
    PairFoldSeg segs[2] = {{1, (int)h->curN, 0}, {1 + (int)h->curN, (int)h->curM, (int)nN}};
    const Affine* in = h->curp < 0 ? h->gens->base.p : h->pts[h->curp].p;
    size_t in_stride = h->curp < 0 ? 0 : h->P2;
    int nxt = h->curp < 0 ? 0 : (h->curp ^ 1);
    // both segments always launched (an empty segment contributes zero blocks) so that the
    // [batch][2] layout of kb/ka/sgn holds
    if ((rc = launch_pair_fold(ctx, in, in_stride, h->jscratch.p, nN + nM, segs, 2, cptr(h, C_KB), cptr(h, C_KA),
                               h->sgn.p, B)))
        return rc;
    if ((rc = to_affine(ctx, h->jscratch.p, nN + nM, h->pts[nxt].p, h->P2, 1, (int)(nN + nM), B * (nN + nM))))
        return rc;
    h->curp = nxt;
    h->cur ^= 1;
    h->curN = nN;
    h->curM = nM;
    h->round++;
    if (B == 1 && !h->tail_gens && !h->no_rebase && 1 + nN + nM <= rebase_limit() && nN + nM >= 16) return nl_rebase_tensor(h);
    return BPPP_OK;
}

extern "C" int bppp_nl_final(bppp_nl* h, uint8_t* s, uint8_t* w, uint8_t* l) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->ctx;
    ENTER(ctx);
    if (h->kind == BPPP_ARG_IP) return ip_final(h, s, w, l);
    const size_t B = h->B, cn = h->curN, cl = h->curM;
    std::vector<Fr> hw(B * cn), hl(B * cl);
    if (w && cn) { ctx->d2h += B * cn * 32; CK(cudaMemcpy2DAsync(hw.data(), cn * 32, h->w[h->cur].p, h->wstride[h->cur] * 32, cn * 32, B, cudaMemcpyDeviceToHost, ctx->st)); }
    if (l && cl) { ctx->d2h += B * cl * 32; CK(cudaMemcpy2DAsync(hl.data(), cl * 32, h->l[h->cur].p, h->lstride[h->cur] * 32, cl * 32, B, cudaMemcpyDeviceToHost, ctx->st)); }
    CK(ctx_sync(ctx));
    for (size_t b = 0; b < B; b++) {
        if (s) h64::to_bytes(s + 32 * b, h->s[b]);
        if (w) for (size_t i = 0; i < cn; i++) h64::to_bytes(w + 32 * (b * cn + i), h64::mul(h->nn[b], hw[b * cn + i]));
        if (l) for (size_t i = 0; i < cl; i++) h64::to_bytes(l + 32 * (b * cl + i), h64::mul(h->nl[b], hl[b * cl + i]));
    }
    return BPPP_OK;
}

// Bare arguments (bppp_nl_create / bppp_nl_create_gens): continue the device transcript `t` (same context, same batch;
// not owned) -- e.g. after bppp_dtr_absorb of the initial commitment.  NULL detaches.
extern "C" int bppp_nl_attach_transcript(bppp_nl* h, bppp_dtr* t) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->ctx;
    if (t && (t->ctx != ctx || t->B != h->B)) FAIL(BPPP_ERR_ARG, "bppp_nl_attach_transcript: transcript of another context / batch size");
    h->dtr = t;
    return BPPP_OK;
}

// ---- NCCL communicator for one argument sharded over several GPUs.  libnccl is loaded at run time
// (dlopen): the library has no link-time dependency on it and single-GPU users never touch it.
namespace {
typedef struct { char internal[128]; } nccl_unique_id;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(void**, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
bool nccl_load(const char* path) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return true;
    void* lib = nullptr;
    if (path && *path) lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);      // the copy the process already has (torch's)
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { g_nccl.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
    NcclApi a;
    a.lib = lib;
    a.GetUniqueId = (int (*)(nccl_unique_id*))dlsym(lib, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(void**, int, nccl_unique_id, int))dlsym(lib, "ncclCommInitRank");
    a.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    a.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    a.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString) {
        g_nccl.err = "libnccl.so.2 lacks an expected symbol";
        return false;
    }
    g_nccl = a;
    return true;
}
}  // namespace
struct bppp_comm {
    bppp_ctx* ctx;
    void* comm = nullptr;               // ncclComm_t; null for a one-rank communicator
    int world = 1, rank = 0;
};
#define NCCL_UINT8 1                     // ncclUint8 (nccl.h)
// path = the libnccl.so.2 to use (NULL: the one already in the process, else the system's)
extern "C" int bppp_comm_load(const char* path) { return nccl_load(path) ? BPPP_OK : BPPP_ERR_STATE; }
extern "C" const char* bppp_comm_last_error(void) { return g_nccl.err.c_str(); }
// rank 0 draws the 128-byte id and hands it to the other ranks out of band (e.g. torch.distributed broadcast)
extern "C" int bppp_comm_unique_id(uint8_t id[128]) {
    if (!id || !nccl_load(nullptr)) return BPPP_ERR_STATE;
    nccl_unique_id u;
    int rc = g_nccl.GetUniqueId(&u);
    if (rc) { g_nccl.err = g_nccl.GetErrorString(rc); return BPPP_ERR_CUDA; }
    memcpy(id, u.internal, 128);
    return BPPP_OK;
}
// one communicator per context (its collectives run on the context's stream); world == 1 needs no NCCL at all
extern "C" int bppp_comm_create(bppp_ctx* ctx, int world, int rank, const uint8_t id[128], bppp_comm** out) {
    if (!ctx) return BPPP_ERR_ARG;
    if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) FAIL(BPPP_ERR_ARG, "bppp_comm_create: bad argument");
    *out = nullptr;
    ENTER(ctx);
    bppp_comm* c = new bppp_comm();
    c->ctx = ctx; c->world = world; c->rank = rank;
    if (world > 1) {
        if (!nccl_load(nullptr)) { delete c; FAIL(BPPP_ERR_STATE, g_nccl.err); }
        nccl_unique_id u;
        memcpy(u.internal, id, 128);
        int rc = g_nccl.CommInitRank(&c->comm, world, u, rank);
        if (rc) { delete c; FAIL(BPPP_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc)); }
    }
    *out = c;
    return BPPP_OK;
}
extern "C" void bppp_comm_destroy(bppp_comm* c) {
    if (!c) return;
    if (c->comm) {
        cudaSetDevice(c->ctx->dev);
        cudaStreamSynchronize(c->ctx->st);
        g_nccl.CommDestroy(c->comm);
    }
    delete c;
}
namespace {
// all ranks contribute `bytes` bytes; recv holds world * bytes in rank order (a device copy when there is one rank)
int comm_all_gather(bppp_comm* c, const void* send, void* recv, size_t bytes) {
    bppp_ctx* ctx = c->ctx;
    if (c->world == 1) {
        CK(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, ctx->st));
        return BPPP_OK;
    }
    int rc = g_nccl.AllGather(send, recv, bytes, NCCL_UINT8, c->comm, ctx->st);
    if (rc) FAIL(BPPP_ERR_CUDA, std::string("ncclAllGather: ") + g_nccl.GetErrorString(rc));
    return BPPP_OK;
}

// ---- the device-resident round loop (rounds.cuh)
struct DevRun {                          // outputs of all rounds, newest first
    DBuf<Affine> resp;                   // [B][total][2]
    DBuf<u256> esd;                      // [B][total]
    size_t total = 0;
    DBuf<unsigned char> send, recv;      // sharded runs: one record per rank
};
RoundState nl_round_state(bppp_nl* h) {
    RoundState S;
    S.q = cptr(h, C_Q); S.qinv = cptr(h, C_QINV); S.nn = cptr(h, C_NN); S.nl = cptr(h, C_NL); S.s = cptr(h, C_S);
    S.rho = cptr(h, C_RHO); S.k1 = cptr(h, C_K1); S.k2 = cptr(h, C_K2); S.coef = cptr(h, C_COEF);
    S.au = cptr(h, C_AU); S.bu = cptr(h, C_BU); S.al = cptr(h, C_AL); S.bl = cptr(h, C_BL); S.ac = cptr(h, C_AC); S.bc = cptr(h, C_BC);
    S.a0n = cptr(h, C_A0N); S.b0n = cptr(h, C_B0N); S.a0l = cptr(h, C_A0L); S.b0l = cptr(h, C_B0L);
    S.kb = cptr(h, C_KB); S.ka = cptr(h, C_KA); S.sgn = h->sgn.p; S.inv = cptr(h, C_INV);
    S.chal = h->dtr->chal.p; S.dots = h->dots.p; S.B = (int)h->B; S.tensor = 0; S.pair_off = 0;
    return S;
}
// the per-proof round state (host vectors of a fresh handle) moves to the device
int nl_dev_upload_state(bppp_nl* h) {
    bppp_ctx* ctx = h->ctx;
    const size_t B = h->B;
    CK(H2D(cptr(h, C_Q), h->q.data(), B * 32)); CK(H2D(cptr(h, C_QINV), h->qinv.data(), B * 32));
    CK(H2D(cptr(h, C_NN), h->nn.data(), B * 32)); CK(H2D(cptr(h, C_NL), h->nl.data(), B * 32));
    CK(H2D(cptr(h, C_S), h->s.data(), B * 32));
    CK(cudaMemsetAsync(cptr(h, C_COEF), 0, B * 8 * 32, ctx->st));
    return BPPP_OK;
}
// rounds [done, done + n) of R.total; with `comm` the handle is one rank's shard: partial commitments and scalar
// parts are all-gathered (256 bytes per rank) and summed in rank order before the transcript sees them
int nl_dev_rounds(bppp_nl* h, size_t n, size_t done, DevRun& R, bppp_comm* comm) {
    bppp_ctx* ctx = h->ctx;
    const size_t B = h->B, total = R.total;
    RoundState S = nl_round_state(h);
    int rc;
    for (size_t r = done; r < done + n; r++) {
        if (h->curN + h->curM == 0 && !comm) FAIL(BPPP_ERR_STATE, "device round loop: nothing left to fold");
        S.pair_off = h->shard_lo ? (unsigned long long)(h->shard_lo >> (h->round + 1)) : 0ull;
        { ProfScope ps_(ctx, K_ROUND_STATE, 0);
        k_round_pre<<<(unsigned)((B + 127) / 128), 128, 0, ctx->st>>>(S);
        }
        CK(cudaGetLastError());
        if (!h->have_partials) {
            if ((rc = launch_fold_dots(h, 0))) return rc;
            h->have_partials = true;
        }
        if ((rc = nl_enqueue_commit(h, comm == nullptr))) return rc;
        if (comm) {
            { ProfScope ps_(ctx, K_ROUND_STATE, 0);
            k_shard_pack<<<1, 32, 0, ctx->st>>>(h->res.p, h->dots.p, R.send.p);
            }
            CK(cudaGetLastError());
            if ((rc = comm_all_gather(comm, R.send.p, R.recv.p, SHARD_REC_BYTES))) return rc;
            { ProfScope ps_(ctx, K_ROUND_STATE, 0);
            k_shard_combine<<<1, 32, 0, ctx->st>>>(R.recv.p, comm->world, h->res.p, h->dots.p);
            }
            CK(cudaGetLastError());
            if ((rc = to_affine(ctx, h->res.p, 1, h->aff.p, 1, 0, 1, B * 2))) return rc;
        }
        // responses are consed: newest first (Bulletproof.hs:357-359)
        CK(cudaMemcpy2DAsync(R.resp.p + 2 * (total - 1 - r), total * 128, h->aff.p, 128, 128, B, cudaMemcpyDeviceToDevice, ctx->st));
        // e <- head <$> oracle [X, R]  (Bulletproof.hs:351)
        if ((rc = dtr_absorb_dev(h->dtr, h->aff.p, 2, 2))) return rc;
        if ((rc = dtr_squeeze_first(h->dtr, 1))) return rc;
        CK(cudaMemcpy2DAsync(R.esd.p + (total - 1 - r), total * 32, h->dtr->chal.p, 32, 32, B, cudaMemcpyDeviceToDevice, ctx->st));
        S.tensor = h->tensor ? 1 : 0;
        if (!S.tensor) {
            ProfScope ps_(ctx, K_ROUND_STATE, 0);
            k_round_ratio<<<(unsigned)((2 * B + 63) / 64), 64, 0, ctx->st>>>(S);
        }
        CK(cudaGetLastError());
        { ProfScope ps_(ctx, K_ROUND_STATE, 0);
        k_round_post<<<(unsigned)((B + 127) / 128), 128, 0, ctx->st>>>(S);
        }
        CK(cudaGetLastError());
        if ((rc = nl_enqueue_fold(h))) return rc;
    }
    return BPPP_OK;
}
// getWitness: the final vectors with their normalisations and the opening scalar; then everything goes to the host
int nl_dev_finish(bppp_nl* h, DevRun& R, uint8_t* responses, uint8_t* es, uint8_t* s, uint8_t* w, uint8_t* l) {
    bppp_ctx* ctx = h->ctx;
    const size_t B = h->B, cn = h->curN, cl = h->curM;
    DBuf<u256> fin;
    CK(fin.alloc(B * (1 + cn + cl)));
    u256* fs = fin.p; u256* fw = fin.p + B; u256* fl = fw + B * cn;
    struct { const u256* v; size_t stride; const u256* sc; size_t n; u256* out; } jobs[3] = {
        {cptr(h, C_S), 1, nullptr, 1, fs}, {h->w[h->cur].p, h->wstride[h->cur], cptr(h, C_NN), cn, fw},
        {h->l[h->cur].p, h->lstride[h->cur], cptr(h, C_NL), cl, fl}};
    for (auto& j : jobs) {
        if (!j.n) continue;
        { ProfScope ps_(ctx, K_ROUND_STATE, 0);
        k_scale_rows<<<(unsigned)((j.n * B + 127) / 128), 128, 0, ctx->st>>>(j.v, j.stride, j.sc, (int)j.n, (int)B, j.out);
        }
        CK(cudaGetLastError());
    }
    CK(D2H(responses, R.resp.p, B * R.total * 128));
    if (es) CK(D2H(es, R.esd.p, B * R.total * 32));
    if (s) CK(D2H(s, fs, B * 32));
    if (w && cn) CK(D2H(w, fw, B * cn * 32));
    if (l && cl) CK(D2H(l, fl, B * cl * 32));
    CK(ctx_sync(ctx));
    return BPPP_OK;
}
}  // namespace

// proveBPM (src/Bulletproof.hs:357-359) for the whole batch as ONE stream of launches: per round the two
// commitments, `oracle [X, R]` on the device transcript, rationalReduceScalar and the fold factors in
// k_round_ratio / k_round_post (rounds.cuh), the folds -- no host synchronisation until the results are read.
// The handle must be fresh (no round taken yet) and have a transcript (bppp_nl_create_trrp after the _tr phases, or
// bppp_nl_attach_transcript).  responses = [batch][rounds][2] points and es = [batch][rounds] challenges (may be
// NULL), NEWEST FIRST; s / w / l as bppp_nl_final.  Bit-identical to the step-by-step calls (tests).
extern "C" int bppp_nl_prove_device(bppp_nl* h, size_t rounds, uint8_t* responses, uint8_t* es, uint8_t* s, uint8_t* w, uint8_t* l) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->ctx;
    if (!responses || rounds == 0 || rounds > 64) FAIL(BPPP_ERR_ARG, "bppp_nl_prove_device: bad argument");
    if (h->kind != BPPP_ARG_NL) FAIL(BPPP_ERR_STATE, "bppp_nl_prove_device: norm-linear arguments only");
    if (!h->dtr) FAIL(BPPP_ERR_STATE, "bppp_nl_prove_device: no device transcript attached");
    if (h->round != 0 || h->have_partials || h->shard_lo) FAIL(BPPP_ERR_STATE, "bppp_nl_prove_device: needs a fresh, unsharded argument");
    ENTER(ctx);
    DevRun R;
    R.total = rounds;
    CK(R.resp.alloc(h->B * rounds * 2)); CK(R.esd.alloc(h->B * rounds));
    int rc;
    if ((rc = nl_dev_upload_state(h))) return rc;
    if ((rc = nl_dev_rounds(h, rounds, 0, R, nullptr))) return rc;
    return nl_dev_finish(h, R, responses, es, s, w, l);
}

// One large argument sharded over the GPUs of a box.  Every rank holds a contiguous slice of the norm
// vector and of its generators (bppp_nl_create over the slice, then bppp_nl_set_shard with the slice's first index;
// equal power-of-two slice lengths); the linear part lives on rank 0 (`lin_len` = its length, the other ranks create
// their handles with M = 0); every rank passes the SAME q and opening scalar s, and a transcript in the same state.
//   rounds 1 .. local_rounds: adjacent-pair folds keep a contiguous slice local (src/Bulletproof.hs:77-90); the
//     commitments are sums of per-rank partial MSMs -- 256 bytes per rank all-gathered with NCCL on the context's
//     stream and added in rank order on the device, so every rank's transcript absorbs the same X, R;
//   then the folded slices (vectors, generators, rank 0's linear part) are all-gathered once and every rank finishes
//     the remaining rounds on the whole (short) argument, re-based to tensor mode -- identical results on all ranks.
// Outputs as bppp_nl_prove_device, on every rank.  The proof is bit-identical to the unsharded one (tests, tools).
extern "C" int bppp_nl_prove_sharded(bppp_nl* h, bppp_comm* comm, size_t rounds, size_t local_rounds, size_t lin_len,
                                     uint8_t* responses, uint8_t* es, uint8_t* s, uint8_t* w, uint8_t* l) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->ctx;
    if (!comm || comm->ctx != ctx || !responses || rounds == 0 || rounds > 64 || local_rounds > rounds)
        FAIL(BPPP_ERR_ARG, "bppp_nl_prove_sharded: bad argument");
    if (h->kind != BPPP_ARG_NL || h->B != 1) FAIL(BPPP_ERR_STATE, "bppp_nl_prove_sharded: one norm-linear argument");
    if (!h->dtr) FAIL(BPPP_ERR_STATE, "bppp_nl_prove_sharded: no device transcript attached");
    if (h->round != 0 || h->have_partials) FAIL(BPPP_ERR_STATE, "bppp_nl_prove_sharded: needs a fresh argument");
    const size_t W = (size_t)comm->world, len0 = h->N;
    if (len0 == 0 || (len0 & (len0 - 1)) || (len0 >> local_rounds) == 0) FAIL(BPPP_ERR_ARG, "bppp_nl_prove_sharded: slice length must be a power of two >= 2^local_rounds");
    if (h->shard_lo != (size_t)comm->rank * len0) FAIL(BPPP_ERR_STATE, "bppp_nl_prove_sharded: bppp_nl_set_shard(rank * slice length) first");
    if ((comm->rank == 0) != (h->M == lin_len) && lin_len) FAIL(BPPP_ERR_ARG, "bppp_nl_prove_sharded: the linear part lives on rank 0 only");
    ENTER(ctx);
    int rc;
    if (h->tensor) {                     // slices fold their generators for real: the gathered tail needs them
        CK(h->pts[0].ensure(h->B * h->P2)); CK(h->pts[1].ensure(h->B * h->P2));
        CK(h->jscratch.ensure(h->B * (h->N2 + h->M2)));
        for (int k = 0; k < 2; k++) {
            k_bcast_point<<<1, 128, 0, ctx->st>>>(h->gens->base.p, h->pts[k].p, h->P2, h->B);
            CK(cudaGetLastError());
        }
        h->tensor = false;
        h->curp = -1;
    }
    DevRun R;
    R.total = rounds;
    CK(R.resp.alloc(rounds * 2)); CK(R.esd.alloc(rounds));
    CK(R.send.alloc(SHARD_REC_BYTES)); CK(R.recv.alloc(W * SHARD_REC_BYTES));
    if ((rc = nl_dev_upload_state(h))) return rc;
    // no re-base of a slice to tensor mode while it is a shard
    h->no_rebase = true;
    if ((rc = nl_dev_rounds(h, local_rounds, 0, R, comm))) return rc;
    h->no_rebase = false;
    // ---- gather: [w slice | G slice] of every rank, rank 0's [l | c | H]
    const size_t L = h->curN;
    size_t cl = lin_len;
    for (size_t i = 0; i < local_rounds; i++) cl = (cl + 1) / 2;
    if (L != (len0 >> local_rounds)) FAIL(BPPP_ERR_STATE, "bppp_nl_prove_sharded: unexpected slice length");
    const size_t rec = L * 32 + L * 64 + cl * (32 + 32 + 64);
    DBuf<unsigned char> snd, rcv;
    CK(snd.alloc(rec)); CK(rcv.alloc(W * rec));
    CK(cudaMemsetAsync(snd.p, 0, rec, ctx->st));
    const Affine* cur_pts = h->curp < 0 ? h->gens->base.p : h->pts[h->curp].p;        // [g | G' | H']
    CK(cudaMemcpyAsync(snd.p, h->w[h->cur].p, L * 32, cudaMemcpyDeviceToDevice, ctx->st));
    CK(cudaMemcpyAsync(snd.p + L * 32, cur_pts + 1, L * 64, cudaMemcpyDeviceToDevice, ctx->st));
    if (cl && comm->rank == 0) {
        CK(cudaMemcpyAsync(snd.p + L * 96, h->l[h->cur].p, cl * 32, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(snd.p + L * 96 + cl * 32, h->c[h->cur].p, cl * 32, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(snd.p + L * 96 + cl * 64, cur_pts + 1 + L, cl * 64, cudaMemcpyDeviceToDevice, ctx->st));
    }
    if ((rc = comm_all_gather(comm, snd.p, rcv.p, rec))) return rc;
    const size_t NT = W * L;
    DBuf<Affine> tp;                     // [g | G (NT) | H (cl)]
    DBuf<u256> tw, tl, tc;
    CK(tp.alloc(1 + NT + cl)); CK(tw.alloc(std::max<size_t>(NT, 1))); CK(tl.alloc(std::max<size_t>(cl, 1))); CK(tc.alloc(std::max<size_t>(cl, 1)));
    CK(cudaMemcpyAsync(tp.p, h->gens->base.p, 64, cudaMemcpyDeviceToDevice, ctx->st));
    for (size_t r = 0; r < W; r++) {
        CK(cudaMemcpyAsync(tw.p + r * L, rcv.p + r * rec, L * 32, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(tp.p + 1 + r * L, rcv.p + r * rec + L * 32, L * 64, cudaMemcpyDeviceToDevice, ctx->st));
    }
    if (cl) {
        CK(cudaMemcpyAsync(tl.p, rcv.p + L * 96, cl * 32, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(tc.p, rcv.p + L * 96 + cl * 32, cl * 32, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(tp.p + 1 + NT, rcv.p + L * 96 + cl * 64, cl * 64, cudaMemcpyDeviceToDevice, ctx->st));
    }
    // the round state is the same on every rank (same challenges): read it back once
    std::vector<Fr> st(5);
    const int slots[5] = {C_Q, C_QINV, C_NN, C_NL, C_S};
    for (int i = 0; i < 5; i++) CK(D2H(&st[i], cptr(h, slots[i]), 32));
    CK(ctx_sync(ctx));
    bppp_gens* tg = nullptr;
    if ((rc = gens_create_impl(ctx, NT, cl, nullptr, nullptr, nullptr, tp.p, &tg))) return rc;
    bppp_nl* t = nullptr;
    uint8_t qb[32], sb[32];
    h64::to_bytes(qb, st[0]); h64::to_bytes(sb, st[4]);
    if ((rc = nl_create_impl(tg, true, BPPP_ARG_NL, 1, qb, sb, nullptr, nullptr, nullptr, &t, tw.p, tl.p, tc.p))) { bppp_gens_destroy(tg); return rc; }
    t->q[0] = st[0]; t->qinv[0] = st[1]; t->nn[0] = st[2]; t->nl[0] = st[3]; t->s[0] = st[4];
    t->dtr = h->dtr;
    if ((rc = nl_dev_upload_state(t)) == BPPP_OK && (rc = nl_dev_rounds(t, rounds - local_rounds, local_rounds, R, nullptr)) == BPPP_OK)
        rc = nl_dev_finish(t, R, responses, es, s, w, l);
    if (rc) ctx->err = std::string("sharded tail: ") + ctx->err;
    std::string keep = ctx->err;
    bppp_nl_destroy(t);
    ctx->err = keep;
    return rc;
}

// =============================================================================== verifier
namespace {
int nl_verify_impl(bppp_gens* gens, int kind, size_t batch, size_t k, const uint8_t* q, const uint8_t* s_pub,
                   const uint8_t* pub_w, const uint8_t* c, const uint8_t* es, const uint8_t* XR, size_t n_norm,
                   size_t n_lin, const uint8_t* fw, const uint8_t* fl, size_t n_init, const uint8_t* init_s,
                   const uint8_t* init_p, int* ok, const u256* pub_dev = nullptr, const uint8_t* weights = nullptr) {
    bppp_ctx* ctx = gens->ctx;
    const size_t N = gens->N, M = gens->M;
    const size_t B = batch, P0 = 1 + N + M, NX = n_init + 2 * k;
    if (weights && !check_fr(weights, B)) FAIL(BPPP_ERR_RANGE, "batch weight >= group order");
    if (k > 30) FAIL(BPPP_ERR_ARG, "too many rounds");
    {   // the final witness has exactly the lengths k rounds leave (roundReduce, src/Bulletproof.hs:300-304; the
        // reference's decodeProof' derives them from the setup, src/RangeProof.hs:70-71): a surplus final scalar
        // would enter the scalar check sum (wgt * v^2) without being bound to any generator
        size_t en = kind == BPPP_ARG_IP ? (N + 1) / 2 : N, el = M;
        for (size_t r = 0; r < k; r++) { en = en / 2 + en % 2; el = el / 2 + el % 2; }
        if (kind == BPPP_ARG_IP) en *= 2;
        if (n_norm != en || n_lin != el) FAIL(BPPP_ERR_ARG, "final witness lengths do not match the number of rounds");
    }
    if (!check_fq(XR, 4 * k * B) || !check_fq(init_p, 2 * n_init * B)) FAIL(BPPP_ERR_RANGE, "coordinate >= field modulus");
    if (!check_fr(q, B) || !check_fr(s_pub, B) || (!pub_dev && !check_fr(pub_w, B * N)) || !check_fr(c, B * M) || !check_fr(es, B * k) ||
        !check_fr(fw, B * n_norm) || !check_fr(fl, B * n_lin) || !check_fr(init_s, B * n_init))
        FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    DBuf<Affine> extra, aff;
    DBuf<u256> sc, xsc, pub, vs_n, vs_l, f0n, f1, f0l, tmp;
    DBuf<Jac> res, res2, resx;
    CK(extra.alloc(B * std::max<size_t>(NX, 1))); CK(aff.alloc(B));
    CK(sc.alloc(B * P0)); CK(xsc.alloc(B * std::max<size_t>(NX, 1)));
    if (!pub_dev) CK(pub.alloc(B * N));
    CK(vs_n.alloc(B * n_norm)); CK(vs_l.alloc(B * n_lin));
    CK(f0n.alloc(B * k)); CK(f1.alloc(B * k)); CK(f0l.alloc(B * k)); CK(res.alloc(B)); CK(res2.alloc(B)); CK(resx.alloc(B));
    CK(tmp.alloc(std::max(B * N, B * M)));
    if (N && !pub_dev) {
        CK(H2D(tmp.p, pub_w, B * N * 32));
        { ProfScope ps_(ctx, K_FR_CONVERT, 0);
        k_fr_convert<<<(unsigned)((B * N + 255) / 256), 256, 0, ctx->st>>>(tmp.p, pub.p, B * N, 1);
        }
        CK(cudaGetLastError());
    }
    const u256* pubp = pub_dev ? pub_dev : pub.p;            // Montgomery, [B][N]
    // host: challenges, tensor factors, final-witness scalar sc
    //   NL: NormArgument.hs:131-145, 73-81        IP: InnerProductArgument.hs:103-124, 172-181
    const bool ip = kind == BPPP_ARG_IP;
    const size_t nvs = ip ? n_norm / 2 : n_norm;
    if (ip && (n_norm & 1)) FAIL(BPPP_ERR_ARG, "IP final norm witness has an even number of scalars");
    std::vector<Fr> hf0n(B * k), hf1(B * k), hf0l(B * k, h64::one()), hvn(B * std::max<size_t>(nvs, 1)), hvl(B * n_lin), scn(B);
    std::vector<Fr> hf1y(B * k), hvy(B * std::max<size_t>(nvs, 1)), hr(B), einv(B * std::max<size_t>(k, 1));
    std::vector<u256> hx(B * std::max<size_t>(NX, 1));
    std::vector<Affine> hp(B * std::max<size_t>(NX, 1));
    if (ip) {
        for (size_t i = 0; i < B * k; i++) einv[i] = h64::from_bytes(es + 32 * i);
        h64::batch_inv(einv.data(), B * k);
    }
    const Fr half = h64::inv(h64::from_u64(2)), four = h64::from_u64(4);
    host_parallel_for(B, [&](size_t b) {
        Fr qq = h64::from_bytes(q + 32 * b);
        if (ip) { hr[b] = qq; qq = h64::sqr(h64::sqr(qq)); }          // q = r^4  (makeNorm :196)
        for (size_t j = 0; j < k; j++) {
            // round j+1's challenge is es[k-1-j] (newest first)
            const size_t ei = b * k + (k - 1 - j);
            hf1[b * k + j] = ip ? einv[ei] : h64::from_bytes(es + 32 * ei);       // IP: esX = recip <$> esY
            if (ip) hf1y[b * k + j] = h64::from_bytes(es + 32 * ei);
            hf0n[b * k + j] = qq;
            qq = h64::sqr(qq);
        }
        Fr acc = h64::zero();
        if (ip) {
            Fr wgt = qq;                                                          // powers' qF
            for (size_t i = 0; i < nvs; i++) {
                Fr s0 = h64::from_bytes(fw + 32 * (b * n_norm + 2 * i)), s1 = h64::from_bytes(fw + 32 * (b * n_norm + 2 * i + 1));
                Fr vx = h64::mul(half, h64::add(s0, s1)), vy = h64::mul(half, h64::sub(s1, s0));   // makeNorm 1 (RangeProof.hs:80)
                hvn[b * nvs + i] = vx; hvy[b * nvs + i] = vy;
                acc = h64::add(acc, h64::mul(wgt, h64::mul(vx, vy)));
                wgt = h64::mul(wgt, qq);
            }
            acc = h64::mul(four, acc);                                            // sIP = 4
        } else {
            Fr qF2 = h64::sqr(qq), wgt = qF2;                                     // powers' (qF^2)
            for (size_t i = 0; i < n_norm; i++) {
                Fr v = h64::from_bytes(fw + 32 * (b * n_norm + i));
                hvn[b * n_norm + i] = v;
                acc = h64::add(acc, h64::mul(wgt, h64::sqr(v)));
                wgt = h64::mul(wgt, qF2);
            }
        }
        scn[b] = acc;
        for (size_t i = 0; i < n_lin; i++) hvl[b * n_lin + i] = h64::from_bytes(fl + 32 * (b * n_lin + i));
        // extra terms: initCom opening, then (e0, X), (e1, R) per round, newest first (verifyWith)
        for (size_t i = 0; i < n_init; i++) {
            hx[b * NX + i] = host::from_bytes(init_s + 32 * (b * n_init + i));
            memcpy(&hp[b * NX + i], init_p + 64 * (b * n_init + i), 64);
        }
        for (size_t r = 0; r < k; r++) {
            Fr e = h64::from_bytes(es + 32 * (b * k + r));
            if (ip) {                                                             // makeEs = (recip e, e)
                hx[b * NX + n_init + 2 * r] = fr_canon_u256(einv[b * k + r]);
                hx[b * NX + n_init + 2 * r + 1] = host::from_bytes(es + 32 * (b * k + r));
            } else {                                                              // makeEs = (e, e^2 - 1)
                hx[b * NX + n_init + 2 * r] = host::from_bytes(es + 32 * (b * k + r));
                hx[b * NX + n_init + 2 * r + 1] = fr_canon_u256(h64::sub(h64::sqr(e), h64::one()));
            }
            memcpy(&hp[b * NX + n_init + 2 * r], XR + 64 * ((b * k + r) * 2), 128);
        }
    });
    DBuf<u256> vs_y, f1y, rdev;
    if (ip) {
        CK(vs_y.alloc(B * std::max<size_t>(nvs, 1))); CK(f1y.alloc(B * std::max<size_t>(k, 1))); CK(rdev.alloc(B));
        if (nvs) CK(H2D(vs_y.p, hvy.data(), B * nvs * 32));
        if (k) CK(H2D(f1y.p, hf1y.data(), B * k * 32));
        CK(H2D(rdev.p, hr.data(), B * 32));
    }
    if (k) {
        CK(H2D(f0n.p, hf0n.data(), B * k * 32));
        CK(H2D(f1.p, hf1.data(), B * k * 32));
        CK(H2D(f0l.p, hf0l.data(), B * k * 32));
    }
    if (nvs) CK(H2D(vs_n.p, hvn.data(), B * nvs * 32));
    if (n_lin) CK(H2D(vs_l.p, hvl.data(), B * n_lin * 32));
    DBuf<int> offcurve;                                      // per proof: a commitment / response point is off the curve
    if (NX) {
        CK(H2D(xsc.p, hx.data(), B * NX * 32));
        CK(H2D(extra.p, hp.data(), B * NX * 64));
        int rc0 = check_points_async(ctx, extra.p, B * NX, NX, offcurve);
        if (rc0) return rc0;
    }
    if (N && ip) {
        IpVerifyArgs A;
        A.pub = pubp; A.pub_stride = N; A.vx = vs_n.p; A.vy = vs_y.p; A.n_vs = (int)nvs;
        A.f0x = f0n.p; A.f1x = f1.p; A.f1y = f1y.p; A.k = (int)k; A.r = rdev.p;
        A.out = sc.p; A.out_stride = P0; A.off = 1; A.n = (int)N;
        { ProfScope ps_(ctx, K_TENSOR, 0);
        k_ip_verify_scalars<<<dim3((unsigned)(((N + 1) / 2 + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    } else if (N) {
        TensorArgs A;
        A.pub = pubp; A.pub_stride = N; A.vs = vs_n.p; A.n_vs = (int)n_norm; A.f0 = f0n.p; A.f1 = f1.p; A.k = (int)k;
        A.out = sc.p; A.out_stride = P0; A.off = 1; A.n = (int)N;
        { ProfScope ps_(ctx, K_TENSOR, 0);
        k_tensor_expand<<<dim3((unsigned)((N + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    }
    if (M) {
        TensorArgs A;
        A.pub = nullptr; A.pub_stride = 0; A.vs = vs_l.p; A.n_vs = (int)n_lin; A.f0 = f0l.p; A.f1 = f1.p; A.k = (int)k;
        A.out = sc.p; A.out_stride = P0; A.off = 1 + (int)N; A.n = (int)M;
        { ProfScope ps_(ctx, K_TENSOR, 0);
        k_tensor_expand<<<dim3((unsigned)((M + 255) / 256), (unsigned)B), 256, 0, ctx->st>>>(A);
        }
        CK(cudaGetLastError());
    }
    // the scalar on g: s_pub - (norm part) + sum_j c_j * (-tensor_j) -- sc_lin = contract' . tensor' (NormArgument.hs:75-78)
    // is a dot product with what the kernel above just wrote, so it stays on the device (no synchronisation here)
    {
        std::vector<u256> d0(B);
        for (size_t b = 0; b < B; b++) d0[b] = fr_canon_u256(h64::sub(h64::from_bytes(s_pub + 32 * b), scn[b]));
        DBuf<u256> cdev, ddev;
        CK(cdev.alloc(B * std::max<size_t>(M, 1))); CK(ddev.alloc(B));
        if (M) CK(H2D(cdev.p, c, B * M * 32));
        CK(H2D(ddev.p, d0.data(), B * 32));
        { ProfScope ps_(ctx, K_TENSOR, 0);
        k_verify_s0<<<(unsigned)B, VS0_THREADS, 0, ctx->st>>>(cdev.p, ddev.p, sc.p, P0, 1 + (int)N, (int)M);
        }
        CK(cudaGetLastError());
    }
    int rc;
    if (weights && B > 1) {
        // Batch verification across proofs: sum_b rho_b (sum_i sc[b][i] G_i + sum_j xsc[b][j] P[b][j]) = 0 -- ONE fixed-base
        // MSM over the shared generators with the weighted column sums, ONE Pippenger over all B * NX per-proof points.
        // A pass accepts every proof of the batch (a false proof survives with probability ~2^-128 over the weights);
        // a failure says nothing about which proof is bad, so it falls through to the per-proof checks below.
        DBuf<u256> rho, wsc, wx;
        DBuf<Jac> r1, r2;
        DBuf<Affine> a1;
        CK(rho.alloc(B)); CK(wsc.alloc(P0)); CK(wx.alloc(B * std::max<size_t>(NX, 1))); CK(r1.alloc(3)); CK(a1.alloc(1));
        CK(H2D(rho.p, weights, B * 32));
        { ProfScope ps_(ctx, K_FR_CONVERT, 0);
        k_fr_convert<<<(unsigned)((B + 255) / 256), 256, 0, ctx->st>>>(rho.p, rho.p, B, 1);
        }
        CK(cudaGetLastError());
        { ProfScope ps_(ctx, K_BATCH_WEIGHT, 0);
        k_weight_columns<<<(unsigned)P0, WCOL_THREADS, 0, ctx->st>>>(sc.p, P0, rho.p, (int)B, wsc.p);
        }
        CK(cudaGetLastError());
        if ((rc = run_msm_gens(gens, P0, wsc.p, P0, 0, 1, 1, r1.p, msm_alg_imads((double)P0)))) return rc;
        if (NX) {
            { ProfScope ps_(ctx, K_BATCH_WEIGHT, 0);
            k_weight_rows<<<(unsigned)((B * NX + 255) / 256), 256, 0, ctx->st>>>(xsc.p, rho.p, (int)NX, B * NX, wx.p);
            }
            CK(cudaGetLastError());
            MsmPlan plan;
            plan.add(extra.p, 0, wx.p, 0, 0, B * NX);
            if ((rc = run_msm(ctx, plan, 1, 1, r1.p + 1, msm_alg_imads((double)(B * NX))))) return rc;
            { ProfScope ps_(ctx, K_JAC_SUM, 0);
            k_jac_sum<<<1, 128, 0, ctx->st>>>(r1.p, 1, r1.p + 1, 1, r1.p + 2, 1);
            }
            CK(cudaGetLastError());
        }
        if ((rc = to_affine(ctx, NX ? r1.p + 2 : r1.p, 1, a1.p, 1, 0, 1, 1))) return rc;
        Affine total;
        std::vector<int> bad(B, 0);
        CK(D2H(&total, a1.p, 64));
        if (NX) CK(D2H(bad.data(), offcurve.p, B * sizeof(int)));
        CK(ctx_sync(ctx));
        bool all_ok = aff_is_inf(total);
        for (size_t b = 0; b < B; b++) all_ok = all_ok && !bad[b];
        if (all_ok) {
            for (size_t b = 0; b < B; b++) ok[b] = 1;
            return BPPP_OK;
        }
    }
    rc = run_msm_gens(gens, P0, sc.p, P0, 0, B, 1, res.p, msm_alg_imads((double)(P0 + NX)));
    if (rc) return rc;
    if (NX) {
        MsmPlan plan;
        plan.add(extra.p, NX, xsc.p, NX, 0, NX);
        if ((rc = run_msm(ctx, plan, B, 1, resx.p, 0))) return rc;
        { ProfScope ps_(ctx, K_JAC_SUM, 0);
        k_jac_sum<<<(unsigned)((B + 127) / 128), 128, 0, ctx->st>>>(res.p, 1, resx.p, 1, res2.p, B);
        }
        CK(cudaGetLastError());
    }
    if ((rc = to_affine(ctx, NX ? res2.p : res.p, 1, aff.p, 1, 0, 1, B))) return rc;
    std::vector<Affine> out(B);
    std::vector<int> bad(B, 0);
    CK(D2H(out.data(), aff.p, B * 64));
    if (NX) CK(D2H(bad.data(), offcurve.p, B * sizeof(int)));
    CK(ctx_sync(ctx));
    for (size_t b = 0; b < B; b++) ok[b] = (aff_is_inf(out[b]) && !bad[b]) ? 1 : 0;
    return BPPP_OK;
}
}  // namespace

// verifier: chal = [batch][8] = (e, 1/e, x, x', q0, 1/q0, t, 0).  The norm part of the public constants
// (makePublicConsts, TypedReciprocal.hs:236-263) stays on the device for bppp_nl_verify_trrp;
// sums = [batch][3] as in bppp_trrp_phase4.
extern "C" int bppp_trrp_verify_pub(bppp_trrp* h, size_t batch, const uint8_t* chal, uint8_t* sums) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->gens->ctx;
    if (!chal || !sums || batch == 0) FAIL(BPPP_ERR_ARG, "bppp_trrp_verify_pub: null/empty argument");
    ENTER(ctx);
    if (!check_fr(chal, batch * 8)) FAIL(BPPP_ERR_RANGE, "scalar >= group order");
    h->B = batch;
    CK(h->chalv.ensure(batch * 8)); CK(h->xp.ensure(batch * h->n_ranges)); CK(h->vt.ensure(batch * h->n_bases));
    CK(h->w.ensure(batch * h->n_ent)); CK(h->small.ensure(batch * 8));
    CK(H2D(h->chalv.p, chal, batch * 8 * 32));
    { ProfScope ps_(ctx, K_TRRP, 0);
    k_trrp_tables<<<(unsigned)((batch + 63) / 64), 64, 0, ctx->st>>>(h->chalv.p, 8, 2, (int)h->n_ranges, (int)h->n_bases, h->xp.p, h->vt.p, (int)batch);
    }
    CK(cudaGetLastError());
    TrrpVArgs A;
    A.st = trrp_static(h); A.chal = h->chalv.p; A.xp = h->xp.p; A.vt = h->vt.p; A.pub = h->w.p; A.sums = h->small.p;
    { ProfScope ps_(ctx, K_TRRP, 0);
    k_trrp_verify_pub<<<(unsigned)batch, TRRP_THREADS, 0, ctx->st>>>(A);
    }
    CK(cudaGetLastError());
    CK(D2H(sums, h->small.p, batch * 3 * 32));
    CK(ctx_sync(ctx));
    h->phase = 10;
    return BPPP_OK;
}
// bppp_nl_verify_gens with the public vector of bppp_trrp_verify_pub (norm-linear argument)
extern "C" int bppp_nl_verify_trrp(bppp_trrp* h, size_t k, const uint8_t* q, const uint8_t* s_pub, const uint8_t* c,
                                   const uint8_t* es, const uint8_t* XR, size_t n_norm, size_t n_lin, const uint8_t* fw,
                                   const uint8_t* fl, size_t n_init, const uint8_t* init_s, const uint8_t* init_p, int* ok) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->gens->ctx;
    if (h->phase != 10) FAIL(BPPP_ERR_STATE, "bppp_nl_verify_trrp: call bppp_trrp_verify_pub first");
    if (!q || !s_pub || !ok || (h->gens->M && !c) || (k && (!es || !XR)) || (n_norm && !fw) || (n_lin && !fl) || (n_init && (!init_s || !init_p)))
        FAIL(BPPP_ERR_ARG, "bppp_nl_verify_trrp: null argument");
    ENTER(ctx);
    h->phase = 0;
    return nl_verify_impl(h->gens, BPPP_ARG_NL, h->B, k, q, s_pub, nullptr, c, es, XR, n_norm, n_lin, fw, fl, n_init, init_s, init_p, ok, h->w.p);
}

// Batch verification across proofs: the same inputs plus weights = [batch] scalars the caller drew at
// random AFTER seeing the proofs (128 bits are enough).  One combined check; all-accept sets every ok[b] = 1, a
// failure is resolved by the per-proof checks, so the verdicts are always exact per proof.
extern "C" int bppp_nl_verify_trrp_rlc(bppp_trrp* h, size_t k, const uint8_t* q, const uint8_t* s_pub, const uint8_t* c,
                                       const uint8_t* es, const uint8_t* XR, size_t n_norm, size_t n_lin, const uint8_t* fw,
                                       const uint8_t* fl, size_t n_init, const uint8_t* init_s, const uint8_t* init_p,
                                       const uint8_t* weights, int* ok) {
    if (!h) return BPPP_ERR_ARG;
    bppp_ctx* ctx = h->gens->ctx;
    if (h->phase != 10) FAIL(BPPP_ERR_STATE, "bppp_nl_verify_trrp: call bppp_trrp_verify_pub first");
    if (!q || !s_pub || !ok || !weights || (h->gens->M && !c) || (k && (!es || !XR)) || (n_norm && !fw) || (n_lin && !fl) ||
        (n_init && (!init_s || !init_p)))
        FAIL(BPPP_ERR_ARG, "bppp_nl_verify_trrp_rlc: null argument");
    ENTER(ctx);
    h->phase = 0;
    return nl_verify_impl(h->gens, BPPP_ARG_NL, h->B, k, q, s_pub, nullptr, c, es, XR, n_norm, n_lin, fw, fl, n_init, init_s, init_p, ok,
                          h->w.p, weights);
}
extern "C" int bppp_nl_verify_gens_rlc(bppp_gens* gens, int kind, size_t batch, size_t k, const uint8_t* q, const uint8_t* s_pub,
                                       const uint8_t* pub_w, const uint8_t* c, const uint8_t* es, const uint8_t* XR,
                                       size_t n_norm, size_t n_lin, const uint8_t* fw, const uint8_t* fl, size_t n_init,
                                       const uint8_t* init_s, const uint8_t* init_p, const uint8_t* weights, int* ok) {
    if (!gens) return BPPP_ERR_ARG;
    bppp_ctx* ctx = gens->ctx;
    if (kind != BPPP_ARG_NL && kind != BPPP_ARG_IP) FAIL(BPPP_ERR_ARG, "bppp_nl_verify: unknown argument kind");
    if (!q || !s_pub || !ok || !weights || batch == 0 || (gens->N && !pub_w) || (gens->M && !c) || (k && (!es || !XR)) ||
        (n_norm && !fw) || (n_lin && !fl) || (n_init && (!init_s || !init_p)))
        FAIL(BPPP_ERR_ARG, "bppp_nl_verify_gens_rlc: null/empty argument");
    if (k > 30) FAIL(BPPP_ERR_ARG, "too many rounds");
    ENTER(ctx);
    return nl_verify_impl(gens, kind, batch, k, q, s_pub, pub_w, c, es, XR, n_norm, n_lin, fw, fl, n_init, init_s, init_p, ok, nullptr, weights);
}

extern "C" int bppp_nl_verify_gens(bppp_gens* gens, int kind, size_t batch, size_t k, const uint8_t* q, const uint8_t* s_pub,
                                   const uint8_t* pub_w, const uint8_t* c, const uint8_t* es, const uint8_t* XR,
                                   size_t n_norm, size_t n_lin, const uint8_t* fw, const uint8_t* fl, size_t n_init,
                                   const uint8_t* init_s, const uint8_t* init_p, int* ok) {
    if (!gens) return BPPP_ERR_ARG;
    bppp_ctx* ctx = gens->ctx;
    if (kind != BPPP_ARG_NL && kind != BPPP_ARG_IP) FAIL(BPPP_ERR_ARG, "bppp_nl_verify: unknown argument kind");
    if (!q || !s_pub || !ok || batch == 0 || (gens->N && !pub_w) || (gens->M && !c) || (k && (!es || !XR)) ||
        (n_norm && !fw) || (n_lin && !fl) || (n_init && (!init_s || !init_p)))
        FAIL(BPPP_ERR_ARG, "bppp_nl_verify: null/empty argument");
    if (k > 30) FAIL(BPPP_ERR_ARG, "too many rounds");
    ENTER(ctx);
    return nl_verify_impl(gens, kind, batch, k, q, s_pub, pub_w, c, es, XR, n_norm, n_lin, fw, fl, n_init, init_s, init_p, ok);
}
extern "C" int bppp_nl_verify(bppp_ctx* ctx, int kind, size_t batch, size_t N, size_t M, size_t k, const uint8_t* g,
                              const uint8_t* G, const uint8_t* H, const uint8_t* q, const uint8_t* s_pub,
                              const uint8_t* pub_w, const uint8_t* c, const uint8_t* es, const uint8_t* XR,
                              size_t n_norm, size_t n_lin, const uint8_t* fw, const uint8_t* fl, size_t n_init,
                              const uint8_t* init_s, const uint8_t* init_p, int* ok) {
    if (!ctx) return BPPP_ERR_ARG;
    if (!g || (N && !G) || (M && !H)) FAIL(BPPP_ERR_ARG, "bppp_nl_verify: null generators");
    bppp_gens* gens = nullptr;
    int rc = bppp_gens_create(ctx, N, M, g, G, H, &gens);
    if (rc) return rc;
    rc = bppp_nl_verify_gens(gens, kind, batch, k, q, s_pub, pub_w, c, es, XR, n_norm, n_lin, fw, fl, n_init, init_s, init_p, ok);
    bppp_gens_destroy(gens);
    return rc;
}

// =============================================================================== debug
extern "C" int bppp_dbg_field(bppp_ctx* ctx, int op, size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    if (!ctx || !a || !b || !out) return BPPP_ERR_ARG;
    ENTER(ctx);
    DBuf<u256> da, db, dc;
    CK(da.alloc(n)); CK(db.alloc(n)); CK(dc.alloc(n));
    CK(H2D(da.p, a, n * 32));
    CK(H2D(db.p, b, n * 32));
    { ProfScope ps_(ctx, K_DBG, WORK_K_DBG_FIELD);
    k_dbg_field<<<(unsigned)((n + 127) / 128), 128, 0, ctx->st>>>(da.p, db.p, dc.p, n, op);
    }
    CK(cudaGetLastError());
    CK(D2H(out, dc.p, n * 32));
    CK(ctx_sync(ctx));
    return BPPP_OK;
}
extern "C" int bppp_dbg_ec(bppp_ctx* ctx, int op, size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    if (!ctx || !a || !b || !out) return BPPP_ERR_ARG;
    ENTER(ctx);
    DBuf<Affine> da, db, dd;
    DBuf<Jac> dc;
    CK(da.alloc(n)); CK(db.alloc(n)); CK(dc.alloc(n)); CK(dd.alloc(n));
    CK(H2D(da.p, a, n * 64));
    CK(H2D(db.p, b, n * 64));
    { ProfScope ps_(ctx, K_DBG, WORK_K_DBG_EC);
    k_dbg_ec<<<(unsigned)((n + 127) / 128), 128, 0, ctx->st>>>(da.p, db.p, dc.p, n, op);
    }
    CK(cudaGetLastError());
    int rc = to_affine(ctx, dc.p, n, dd.p, n, 0, (int)n, n);
    if (rc) return rc;
    CK(D2H(out, dd.p, n * 64));
    CK(ctx_sync(ctx));
    return BPPP_OK;
}
