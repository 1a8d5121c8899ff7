// Host side of the range-proof provers / verifiers, above the device C ABI.
//
// The reference keeps three things on the CPU and so does this file: the Fiat-Shamir transcript,
// the round sequencing, and the range proofs' scalar phases.  Every elliptic-curve operation goes
// to the GPU through bppp_msm_batch / bppp_nl_* (include/bppp_b200.h); there is no CPU group law
// here.  Mirrors, with the same names and argument meaning:
//   src/RangeProof/Internal.hs         RPWitness, commitRPW, blindWitness, blindErrWitness,
//                                      blindBlindingTerm, makePolyTerms
//   src/RangeProof/TypedReciprocal.hs  makeRangeData, digits, makePhase1s, makePhase2s,
//                                      makeSharedCoeffs, makeErrorTerms, makePublicConsts,
//                                      inputCoeffs, setup, witnessTRRP, proveTRRPM, verifyTRRPM
//   src/RangeProof/Binary.hs           makeRangeData, makeDigits, makePublicConsts, inputCoeffs,
//                                      setupBRP, witnessBRP, proveBRPM, verifyBRPM
//   src/RangeProof.hs                  RangeProof.proveM / verifyM, infoRP
//   src/Bulletproof.hs                 proveBPM / verifyBPM round loops, optimalWitnessSize
// A batch of proofs runs in lock-step: host phases are spread over worker threads, each device
// call covers the whole batch.
#include <malloc.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <algorithm>
#include <atomic>
#include <dlfcn.h>
#include <signal.h>
#include <sys/time.h>
#include <sys/random.h>
#include <ucontext.h>
#include <map>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/bppp_b200.h"
#include "transcript.hpp"
#include "../host_math.hpp"

using namespace bppp;
using h64::Fr;
typedef __int128 I128;
typedef unsigned __int128 U128;

namespace {

// ------------------------------------------------------------------------------ utilities
static bool t_lane_threads_is_main();
extern thread_local int t_lane_id;
struct TraceEv { int lane; const char* name; double t0, t1; };
struct Timing {                                // BPPP_TIMING=1: coarse wall-clock split printed to stderr
    std::map<std::string, double> ms;
    bool on = getenv("BPPP_TIMING") != nullptr;
    bool trace = getenv("BPPP_TRACE") != nullptr;   // BPPP_TRACE=1: every lane's phases as (lane, name, start, end)
    std::mutex tmu;
    std::vector<TraceEv> evs;
    static double now() {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    void start() {
        t_progress() = 0;
        if (trace) t_last() = now();
        if (on && t_lane_threads_is_main()) last = now();
    }
    void lap(const char* name) {
        t_progress()++;
        if (trace) {
            double t = now();
            { std::lock_guard<std::mutex> lk(tmu); evs.push_back({t_lane_id, name, t_last(), t}); }
            t_last() = t;
        }
        if (on && t_lane_threads_is_main()) { double t = now(); ms[name] += t - last; last = t; }
    }
    void dump(const char* what) {
        if (!on || !t_lane_threads_is_main()) return;
        fprintf(stderr, "[bppp timing] %s:", what);
        for (auto& kv : ms) fprintf(stderr, " %s=%.1fms", kv.first.c_str(), kv.second);
        fprintf(stderr, "\n");
        ms.clear();
    }
    void dump_trace(const char* what) {
        if (!trace) return;
        std::lock_guard<std::mutex> lk(tmu);
        for (auto& e : evs) fprintf(stderr, "[bppp trace] %s lane=%d %s %.3f %.3f\n", what, e.lane, e.name, e.t0, e.t1);
        evs.clear();
    }
    double last = 0;
    static double& t_last() { static thread_local double v = 0; return v; }
    // phase boundaries passed by this lane in the current call (its place in the pipeline)
    static int& t_progress() { static thread_local int v = 0; return v; }
};
Timing g_tm;
// fine-grained per-proof section timers (summed over threads), BPPP_TIMING=1
enum { S_WITNESS, S_RANDOM, S_SCALARS, S_ORACLE, S_PHASE2, S_COEFFS, S_ERRTERMS, S_BLIND, S_PUB, S_COMBINE, S_TOBYTES, S_V_ORACLE,
       S_V_PHASE2, S_V_PUB, S_V_MISC, S_COUNT };
const char* const kSectNames[S_COUNT] = {"witness", "random", "commit_scalars", "oracle", "phase2", "coeffs", "errterms", "blind",
                                          "pub", "combine", "tobytes", "v_oracle", "v_phase2", "v_pub", "v_misc"};
std::atomic<uint64_t> g_sect[S_COUNT];
struct Sect {
    double t;
    bool on;
    Sect() : t(0), on(g_tm.on) { if (on) t = Timing::now(); }
    void lap(int id) { if (on) { double n = Timing::now(); g_sect[id] += (uint64_t)((n - t) * 1e6); t = n; } }
};
extern std::atomic<uint64_t> g_turn_wait_us, g_turn_hold_us, g_turn_calls;
void dump_sections(const char* what, size_t proofs) {
    if (!g_tm.on) return;
    fprintf(stderr, "[bppp sections] %s (us per proof):", what);
    for (int i = 0; i < S_COUNT; i++) {
        uint64_t v = g_sect[i].exchange(0);
        if (v) fprintf(stderr, " %s=%.0f", kSectNames[i], v / 1e3 / (double)proofs);
    }
    fprintf(stderr, "\n");
    fprintf(stderr, "[bppp jobs] %s: host jobs=%llu, lane-time inside them (sum over lanes)=%.1fms\n", what,
            (unsigned long long)g_turn_calls.exchange(0), g_turn_hold_us.exchange(0) / 1e3);
}
extern thread_local bool t_is_lane0;
static bool t_lane_threads_is_main() { return t_is_lane0; }
int g_threads = 0;
extern thread_local int t_lane_threads;
int n_threads() {
    if (t_lane_threads > 0) return t_lane_threads;
    if (g_threads > 0) return g_threads;
    unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 4;
}
// Host phases of the concurrent lanes run on ONE pool of worker threads (as many as the host threads
// this process may use).  A phase is a job of per-proof items; workers always take the next items
// of the pending job whose lane is FURTHEST along (ties: lowest lane, then first come).  So
//  * lanes that start together leave their host phases one after the other instead of all at
//    once, and the first lane's device rounds overlap the later lanes' host phases;
//  * a lane in its argument rounds -- a few hundred microseconds of transcript hashing that gate
//    its next device launch -- is served at the next item boundary, never behind a long phase;
//  * the pool is work-conserving: when the preferred job has no items left to hand out, idle
//    workers start on the next one (no gap while a phase drains).
// The lanes' own threads only sequence device calls and sleep while their job runs.
std::atomic<uint64_t> g_turn_wait_us{0}, g_turn_hold_us{0}, g_turn_calls{0};   // BPPP_TIMING=1
class Scheduler {
  public:
    explicit Scheduler(int workers) {
        for (int i = 0; i < workers; i++) th_.emplace_back([this] { loop(); });
    }
    ~Scheduler() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        work_.notify_all();
        for (auto& t : th_) t.join();
    }
    int workers() const { return (int)th_.size(); }
    // fn(i) for i in [0, n); returns when all are done
    void run(size_t n, int64_t prio, const std::function<void(size_t)>& fn) {
        if (n == 0) return;
        Job j;
        j.prio = prio; j.n = n; j.fn = &fn;
        j.grain = std::max<size_t>(1, std::min<size_t>(4, n / ((size_t)th_.size() * 16 + 1)));
        std::unique_lock<std::mutex> lk(mu_);
        j.seq = seq_++;
        auto pos = jobs_.begin();
        while (pos != jobs_.end() && ((*pos)->prio > j.prio || ((*pos)->prio == j.prio && (*pos)->seq < j.seq))) ++pos;
        jobs_.insert(pos, &j);
        work_.notify_all();
        j.cv.wait(lk, [&] { return j.done == j.n; });
    }

  private:
    struct Job {
        int64_t prio = 0;
        uint64_t seq = 0;
        size_t n = 0, next = 0, done = 0, grain = 1;
        const std::function<void(size_t)>* fn = nullptr;
        std::condition_variable cv;
    };
    void loop() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            work_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
            if (stop_) return;
            Job* j = jobs_.front();                         // highest priority job that still has items to hand out
            const size_t i0 = j->next, i1 = std::min(j->n, i0 + j->grain);
            j->next = i1;
            if (i1 == j->n) jobs_.erase(jobs_.begin());
            lk.unlock();
            for (size_t i = i0; i < i1; i++) (*j->fn)(i);
            lk.lock();
            j->done += i1 - i0;
            if (j->done == j->n) j->cv.notify_one();        // the job object lives until its owner wakes up
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable work_;
    std::vector<Job*> jobs_;                                // sorted: priority descending, then submission order
    uint64_t seq_ = 0;
    bool stop_ = false;
};
Scheduler& pool() {
    static Scheduler p(std::max(1, g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency()));
    return p;
}
template <class F>
void parallel_for(size_t n, F fn) {
    int nt = (int)std::min<size_t>(n, (size_t)n_threads());
    if (nt <= 1) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::function<void(size_t)> f = fn;
    const double t0 = g_tm.on ? Timing::now() : 0;
    pool().run(n, (int64_t)Timing::t_progress() * 1024 - t_lane_id, f);
    if (g_tm.on) {
        g_turn_hold_us += (uint64_t)((Timing::now() - t0) * 1e3);
        g_turn_calls++;
    }
}

I128 load_i128(const uint8_t b[16]) {
    U128 v = 0;
    for (int i = 15; i >= 0; i--) v = (v << 8) | b[i];
    return (I128)v;
}
Fr fr_pow(Fr b, uint64_t e) { return h64::pow_u64(b, e); }
// Montgomery forms of small integers (digits, multiplicities, symbols) without a multiplication
const Fr* small_table() {
    static std::vector<Fr> t;
    static std::once_flag once;
    std::call_once(once, [] {
        t.resize(2048);
        Fr acc = h64::zero(), one = h64::one();
        for (size_t i = 0; i < t.size(); i++) { t[i] = acc; acc = h64::add(acc, one); }
    });
    return t.data();
}
inline Fr fr_small(U128 v) { return v < 2048 ? small_table()[(size_t)v] : h64::from_u128(v); }
int integer_log(U128 b, U128 n) {            // src/Utils.hs:91-92
    int r = 0;
    while (n >= b) { n /= b; r++; }
    return r;
}
U128 ipow(U128 b, int e) {
    U128 r = 1;
    while (e-- > 0) r *= b;
    return r;
}

// ------------------------------------------------------------------------------ RPWitness
struct RPW {                                   // Internal.hs:22-41
    Fr sc = h64::zero();
    std::vector<Fr> lin, nrm;
};
void vadd(std::vector<Fr>& a, const std::vector<Fr>& b) {
    if (b.size() > a.size()) a.resize(b.size(), h64::zero());
    for (size_t i = 0; i < b.size(); i++) a[i] = h64::add(a[i], b[i]);
}
void rpw_add(RPW& a, const RPW& b) {
    a.sc = h64::add(a.sc, b.sc);
    vadd(a.lin, b.lin);
    vadd(a.nrm, b.nrm);
}
RPW rpw_scale(const RPW& a, const Fr& s) {
    RPW r;
    r.sc = h64::mul(a.sc, s);
    r.lin.resize(a.lin.size());
    r.nrm.resize(a.nrm.size());
    for (size_t i = 0; i < a.lin.size(); i++) r.lin[i] = a.lin[i].is_zero() ? a.lin[i] : h64::mul(a.lin[i], s);
    for (size_t i = 0; i < a.nrm.size(); i++) r.nrm[i] = a.nrm[i].is_zero() ? a.nrm[i] : h64::mul(a.nrm[i], s);
    return r;
}

// ------------------------------------------------------------------------------ setup
struct Range {
    I128 mn, mx;
    U128 base = 2;
    bool is_shared = false, is_output = false, is_assumed = false, has_bit = false;
    std::vector<U128> coeffs;                  // baseCoeffs
    std::vector<Fr> coeffs_fr;                 // the same as field elements (cached)
};
struct Public {
    I128 amount, type;
    bool is_output;
};
// Phase1 entry (TypedReciprocal.hs:51-56).  kind: 'T' typing, 'I' inline, 'S' shared
struct Ph1 {
    char kind;
    int ind;
    U128 base = 0;
    Fr b, s;                                   // public: digit coefficient, symbol
    bool s_zero = true;
    bool io = false, ia = false;
    Fr d, m;                                   // private: digit (or type), multiplicity; for 'T' v in `m`
    int di = -1;                               // the digit as a small integer (< 2048) when it is one, else -1
    bool m_zero = true;                        // multiplicity known to be zero (shared digits, typing)
};
struct Ph2 {                                   // TypedReciprocal.hs:165-166
    bool isT;
    Fr d, m, u, v, r, c;
    bool m_zero = true, c_zero = true;         // known zeros let the per-entry sums skip multiplications
    int di = -1;
};

}  // namespace

struct bppp_rp {
    bppp_ctx* ctx = nullptr;
    bool binary = false;
    int arg = BPPP_ARG_NL;
    bool flag = false;                         // TRRP: hasTypes (typed || conserved); Binary: conserved
    std::vector<Range> rds;
    std::vector<Public> pubs;
    I128 net_pub = 0;                          // Binary: net public amount
    int fmt = tr::PREFIXED_P, root = tr::ROOT_EXP;
    std::string basis_seed;
    std::vector<U128> m_bases, sorted_bases;
    size_t nrm_len = 0, lin_len = 0, n_inputs = 0, num_rp_coms = 4;
    size_t rounds = 0, prover_rounds = 0, fin_n = 0, fin_l = 0, prover_fin_n = 0, prover_fin_l = 0;
    std::vector<Affine> pts;                   // h, g, then per kind
    Affine g;
    std::vector<uint8_t> table;                // [g | gs | hs] as bytes (device MSM base table)
    std::vector<std::vector<Ph1>> ph1_tmpl;    // per range: the public part of its Phase-1 entries (shared digits)
    std::vector<uint8_t> in_pts;               // TRRP: [g, hs0, hs1]; Binary: [g, h0]
    bppp_fb* fb = nullptr;                     // fixed-base tables over in_pts        (lane 0)
    bppp_gens* gens = nullptr;                 // resident [g | gs | hs] with window tables (lane 0)
    // extra lanes: sub-batches of one call run concurrently, each on its own context (stream) and
    // driver thread, so the host phases of one lane overlap the device work of another
    // page-locked staging buffers, one set per lane, grown on demand and reused across calls
    struct Pinned {
        void* p = nullptr; size_t cap = 0; bool pageable = false;
        void release() { if (p) { if (pageable) free(p); else bppp_pinned_free(p); } p = nullptr; cap = 0; }
    };
    std::vector<std::vector<Pinned>> pinned;
    std::vector<bppp_ctx*> lane_ctx;
    std::vector<bppp_fb*> lane_fb;
    std::vector<bppp_gens*> lane_gens;
    bppp_trrp* trrp = nullptr;                 // device scalar phases, lane 0 / other lanes
    std::vector<bppp_trrp*> lane_trrp;
    bool dev_phases = false;
    // Fiat-Shamir transcript on the device (bppp_rp_set_device_transcript / BPPP_DEVICE_TRANSCRIPT=1):
    // commitments are rendered and hashed where they are produced, the host only reads the challenges.
    // Needs the device scalar phases; bit-identical to the host transcript (tests).
    bool dev_transcript = false;
    // verify a lane's sub-batch by one random linear combination first, per-proof checks only if it fails
    bool batch_verify = false;
    size_t n_shared = 0;                       // shared-multiplicity coefficient slots computed on the device
    std::vector<bppp_dtr*> lane_vtr;           // verifier transcripts, one per lane (created on first use)
    std::mutex err_mu;
    std::string err;
};

namespace {

// ---- round counting (src/Bulletproof.hs:300-316, NormArgument.hs:165-178)
size_t round_reduce(size_t n) { return n / 2 + n % 2; }
void number_rounds_reduce(size_t n, size_t& r, size_t& out) {
    r = 0;
    while (n >= 5) { n = round_reduce(n); r++; }
    out = n;
}
void optimal_witness_size_nl(size_t n_len, size_t l_len, size_t& rounds, size_t& fn, size_t& fl) {
    size_t nR, n1, lR, l1;
    number_rounds_reduce(n_len, nR, n1);
    number_rounds_reduce(l_len, lR, l1);
    size_t r = std::max(nR, lR);
    for (size_t i = nR; i < r; i++) n1 = round_reduce(n1);
    for (size_t i = lR; i < r; i++) l1 = round_reduce(l1);
    if (n1 + l1 > 5) { rounds = r + 1; fn = round_reduce(n1); fl = round_reduce(l1); }
    else { rounds = r; fn = n1; fl = l1; }
}
// IP.NormLinear.optimalWitnessSize (InnerProductArgument.hs:253-267): nLen is the norm length
void optimal_witness_size_ip(size_t n_len, size_t l_len, size_t& rounds, size_t& fn, size_t& fl) {
    size_t n_even = (n_len + (n_len % 2)) / 2, nR, n1, lR, l1;
    number_rounds_reduce(n_even, nR, n1);                       // numberRoundsReduce' (Bulletproof.hs:306-308)
    if (n1 > 2) { nR++; n1 = round_reduce(n1); }
    number_rounds_reduce(l_len, lR, l1);
    size_t r = std::max(nR, lR);
    for (size_t i = nR; i < r; i++) n1 = round_reduce(n1);
    for (size_t i = lR; i < r; i++) l1 = round_reduce(l1);
    if (2 * n1 + l1 > 5) { rounds = r + 1; fn = 2 * round_reduce(n1); fl = round_reduce(l1); }
    else { rounds = r; fn = 2 * n1; fl = l1; }
}
void lengths_after(size_t n, size_t l, size_t rounds, size_t& fn, size_t& fl) {
    for (size_t i = 0; i < rounds; i++) { n = round_reduce(n); l = round_reduce(l); }
    fn = n; fl = l;
}

// ---- TypedReciprocal.makeRangeData (TypedReciprocal.hs:89-115)
bool make_range_trrp(Range& rd) {
    if (!(rd.mx > rd.mn) || rd.base <= 1) return false;
    U128 w = (U128)(rd.mx - rd.mn);
    U128 b = rd.base;
    int n1 = integer_log(b, w - 1);
    rd.has_bit = ((w - 1) % (b - 1)) != 0;
    std::vector<U128> tail;
    for (int i = 1; i <= n1; i++) tail.push_back(ipow(b, n1 - i));
    std::vector<U128> bs;
    U128 bn = ipow(b, n1);
    if (!rd.has_bit) bs.push_back((w - bn) / (b - 1));
    else if (w < 2 * bn) bs.push_back(w - bn);
    else {
        U128 bn1 = 1 + w / (2 * (b - 1)) - (bn - 1) / (b - 1);
        bs.push_back(w - bn1 * (b - 1) - bn);
        bs.push_back(bn1);
    }
    bs.insert(bs.end(), tail.begin(), tail.end());
    rd.coeffs = rd.is_assumed ? std::vector<U128>() : bs;
    rd.coeffs_fr.clear();
    for (auto c : rd.coeffs) rd.coeffs_fr.push_back(h64::from_u128(c));
    return true;
}
// ---- Binary.makeRangeData (Binary.hs:48-54)
bool make_range_bin(Range& rd) {
    if (!(rd.mx > rd.mn)) return false;
    U128 w = (U128)(rd.mx - rd.mn);
    int n1 = integer_log(2, w - 1);
    rd.base = 2;
    rd.coeffs.clear();
    rd.coeffs.push_back(w - ipow(2, n1));
    for (int i = 1; i <= n1; i++) rd.coeffs.push_back(ipow(2, n1 - i));
    return true;
}

void finish_setup(bppp_rp* s) {
    // device base table [g | gs | hs] and the generator triple of the input commitments
    size_t P0 = 1 + s->nrm_len + s->lin_len;
    s->table.resize(P0 * 64);
    size_t hs_off, gs_off;
    if (s->binary) { hs_off = 2; gs_off = 4; }                     // [h,g,h0,h1] ++ gs (Binary.hs:147-148)
    else { hs_off = 2; gs_off = 2 + s->lin_len; }                  // h : g : hs ++ gs (TypedReciprocal.hs:334,348-349)
    s->g = s->pts[1];
    memcpy(&s->table[0], &s->pts[1], 64);
    for (size_t i = 0; i < s->nrm_len; i++) memcpy(&s->table[64 * (1 + i)], &s->pts[gs_off + i], 64);
    for (size_t i = 0; i < s->lin_len; i++) memcpy(&s->table[64 * (1 + s->nrm_len + i)], &s->pts[hs_off + i], 64);
    if (s->binary) {             // value on g, blind on h0 (Internal.hs:53-54, app/Main.hs:315)
        s->in_pts.resize(2 * 64);
        memcpy(&s->in_pts[0], &s->pts[1], 64);
        memcpy(&s->in_pts[64], &s->pts[2], 64);
    } else {                     // value on g, type on hs[0] (= ht), blind on hs[1] (Internal.hs:56-57)
        s->in_pts.resize(3 * 64);
        memcpy(&s->in_pts[0], &s->pts[1], 64);
        memcpy(&s->in_pts[64], &s->pts[2], 64);
        memcpy(&s->in_pts[128], &s->pts[3], 64);
    }
}

// scalars of commitRPW over the table [g | gs | hs] (Internal.hs:43-48; dotWith pads with zeros /
// identity points, so entries beyond the generator lists contribute nothing)
void commit_scalars(const bppp_rp* s, const RPW& w, uint8_t* out) {
    size_t P0 = 1 + s->nrm_len + s->lin_len;
    memset(out, 0, P0 * 32);
    h64::to_bytes(out, w.sc);
    for (size_t i = 0; i < w.nrm.size() && i < s->nrm_len; i++)
        if (!w.nrm[i].is_zero()) h64::to_bytes(out + 32 * (1 + i), w.nrm[i]);
    for (size_t i = 0; i < w.lin.size() && i < s->lin_len; i++)
        if (!w.lin[i].is_zero()) h64::to_bytes(out + 32 * (1 + s->nrm_len + i), w.lin[i]);
}

// qPowers' of the Weighted instances: NL powers' (q^2) (NormArgument.hs:147-148);
// IP norm powers' (-(q^2)) (InnerProductArgument.hs:230-231)
Fr q0_of(int arg, const Fr& q) { return arg == BPPP_ARG_IP ? h64::neg(h64::sqr(q)) : h64::sqr(q); }
std::vector<Fr> q_powers(int arg, const Fr& q, size_t n) { return h64::powers1(q0_of(arg, q), n); }

// ---- blinding helpers (Internal.hs:134-195)
std::vector<Fr> pad_right(size_t n, std::vector<Fr> xs) {
    xs.resize(n, h64::zero());
    return xs;
}
RPW blind_witness(tr::Zkpt& zk, int n, int k, const std::vector<Fr>& ls, const std::vector<Fr>& ns) {
    int n_bls = (k == 1) ? 2 * n - 1 : 2 * n - k + 1;
    std::vector<Fr> bls;
    for (int i = 0; i < n_bls; i++) bls.push_back(zk.random());
    bls.insert(bls.begin() + (2 * n - k), h64::zero());
    bls = pad_right(2 * n + 1, bls);
    RPW w;
    w.sc = bls[0];
    w.lin.assign(bls.begin() + 1, bls.end());
    w.lin.insert(w.lin.end(), ls.begin(), ls.end());
    w.nrm = ns;
    return w;
}
RPW blind_err_witness(tr::Zkpt& zk, int n, const std::vector<Fr>& es, const std::vector<Fr>& ls, const std::vector<Fr>& ns) {
    std::vector<Fr> bls;
    for (int i = 0; i < n + 1; i++) bls.push_back(zk.random());
    bls.insert(bls.begin() + n, h64::zero());
    bls.insert(bls.end(), es.begin(), es.end());
    bls = pad_right(2 * n + 1, bls);
    RPW w;
    w.sc = bls[0];
    w.lin.assign(bls.begin() + 1, bls.end());
    w.lin.insert(w.lin.end(), ls.begin(), ls.end());
    w.nrm = ns;
    return w;
}
std::vector<Fr> scale_errs(int n, const Fr& k, const std::vector<Fr>& xs) {
    std::vector<Fr> o = xs;
    for (size_t i = n + 1; i < xs.size() && i < (size_t)(n + 1 + n - 2); i++) o[i] = h64::mul(k, xs[i]);
    return o;
}
RPW blind_blinding_term(const RPW& bl, const Fr& tC, const Fr& r0, const Fr& r0i, const Fr& r1, const Fr& r1i,
                        const std::vector<Fr>& errs, const std::vector<const RPW*>& wits, const Fr& input_bl) {
    (void)r0;
    Fr blT = bl.lin[0];
    Fr rs_inv = h64::mul(r0i, r1i);
    int n = (int)wits.size();
    const RPW* wit_err = wits[n - 1];
    auto take = [](const std::vector<Fr>& v, size_t k) {
        return std::vector<Fr>(v.begin(), v.begin() + std::min(k, v.size()));
    };
    std::vector<std::vector<Fr>> rows;
    for (int i = 0; i < n - 1; i++) {
        std::vector<Fr> r = {wits[i]->sc};
        auto t = take(wits[i]->lin, 2 * n);
        r.insert(r.end(), t.begin(), t.end());
        rows.push_back(r);
    }
    {
        std::vector<Fr> r = {wit_err->sc};
        auto t = pad_right(2 * n, take(wit_err->lin, n + 1));
        r.insert(r.end(), t.begin(), t.end());
        rows.push_back(r);
    }
    for (auto& r : rows)
        for (size_t i = 2; i < r.size(); i++) r[i] = h64::neg(r[i]);
    std::vector<Fr> errs1;
    errs1.push_back(h64::neg(h64::sub(errs[0], h64::mul(tC, blT))));
    for (size_t i = 1; i < errs.size(); i++) errs1.push_back(h64::neg(h64::mul(rs_inv, errs[i])));
    std::vector<std::vector<Fr>> table;
    table.push_back(errs1);
    Fr rs_tc = h64::mul(rs_inv, tC);
    for (auto& r : rows) {
        std::vector<Fr> a;
        a.push_back(h64::add(h64::mul(rs_inv, r[0]), h64::mul(rs_tc, r[1])));      // addConsts
        a.insert(a.end(), r.begin() + 2, r.end());
        table.push_back(scale_errs(n, r1i, a));
    }
    for (auto& r : table) r.insert(r.begin() + std::min<size_t>(2 * n - 1, r.size()), h64::zero());   // insertAt (2n-1) 0
    std::map<size_t, Fr> diag;                                                     // sumDiagonals
    for (size_t a = 0; a < table.size(); a++)
        for (size_t b = 0; b < table[a].size(); b++) {
            auto it = diag.find(a + b);
            if (it == diag.end()) diag[a + b] = table[a][b];
            else it->second = h64::add(it->second, table[a][b]);
        }
    std::vector<Fr> sd;
    for (auto& kv : diag) sd.push_back(kv.second);
    sd.erase(sd.begin() + (2 * n - 1));                                            // removeAt (2n-1)
    sd.resize(std::min<size_t>(sd.size(), 2 * n));                                 // take (2n)
    std::vector<Fr> bl_errs = scale_errs(n, r1, sd);
    bl_errs.back() = h64::sub(bl_errs.back(), h64::dbl(input_bl));                 // appLast
    RPW out;
    out.sc = h64::neg(bl_errs[0]);
    out.lin.push_back(blT);
    out.lin.insert(out.lin.end(), bl_errs.begin() + 1, bl_errs.end());
    out.lin.insert(out.lin.end(), bl.lin.begin() + 1, bl.lin.end());
    out.nrm = bl.nrm;
    return out;
}

// ---- TRRP phases
std::vector<U128> digits_trrp(const Range& rd, U128 n) {                // TypedReciprocal.hs:120-122
    std::vector<U128> out;
    for (size_t i = 0; i < rd.coeffs.size(); i++) {
        U128 base = (rd.has_bit && i == 0) ? 2 : rd.base;
        U128 b = rd.coeffs[i];
        U128 d = b ? std::min<U128>(base - 1, n / b) : (base - 1);        // n `quot` 0 never happens for valid ranges
        n -= d * b;
        out.push_back(d);
    }
    return out;
}
// value (Fr canonical bytes) - min as a small integer; false when out of range
bool adjust_value(const Range& rd, const uint8_t val[32], U128& n_adj) {
    Fr v = h64::from_bytes(val);
    Fr a = h64::sub(v, h64::from_i128(rd.mn));
    uint64_t c[4];
    h64::to_canon(c, a);
    if (c[2] | c[3]) return false;
    n_adj = ((U128)c[1] << 64) | c[0];
    return n_adj < (U128)(rd.mx - rd.mn);
}
// makePhase1s (TypedReciprocal.hs:128-152); prover = false builds the verifier's empty-witness copy
bool make_phase1s(int ind, const Range& rd, const uint8_t* val, bool prover, std::vector<Ph1>& out, bool& has_ms,
                  std::vector<U128>& ms_out) {
    has_ms = false;
    if (rd.is_assumed) return true;
    std::vector<U128> ds;
    if (prover) {
        U128 n_adj;
        if (!adjust_value(rd, val, n_adj)) return false;
        ds = digits_trrp(rd, n_adj);
    } else ds.assign(rd.coeffs.size(), 0);
    size_t base = (size_t)rd.base;
    std::vector<U128> ms;
    if (prover) {
        std::vector<U128> cnt(base, 0);
        for (size_t i = rd.has_bit ? 1 : 0; i < ds.size(); i++)
            if (ds[i] < base) cnt[(size_t)ds[i]]++;
        if (rd.has_bit) ms.push_back(ds[0]);
        for (size_t v = 1; v < base; v++) ms.push_back(cnt[v]);
    } else ms.assign(base - 1 + (rd.has_bit ? 1 : 0), 0);
    std::vector<U128> ns;
    if (rd.has_bit) ns.push_back(1);
    for (size_t v = 1; v < base; v++) ns.push_back(v);
    auto base_at = [&](size_t i) { return (rd.has_bit && i == 0) ? (U128)2 : rd.base; };
    if (rd.is_shared) {
        for (size_t i = 0; i < rd.coeffs.size(); i++) {
            Ph1 p;
            p.kind = 'S'; p.ind = ind; p.base = base_at(i);
            p.b = rd.coeffs_fr[i]; p.d = fr_small(ds[i]); p.m = h64::zero();
            p.di = ds[i] < 2048 ? (int)ds[i] : -1;
            out.push_back(p);
        }
        has_ms = true;
        ms_out = ms;
        return true;
    }
    size_t L = std::max(std::max(rd.coeffs.size(), ds.size()), std::max(ms.size(), ns.size()));
    for (size_t i = 0; i < L; i++) {
        Ph1 p;
        p.kind = 'I'; p.ind = ind; p.base = base_at(i);
        p.b = i < rd.coeffs.size() ? rd.coeffs_fr[i] : h64::zero();
        p.d = fr_small(i < ds.size() ? ds[i] : 0);
        p.m = fr_small(i < ms.size() ? ms[i] : 0);
        p.m_zero = p.m.is_zero();
        { U128 dv = i < ds.size() ? ds[i] : 0; p.di = dv < 2048 ? (int)dv : -1; }
        U128 sym = i < ns.size() ? ns[i] : 0;
        p.s = fr_small(sym);
        p.s_zero = (sym == 0);
        out.push_back(p);
    }
    return true;
}
std::map<U128, Fr> make_base_map(const bppp_rp* s, const Fr& x) {     // TypedReciprocal.hs:353
    std::map<U128, Fr> m;
    Fr x2 = h64::sqr(x), cur = h64::mul(x2, x);
    for (auto b : s->sorted_bases) { m[b] = cur; cur = h64::mul(cur, x2); }
    return m;
}
// makePhase2s (TypedReciprocal.hs:171-195)
std::vector<Ph2> make_phase2s(bool prover, const Fr& e, const Fr& e_inv, const Fr& x, std::map<U128, Fr>& bm,
                              const std::vector<Ph1>& ph1s) {
    size_t n = ph1s.size();
    std::vector<Ph2> out(n);
    std::vector<Fr> ss(n, h64::zero()), vs(n);
    Fr x2 = h64::sqr(x);
    std::map<int, Fr> xpow;                      // x^(2(ind+1)) per range index
    // reciprocals 1/(e + d): digits are small integers, so invert the distinct denominators once
    // (batchInverse over all entries in the reference, TypedReciprocal.hs:194); entries whose
    // denominator is not a small digit (types) are inverted individually.
    int dmax = -1;
    std::vector<size_t> slow;
    std::vector<Fr> slow_den;
    bool any_ss = false;
    int last_ind = -1;
    Fr last_xp = h64::zero();
    U128 last_base = 0;
    Fr last_v = h64::zero();
    for (size_t i = 0; i < n; i++) {
        const Ph1& p = ph1s[i];
        if (p.ind != last_ind) {
            auto it = xpow.find(p.ind);
            if (it == xpow.end()) it = xpow.emplace(p.ind, fr_pow(x2, (uint64_t)p.ind + 1)).first;
            last_ind = p.ind;
            last_xp = it->second;
        }
        const Fr& xp = last_xp;
        Ph2& o = out[i];
        if (p.kind == 'T') {
            Fr xpp = p.io ? h64::neg(x) : x;
            if (prover) { slow.push_back(i); slow_den.push_back(h64::add(e, p.d)); }
            o.isT = true; o.d = p.d; o.m = h64::zero(); o.u = p.ia ? h64::zero() : xp; o.v = xpp;
            vs[i] = xpp;
        } else {
            if (p.base != last_base) { last_base = p.base; last_v = bm[p.base]; }
            if (prover) {
                if (p.di >= 0) dmax = std::max(dmax, p.di);
                else { slow.push_back(i); slow_den.push_back(h64::add(e, p.d)); }
            }
            if (p.kind == 'I' && !p.s_zero) { ss[i] = h64::add(e, p.s); any_ss = true; }
            o.isT = false; o.d = p.d; o.m = (p.kind == 'I') ? p.m : h64::zero();
            o.m_zero = (p.kind != 'I') || p.m_zero;
            o.di = p.di;
            o.u = h64::mul(xp, p.b); o.v = last_v;
            vs[i] = last_v;
        }
    }
    std::vector<Fr> tab;
    if (prover) {
        tab.resize(dmax + 1);
        Fr cur = e;
        const Fr one = h64::one();
        for (int d = 0; d <= dmax; d++) { tab[d] = cur; cur = h64::add(cur, one); }
        h64::batch_inv(tab.data(), tab.size());
        h64::batch_inv(slow_den.data(), slow_den.size());
    }
    if (any_ss) h64::batch_inv(ss.data(), n);
    for (size_t i = 0; i < n; i++) {
        const Ph1& p = ph1s[i];
        out[i].r = h64::zero();
        if (prover && p.kind != 'T' && p.di >= 0) out[i].r = tab[p.di];            // ps = 1
        out[i].c_zero = ss[i].is_zero();
        out[i].c = out[i].c_zero ? h64::zero() : h64::mul(vs[i], h64::sub(e_inv, ss[i]));
    }
    for (size_t k = 0; k < slow.size(); k++) {
        const Ph1& p = ph1s[slow[k]];
        out[slow[k]].r = (p.kind == 'T') ? h64::mul(p.m, slow_den[k]) : slow_den[k];   // T: ps = v (the amount)
    }
    return out;
}
std::vector<Fr> make_shared_coeffs(const Fr& e, const Fr& e_inv, const std::vector<U128>& m_bases, std::map<U128, Fr>& bm) {
    std::vector<Fr> xs, ss;                                          // TypedReciprocal.hs:204-206
    for (auto b : m_bases)
        for (U128 s = 1; s < b; s++) { xs.push_back(bm[b]); ss.push_back(h64::add(e, h64::from_u128(s))); }
    h64::batch_inv(ss.data(), ss.size());
    for (size_t i = 0; i < xs.size(); i++) xs[i] = h64::mul(xs[i], h64::sub(e_inv, ss[i]));
    return xs;
}
std::vector<Fr> make_error_terms(const Fr& e, const Fr& xq, const std::vector<Fr>& shared_cs, const std::vector<Fr>& bls_ms,
                                 const std::vector<Ph2>& ph2s, const std::vector<Fr>& q2s, const std::vector<Fr>& bls) {
    std::vector<Fr> tot(6, h64::zero());                             // TypedReciprocal.hs:217-233
    Fr aug = h64::zero();
    for (size_t i = 0; i < shared_cs.size() && i < bls_ms.size(); i++) aug = h64::add(aug, h64::mul(shared_cs[i], bls_ms[i]));
    tot[3] = h64::dbl(aug);
    using namespace h64;
    // the doubled parts are accumulated undoubled and doubled once at the end
    Fr h1 = zero(), h2 = zero(), h3 = zero(), h4 = zero(), h5 = zero();
    for (size_t i = 0; i < ph2s.size() && i < q2s.size() && i < bls.size(); i++) {
        const Ph2& o = ph2s[i];
        const Fr& q2 = q2s[i];
        const Fr& bl = bls[i];
        Fr rC = o.isT ? mul(xq, add(o.u, q2)) : o.u;
        Fr q2e = mul(q2, e);
        Fr dC = add(o.v, q2e);
        Fr q2d = mul(q2, o.d), q2r = mul(q2, o.r);
        Fr qd = add(q2d, dC), qr = add(q2r, rC);
        Fr q2bl = mul(q2, bl);
        tot[0] = add(tot[0], mul(q2bl, bl));                                   // err0 = q2 bl^2
        h2 = add(h2, mul(bl, qd));                                             // err2 = q2 m^2 + 2 bl qd
        h3 = add(h3, mul(bl, qr));                                             // err3 = 2 (bl qr + m qd)
        // err4 = (q2 d^2 + 2 d dC) + 2 (bl c + m qr) = d (q2 d + 2 dC) + ...
        tot[4] = add(tot[4], mul(o.d, add(q2d, dbl(dC))));
        // err6 = (q2 r^2 + 2 r rC) + 2 c d = r (q2 r + 2 rC) + 2 c d
        tot[5] = add(tot[5], mul(o.r, add(q2r, dbl(rC))));
        if (!o.m_zero) {
            h1 = add(h1, mul(q2bl, o.m));                                      // err1 = 2 q2 m bl
            tot[2] = add(tot[2], mul(q2, sqr(o.m)));
            h3 = add(h3, mul(o.m, qd));
            h4 = add(h4, mul(o.m, qr));
        }
        if (!o.c_zero) {
            h4 = add(h4, mul(bl, o.c));
            h5 = add(h5, mul(o.c, o.d));
        }
    }
    tot[1] = add(tot[1], dbl(h1));
    tot[2] = add(tot[2], dbl(h2));
    tot[3] = add(tot[3], dbl(h3));
    tot[4] = add(tot[4], dbl(h4));
    tot[5] = add(tot[5], dbl(h5));
    return tot;
}
// the part of makePublicConsts' scalar that does not involve the norm entries (range minima, public amounts)
Fr public_consts_z_trrp(const bppp_rp* s, const Fr& e, const Fr& x, const Fr& two_t5) {
    using namespace h64;
    Fr x2 = sqr(x), xp = x2, acc = zero();
    for (auto& rd : s->rds) {
        if (!rd.is_assumed) acc = add(acc, mul(from_i128(rd.mn), xp));
        xp = mul(xp, x2);
    }
    Fr z = neg(mul(two_t5, acc));
    if (s->flag) {
        std::vector<Fr> rs;
        for (auto& p : s->pubs) rs.push_back(add(e, from_i128(p.type)));
        batch_inv(rs.data(), rs.size());
        Fr sum = zero();
        for (size_t i = 0; i < s->pubs.size(); i++) {
            Fr term = mul(rs[i], from_i128(s->pubs[i].amount));
            sum = s->pubs[i].is_output ? sub(sum, term) : add(sum, term);
        }
        z = sub(z, mul(mul(two_t5, x), sum));
    }
    return z;
}
RPW make_public_consts_trrp(const bppp_rp* s, const Fr& e, const Fr& e_inv, const Fr& x, const Fr& xq, const Fr& q0,
                            const Fr& q0_inv, const Fr& t, const std::vector<Ph2>& ph2s) {
    using namespace h64;                                             // TypedReciprocal.hs:236-263
    Fr t2 = sqr(t), t3 = mul(t2, t), t4 = sqr(t2), t5 = mul(t4, t);
    Fr two_t5 = dbl(t5);
    Fr z = public_consts_z_trrp(s, e, x, two_t5);
    // p_i = t^2 (e + qi2 v) + t^3 rC + t^4 qi2 c  with rC = qi2 u (digits) or x'(qi2 u + 1) (types)
    //     = const + qi2 * (t^2 v + t^3 u' + t^4 c),   u' = u or x' u,  const = t^2 e (+ t^3 x' for types)
    // sum_i t^5 p2C_i = 2 t^5 (sum q2_i + eInv sum v_i) over the digit entries
    RPW out;
    out.nrm.reserve(ph2s.size());
    const Fr t2e = mul(t2, e), t3xq = mul(t3, xq), constT = add(t2e, t3xq);
    Fr q2 = q0, qi2 = q0_inv, ts0 = zero(), sum_q2 = zero(), sum_v = zero();
    Fr last_v = zero(), last_t2v = zero();
    bool have_v = false;
    for (auto& o : ph2s) {
        if (!have_v || !(o.v == last_v)) { last_v = o.v; last_t2v = mul(t2, o.v); have_v = true; }
        Fr inner = last_t2v;
        if (!o.u.is_zero()) inner = add(inner, mul(o.isT ? t3xq : t3, o.u));
        if (!o.c_zero) inner = add(inner, mul(t4, o.c));
        Fr p = add(o.isT ? constT : t2e, mul(qi2, inner));
        ts0 = add(ts0, mul(q2, sqr(p)));
        if (!o.isT) { sum_q2 = add(sum_q2, q2); sum_v = add(sum_v, o.v); }
        out.nrm.push_back(p);
        q2 = mul(q2, q0);
        qi2 = mul(qi2, q0_inv);
    }
    ts0 = add(ts0, mul(two_t5, add(sum_q2, mul(e_inv, sum_v))));
    out.sc = add(z, ts0);
    return out;
}
std::vector<Fr> input_coeffs_trrp(const bppp_rp* s, const Fr& x, const Fr& q0) {   // TypedReciprocal.hs:325-328
    std::vector<Fr> out;
    Fr x2 = h64::sqr(x), xp = x2, qp = q0;
    for (auto& rd : s->rds) {
        Fr v = rd.is_assumed ? h64::zero() : xp;
        if (s->flag) v = h64::add(qp, v);
        out.push_back(v);
        xp = h64::mul(xp, x2);
        qp = h64::mul(qp, q0);
    }
    return out;
}
std::vector<Fr> make_bp_coeffs(bool has_types, const Fr& xq, const Fr& r0, const Fr& r1, const Fr& t, const std::vector<Fr>& cs) {
    using namespace h64;                                             // TypedReciprocal.hs:391-396
    Fr rs = mul(r0, r1), t2 = sqr(t), t3 = mul(t2, t), t4 = sqr(t2), t6 = sqr(t3);
    std::vector<Fr> o = {has_types ? neg(xq) : zero(), mul(rs, t), mul(rs, t2), mul(rs, t3), mul(r0, t4), mul(rs, t6)};
    Fr two_t3 = dbl(t3);
    for (auto& c : cs) o.push_back(mul(two_t3, c));
    return o;
}
std::vector<Ph1> ph1s_verifier(const bppp_rp* s) {
    std::vector<Ph1> out;
    if (s->flag)
        for (size_t i = 0; i < s->rds.size(); i++) {
            Ph1 p;
            p.kind = 'T'; p.ind = (int)i; p.io = s->rds[i].is_output; p.ia = s->rds[i].is_assumed;
            p.d = p.m = h64::zero();
            out.push_back(p);
        }
    for (size_t i = 0; i < s->rds.size(); i++) {
        bool hm;
        std::vector<U128> ms;
        make_phase1s((int)i, s->rds[i], nullptr, false, out, hm, ms);
    }
    return out;
}

// ---- Binary phases
std::vector<Fr> input_coeffs_brp(const bppp_rp* s, const Fr& x) {     // Binary.hs:128-130
    std::vector<Fr> out;
    Fr x2 = h64::sqr(x), xp = x2;
    for (auto& rd : s->rds) {
        Fr v = rd.is_assumed ? h64::zero() : xp;
        if (s->flag) v = h64::add(v, rd.is_output ? h64::neg(x) : x);
        out.push_back(v);
        xp = h64::mul(xp, x2);
    }
    return out;
}
RPW public_consts_brp(const bppp_rp* s, const Fr& x, const Fr& q0, const Fr& q0_inv) {   // Binary.hs:73-97
    using namespace h64;
    Fr x2 = sqr(x), xp = x2, macc = zero();
    std::vector<Fr> bss;
    for (auto& rd : s->rds) {
        if (!rd.is_assumed) {
            for (auto b : rd.coeffs) bss.push_back(mul(xp, from_u128(b)));
            macc = add(macc, mul(from_i128(rd.mn), xp));
        }
        xp = mul(xp, x2);
    }
    Fr net = s->flag ? mul(neg(x), from_i128(s->net_pub)) : zero();
    Fr z = neg(dbl(add(net, macc)));
    Fr half = inv(from_u64(2));
    RPW out;
    Fr q2 = q0, q2i = q0_inv, sc = z;
    for (auto& bx : bss) {
        Fr p = sub(mul(bx, q2i), half);
        sc = add(sc, mul(q2, sqr(p)));
        q2 = mul(q2, q0);
        q2i = mul(q2i, q0_inv);
        out.nrm.push_back(p);
    }
    out.sc = sc;
    return out;
}

// ------------------------------------------------------------------------------ batch state
struct Proof {
    tr::Zkpt zk;
    std::vector<Ph1> ph1s;
    std::vector<std::pair<U128, std::vector<Fr>>> base_mss;
    std::vector<RPW> n_wits;
    RPW dm, m, r, bl, d;
    std::vector<Ph2> ph2s;
    Fr e, x, r0, q, xq, r1, q0, t, e_inv, r0_inv, q_inv, q0_inv, r1_inv;
    std::vector<Fr> shared_cs, cs, bls_lin, bls_nrm;
    std::map<U128, Fr> base_map;
    RPW pub, wit;
    Fr s_bl;
    bool ok = true;
};

int fail(bppp_rp* s, int code, const std::string& msg) {
    std::lock_guard<std::mutex> lk(s->err_mu);
    s->err = msg;
    return code;
}
struct Lane {
    bppp_ctx* ctx;
    bppp_fb* fb;
    bppp_gens* gens;
    int threads;                               // host threads this lane may use
    int index;                                 // which set of staging buffers
    bppp_trrp* trrp = nullptr;                 // device scalar phases (TypedReciprocal + norm-linear argument)
};
enum { PB_IN = 0, PB_SC1, PB_SC2, PB_C1, PB_C2, PB_NCOMS, PB_Q, PB_S, PB_W, PB_L, PB_C, PB_X, PB_R, PB_E, PB_V0, PB_V1, PB_V2, PB_V3,
       PB_V4, PB_V5, PB_V6, PB_V7, PB_V8, PB_CH, PB_SCLIN, PB_BLN, PB_SMALL, PB_CHT, PB_CHI, PB_SHC, PB_COUNT };
// uninitialised page-locked buffer `slot` of the lane, at least `bytes` long
uint8_t* lane_buf(bppp_rp* s, const Lane& ln, int slot, size_t bytes) {
    auto& pb = s->pinned[ln.index][slot];
    if (pb.cap < bytes) {
        pb.release();
        pb.cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        pb.pageable = false;
        if (bppp_pinned_alloc(want, &pb.p)) {                             // fall back to pageable memory
            pb.p = malloc(want);
            pb.pageable = true;
        }
        pb.cap = pb.p ? want : 0;
    }
    return (uint8_t*)pb.p;
}
thread_local int t_lane_threads = 0;
thread_local int t_lane_id = 0;
thread_local bool t_is_lane0 = true;
const char* ctx_err(const Lane& ln) { return bppp_last_error(ln.ctx); }

// run the argument (proveBPM, src/Bulletproof.hs:357-359) for the whole batch
int run_argument(bppp_rp* s, const Lane& ln, std::vector<Proof>& P, size_t rounds, const uint8_t* q, const uint8_t* sc,
                 const uint8_t* w, const uint8_t* l, const uint8_t* c,
                 uint8_t* responses, uint8_t* finals, size_t fin_n, size_t fin_l, bool device_witness = false,
                 bool device_transcript = false) {
    const size_t B = P.size();
    bppp_nl* h = nullptr;
    int rc = device_witness ? bppp_nl_create_trrp(ln.trrp, q, sc, l, c, &h) : bppp_nl_create_gens(ln.gens, s->arg, B, q, sc, w, l, c, &h);
    if (rc) return fail(s, rc, std::string("bppp_nl_create: ") + ctx_err(ln));
    g_tm.lap("nl_create");
    const bool dev_rounds = !(getenv("BPPP_DEVICE_ROUNDS") && !atoi(getenv("BPPP_DEVICE_ROUNDS")));
    if (device_transcript && dev_rounds && s->arg == BPPP_ARG_NL) {
        // the whole argument as one stream of launches (bppp_nl_prove_device): no per-round host work at all
        size_t cn = 0, cl = 0;
        lengths_after(s->nrm_len, s->lin_len, rounds, cn, cl);
        if (cn != fin_n || cl != fin_l) { bppp_nl_destroy(h); return fail(s, BPPP_ERR_STATE, "final witness lengths differ from infoRP"); }
        uint8_t* fw = lane_buf(s, ln, PB_X, B * cn * 32 + 32);
        uint8_t* fl = lane_buf(s, ln, PB_R, B * cl * 32 + 32);
        rc = bppp_nl_prove_device(h, rounds, responses, nullptr, nullptr, fw, fl);
        bppp_nl_destroy(h);
        g_tm.lap("nl_rounds_device");
        g_tm.dump("prove");
        if (t_lane_threads_is_main()) dump_sections("prove, all lanes", B);
        if (rc) return fail(s, rc, std::string("bppp_nl_prove_device: ") + ctx_err(ln));
        for (size_t b = 0; b < B; b++) {       // getWitness: norm scalars then linear scalars (RangeProof.hs:65)
            memcpy(finals + 32 * b * (cn + cl), &fw[32 * b * cn], 32 * cn);
            memcpy(finals + 32 * (b * (cn + cl) + cn), &fl[32 * b * cl], 32 * cl);
        }
        return BPPP_OK;
    }
    uint8_t* X = lane_buf(s, ln, PB_X, B * 64);
    uint8_t* R = lane_buf(s, ln, PB_R, B * 64);
    uint8_t* E = lane_buf(s, ln, PB_E, B * 32);
    for (size_t r = 0; r < rounds; r++) {
        rc = device_transcript ? bppp_nl_round_challenge(h, X, R, E) : bppp_nl_round_commit(h, X, R);
        g_tm.lap("nl_commit");
        if (rc) { bppp_nl_destroy(h); return fail(s, rc, std::string("bppp_nl_round_commit: ") + ctx_err(ln)); }
        if (device_transcript) {                                    // the challenge came back with the commitments
            for (size_t b = 0; b < B; b++) {
                uint8_t* o = responses + 128 * (b * rounds + (rounds - 1 - r));
                memcpy(o, &X[64 * b], 64);
                memcpy(o + 64, &R[64 * b], 64);
            }
        } else
        parallel_for((B + 1) / 2, [&](size_t pi) {                  // two transcripts per task (two-stream SHA)
            const size_t b0 = 2 * pi, nb = std::min<size_t>(2, B - b0);
            uint8_t xr[2][128];
            Fr e[2];
            Sect sect;
            for (size_t j = 0; j < nb; j++) {
                memcpy(xr[j], &X[64 * (b0 + j)], 64);
                memcpy(xr[j] + 64, &R[64 * (b0 + j)], 64);
            }
            if (nb == 2) tr::Zkpt::oracle_pair(P[b0].zk, xr[0], P[b0 + 1].zk, xr[1], 2, &e[0], &e[1]);
            else P[b0].zk.oracle(xr[0], 2, &e[0], 1);                // e <- head <$> oracle [ac, bc]
            sect.lap(S_ORACLE);
            for (size_t j = 0; j < nb; j++) {
                h64::to_bytes(&E[32 * (b0 + j)], e[j]);
                // responses are consed: newest first (Bulletproof.hs:357-359)
                memcpy(responses + 128 * ((b0 + j) * rounds + (rounds - 1 - r)), xr[j], 128);
            }
        });
        g_tm.lap("round_hash");
        rc = bppp_nl_round_fold(h, E);
        g_tm.lap("nl_fold");
        if (rc) { bppp_nl_destroy(h); return fail(s, rc, std::string("bppp_nl_round_fold: ") + ctx_err(ln)); }
    }
    size_t cn = 0, cl = 0;
    bppp_nl_lengths(h, &cn, &cl);
    if (cn != fin_n || cl != fin_l) { bppp_nl_destroy(h); return fail(s, BPPP_ERR_STATE, "final witness lengths differ from infoRP"); }
    std::vector<uint8_t> fs(B * 32), fw(B * cn * 32), fl(B * cl * 32);
    rc = bppp_nl_final(h, fs.data(), fw.data(), fl.data());
    bppp_nl_destroy(h);
    g_tm.lap("nl_final");
    g_tm.dump("prove");
    if (t_lane_threads_is_main()) dump_sections("prove, all lanes", B);
    if (rc) return fail(s, rc, std::string("bppp_nl_final: ") + ctx_err(ln));
    for (size_t b = 0; b < B; b++) {       // getWitness: norm scalars then linear scalars (RangeProof.hs:65)
        memcpy(finals + 32 * b * (cn + cl), &fw[32 * b * cn], 32 * cn);
        memcpy(finals + 32 * (b * (cn + cl) + cn), &fl[32 * b * cl], 32 * cl);
    }
    return BPPP_OK;
}

}  // namespace

// BPPP_SAMPLE=file: a SIGPROF sampling profiler of the whole process (1 kHz of CPU time, program
// counters only) for finding where the host cores go on a box without perf; tools/sample_report.py
// resolves the counts against the library's symbol table.
// Compiled only with -DBPPP_SAMPLER (a development build): a release library installs no signal handler
// and no interval timer -- SIGPROF / ITIMER_PROF belong to the host application (GHC's RTS uses them).
namespace {
#ifdef BPPP_SAMPLER
std::atomic<size_t> g_ns{0};
void** g_samples = nullptr;
void** g_callers = nullptr;           // the frame's return address (meaningful in a -fno-omit-frame-pointer build)
const size_t kMaxSamples = 1 << 20;
void sample_handler(int, siginfo_t*, void* uc) {
    size_t i = g_ns.fetch_add(1, std::memory_order_relaxed);
    if (i >= kMaxSamples) return;
    const greg_t* r = ((ucontext_t*)uc)->uc_mcontext.gregs;
    g_samples[i] = (void*)r[REG_RIP];
    const uintptr_t bp = (uintptr_t)r[REG_RBP], sp = (uintptr_t)r[REG_RSP];
    g_callers[i] = (bp >= sp && bp - sp < (1u << 20) && (bp & 7) == 0) ? ((void**)bp)[1] : nullptr;   // same stack, plausible frame
}
void sample_dump() {
    const char* path = getenv("BPPP_SAMPLE");
    FILE* f = path ? fopen(path, "w") : nullptr;
    if (!f) return;
    struct itimerval off = {};
    setitimer(ITIMER_PROF, &off, nullptr);
    size_t n = std::min(g_ns.load(), kMaxSamples);
    for (size_t i = 0; i < n; i++) {
        Dl_info di;
        Dl_info dc;
        unsigned long coff = 0;
        if (g_callers[i] && dladdr(g_callers[i], &dc) && dc.dli_fname && strstr(dc.dli_fname, "libbppp_b200"))
            coff = (unsigned long)((char*)g_callers[i] - (char*)dc.dli_fbase);
        if (dladdr(g_samples[i], &di) && di.dli_fname)
            fprintf(f, "%s %lx@%lx %s\n", di.dli_fname, (unsigned long)((char*)g_samples[i] - (char*)di.dli_fbase), coff, di.dli_sname ? di.dli_sname : "?");
        else fprintf(f, "? %lx@0 ?\n", (unsigned long)g_samples[i]);
    }
    fclose(f);
}
void sample_start() {
    static std::once_flag once;
    std::call_once(once, [] {
        if (!getenv("BPPP_SAMPLE")) return;
        g_samples = (void**)calloc(kMaxSamples, sizeof(void*));
        g_callers = (void**)calloc(kMaxSamples, sizeof(void*));
        struct sigaction sa = {};
        sa.sa_sigaction = sample_handler;
        sa.sa_flags = SA_SIGINFO | SA_RESTART;
        sigaction(SIGPROF, &sa, nullptr);
        struct itimerval it = {{0, 1000}, {0, 1000}};
        setitimer(ITIMER_PROF, &it, nullptr);
        atexit(sample_dump);
    });
}
#else
void sample_start() {}
#endif
}  // namespace

// =================================================================================== C ABI
extern "C" {

void bppp_set_host_threads(int n) {
    g_threads = n;
    bppp_set_device_host_threads(n);
}

// Process-wide tuning for a batch-proving process, OPT-IN (nothing here happens unless the embedding
// application asks for it -- a library loaded under a GHC RTS or next to torch must not change global
// state behind its host's back):
//   BPPP_TUNE_MALLOC    keep freed memory in the malloc arenas (the per-proof scratch vectors of 16 host
//                       threads add up to ~1 MB per proof: no mmap/munmap + page faults per proof)
//   BPPP_TUNE_DEVICE    contexts created afterwards ask for cudaDeviceScheduleBlockingSync and pre-grow
//                       the stream-ordered memory pool (BPPP_POOL_PREWARM_MB, default 8192)
extern int g_bppp_tune_device;
int bppp_tune_process(int flags) {
    if (flags & BPPP_TUNE_MALLOC) {
        static std::once_flag once;
        std::call_once(once, [] {
            mallopt(M_MMAP_THRESHOLD, 1 << 30);
            mallopt(M_TRIM_THRESHOLD, 0x7fffffff);
            mallopt(M_TOP_PAD, 256 << 20);
        });
    }
    if (flags & BPPP_TUNE_DEVICE) g_bppp_tune_device = 1;
    return BPPP_OK;
}

int bppp_rp_setup(bppp_ctx* ctx, int binary, int arg_kind, int typed_or_conserved, const char* basis_seed, int show_format,
                  int root_policy, size_t n_ranges, const bppp_range_spec* ranges, size_t n_pub, const bppp_public_spec* pubs,
                  bppp_rp** out) {
    if (!ctx || !out || !basis_seed || (n_ranges && !ranges) || (n_pub && !pubs)) return BPPP_ERR_ARG;
    *out = nullptr;
    sample_start();
    if (!h64::host_cpu_ok()) {
        fprintf(stderr, "bppp_rp_setup: this build's host field arithmetic needs BMI2 + ADX (rebuild with -DBPPP_HOST_PORTABLE_FR)\n");
        return BPPP_ERR_STATE;
    }
    bppp_rp* s = new bppp_rp();
    s->ctx = ctx; s->binary = binary != 0; s->arg = arg_kind; s->flag = typed_or_conserved != 0;
    s->fmt = show_format; s->root = root_policy; s->basis_seed = basis_seed;
    s->n_inputs = n_ranges;
    for (size_t i = 0; i < n_ranges; i++) {
        Range rd;
        rd.mn = load_i128(ranges[i].min); rd.mx = load_i128(ranges[i].max); rd.base = ranges[i].base;
        rd.is_shared = ranges[i].is_shared; rd.is_output = ranges[i].is_output; rd.is_assumed = ranges[i].is_assumed;
        bool ok = s->binary ? make_range_bin(rd) : make_range_trrp(rd);
        if (!ok) { delete s; return BPPP_ERR_RANGE; }
        s->rds.push_back(rd);
    }
    for (size_t i = 0; i < n_pub; i++) {
        Public p;
        p.amount = load_i128(pubs[i].amount); p.type = load_i128(pubs[i].type); p.is_output = pubs[i].is_output;
        s->pubs.push_back(p);
        s->net_pub += p.is_output ? -p.amount : p.amount;                 // app/Main.hs:314
    }
    size_t need;
    if (s->binary) {                                                      // setupBRP (Binary.hs:143-156)
        s->num_rp_coms = 2;
        s->nrm_len = 0;
        for (auto& rd : s->rds) s->nrm_len += rd.coeffs.size();
        s->lin_len = 2;
        need = 4 + s->nrm_len;
    } else {                                                              // setup (TypedReciprocal.hs:332-359)
        s->num_rp_coms = 4;
        bool any_bit = false, any_shared_bit = false;
        std::set<U128> mb, sb;
        for (auto& rd : s->rds) {
            if (rd.is_assumed) continue;
            any_bit |= rd.has_bit;
            any_shared_bit |= (rd.has_bit && rd.is_shared);
            sb.insert(rd.base);
            if (rd.is_shared) mb.insert(rd.base);
        }
        if (any_shared_bit) mb.insert(2);
        if (any_bit) sb.insert(2);
        s->m_bases.assign(mb.begin(), mb.end());
        s->sorted_bases.assign(sb.begin(), sb.end());
        s->nrm_len = 0;
        for (auto& rd : s->rds) s->nrm_len += rd.coeffs.size() + (s->flag ? 1 : 0);
        s->lin_len = 6;
        for (auto b : s->m_bases) s->lin_len += (size_t)(b - 1);
        need = 2 + s->lin_len + s->nrm_len;
    }
    s->pts = tr::get_points(s->basis_seed, need, s->root);
    if (s->arg == BPPP_ARG_IP) optimal_witness_size_ip(s->nrm_len, s->lin_len, s->rounds, s->fin_n, s->fin_l);
    else optimal_witness_size_nl(s->nrm_len, s->lin_len, s->rounds, s->fin_n, s->fin_l);
    if (s->binary) {                       // the prover's own rule (Binary.hs:195)
        s->prover_rounds = (size_t)std::max(0, integer_log(2, s->nrm_len) - 1);
        if (s->arg == BPPP_ARG_IP) {
            size_t a = (s->nrm_len + 1) / 2, l2 = s->lin_len;
            for (size_t i = 0; i < s->prover_rounds; i++) { a = round_reduce(a); l2 = round_reduce(l2); }
            s->prover_fin_n = 2 * a; s->prover_fin_l = l2;
        } else
        lengths_after(s->nrm_len, s->lin_len, s->prover_rounds, s->prover_fin_n, s->prover_fin_l);
    } else {
        s->prover_rounds = s->rounds; s->prover_fin_n = s->fin_n; s->prover_fin_l = s->fin_l;
    }
    if (!s->binary)
        for (size_t i = 0; i < s->rds.size(); i++) {
            std::vector<Ph1> t;
            const Range& rd = s->rds[i];
            if (rd.is_shared && !rd.is_assumed)
                for (size_t k = 0; k < rd.coeffs.size(); k++) {
                    Ph1 p;
                    p.kind = 'S'; p.ind = (int)i; p.base = (rd.has_bit && k == 0) ? (U128)2 : rd.base;
                    p.b = rd.coeffs_fr[k]; p.d = p.m = p.s = h64::zero();
                    t.push_back(p);
                }
            s->ph1_tmpl.push_back(t);
        }
    finish_setup(s);
    int rc = bppp_fb_create(ctx, s->in_pts.size() / 64, s->in_pts.data(), &s->fb);
    if (rc) { delete s; return rc; }
    rc = bppp_gens_create(ctx, s->nrm_len, s->lin_len, &s->table[0], &s->table[64], &s->table[64 * (1 + s->nrm_len)], &s->gens);
    if (rc) { bppp_fb_destroy(s->fb); delete s; return rc; }
    {   // extra lanes (BPPP_LANES, default 8 in total)
        const char* ev = getenv("BPPP_LANES");
        int lanes = ev ? atoi(ev) : 8;
        int dev = bppp_ctx_device(ctx);
        for (int i = 1; i < lanes; i++) {
            bppp_ctx* c2 = nullptr;
            bppp_fb* f2 = nullptr;
            bppp_gens* g2 = nullptr;
            if (bppp_init(dev, &c2)) break;
            if (bppp_fb_create(c2, s->in_pts.size() / 64, s->in_pts.data(), &f2) ||
                bppp_gens_create(c2, s->nrm_len, s->lin_len, &s->table[0], &s->table[64], &s->table[64 * (1 + s->nrm_len)], &g2)) {
                bppp_fb_destroy(f2);
                bppp_free(c2);
                break;
            }
            s->lane_ctx.push_back(c2); s->lane_fb.push_back(f2); s->lane_gens.push_back(g2);
        }
    }
    // device scalar phases: TypedReciprocal proofs over the norm-linear argument (BPPP_RP_HOST_PHASES=1
    // keeps them on the host; binary proofs and the inner-product argument always run them there)
    if (!s->binary && s->arg == BPPP_ARG_NL && !(getenv("BPPP_RP_HOST_PHASES") && atoi(getenv("BPPP_RP_HOST_PHASES")))) {
        std::vector<Ph1> tp = ph1s_verifier(s);
        if (tp.size() == s->nrm_len) {
            const size_t ne = tp.size();
            std::vector<uint8_t> desc(16 * ne), eb(32 * ne), es(32 * ne);
            for (size_t i = 0; i < ne; i++) {
                const Ph1& p = tp[i];
                uint32_t flags = (p.kind == 'T' ? 1u : 0u) | (p.io ? 2u : 0u) | (p.ia ? 4u : 0u) | (p.kind == 'I' ? 8u : 0u) |
                                 (p.s_zero ? 0u : 16u);
                int32_t ind = p.ind, bi = -1, pad = 0;
                if (p.kind != 'T')
                    for (size_t j = 0; j < s->sorted_bases.size(); j++)
                        if (s->sorted_bases[j] == p.base) bi = (int32_t)j;
                memcpy(&desc[16 * i], &flags, 4); memcpy(&desc[16 * i + 4], &ind, 4);
                memcpy(&desc[16 * i + 8], &bi, 4); memcpy(&desc[16 * i + 12], &pad, 4);
                h64::to_bytes(&eb[32 * i], p.kind == 'T' ? h64::zero() : p.b);
                h64::to_bytes(&es[32 * i], p.s_zero ? h64::zero() : p.s);
            }
            int rc2 = bppp_trrp_create(s->gens, ne, desc.data(), eb.data(), es.data(), s->rds.size(), s->sorted_bases.size(), &s->trrp);
            for (size_t i = 0; !rc2 && i < s->lane_gens.size(); i++) {
                bppp_trrp* t = nullptr;
                rc2 = bppp_trrp_create(s->lane_gens[i], ne, desc.data(), eb.data(), es.data(), s->rds.size(), s->sorted_bases.size(), &t);
                if (!rc2) s->lane_trrp.push_back(t);
            }
            // makeSharedCoeffs' static part (TypedReciprocal.hs:204-206): slot -> (shared base, symbol)
            std::vector<int32_t> sh_b;
            std::vector<uint8_t> sh_s;
            for (auto b : s->m_bases) {
                int32_t bi = 0;
                for (size_t j = 0; j < s->sorted_bases.size(); j++)
                    if (s->sorted_bases[j] == b) bi = (int32_t)j;
                for (U128 sv = 1; sv < b; sv++) {
                    sh_b.push_back(bi);
                    sh_s.resize(sh_s.size() + 32, 0);
                    memcpy(&sh_s[sh_s.size() - 32], &sv, 16);
                }
            }
            s->n_shared = sh_b.size();
            if (!rc2 && s->n_shared) {
                rc2 = bppp_trrp_set_shared(s->trrp, s->n_shared, sh_b.data(), sh_s.data());
                for (size_t i = 0; !rc2 && i < s->lane_trrp.size(); i++) rc2 = bppp_trrp_set_shared(s->lane_trrp[i], s->n_shared, sh_b.data(), sh_s.data());
            }
            if (rc2) { bppp_rp_free(s); return rc2; }
            s->dev_phases = true;
            const char* dv = getenv("BPPP_DEVICE_TRANSCRIPT");
            s->dev_transcript = dv && atoi(dv);
        }
    }
    {
        const char* bv = getenv("BPPP_BATCH_VERIFY");
        s->batch_verify = bv && atoi(bv);
        const char* lg = getenv("BPPP_LUT_GB");
        if (lg && atof(lg) > 0) {
            int rc3 = bppp_rp_enable_lut(s, atof(lg), nullptr);
            if (rc3) { bppp_rp_free(s); return rc3; }
        }
    }
    *out = s;
    return BPPP_OK;
}
// Fiat-Shamir transcript of bppp_rp_prove_batch / bppp_rp_verify_batch on the device (on != 0) or on the host
// (default; app/Main.hs:75-80, src/ZKP.hs:96-101).  Proofs and verdicts are bit-identical either way.  Only setups
// that run the device scalar phases (TypedReciprocal over the norm-linear argument) can move it.
int bppp_rp_enable_lut(bppp_rp* s, double budget_gb, int* c_out) {
    if (!s) return BPPP_ERR_ARG;
    int c = 0;
    int rc = bppp_gens_enable_lut(s->gens, budget_gb, &c);
    for (size_t i = 0; !rc && i < s->lane_gens.size(); i++) rc = bppp_gens_enable_lut(s->lane_gens[i], budget_gb, nullptr);
    if (rc) return fail(s, rc, std::string("bppp_gens_enable_lut: ") + bppp_last_error(s->ctx));
    if (c_out) *c_out = c;
    return BPPP_OK;
}
int bppp_rp_set_batch_verify(bppp_rp* s, int on) {
    if (!s) return BPPP_ERR_ARG;
    s->batch_verify = on != 0;
    return BPPP_OK;
}
int bppp_rp_set_device_transcript(bppp_rp* s, int on) {
    if (!s) return BPPP_ERR_ARG;
    if (on && !s->dev_phases) return fail(s, BPPP_ERR_STATE, "device transcript needs the device scalar phases");
    s->dev_transcript = on != 0;
    return BPPP_OK;
}
void bppp_rp_free(bppp_rp* s) {
    if (!s) return;
    for (auto t : s->lane_vtr) bppp_dtr_destroy(t);
    bppp_trrp_destroy(s->trrp);                 // before the generator sets / contexts they refer to
    for (auto t : s->lane_trrp) bppp_trrp_destroy(t);
    bppp_fb_destroy(s->fb);
    bppp_gens_destroy(s->gens);
    for (auto& lane : s->pinned)
        for (auto& pb : lane) pb.release();
    for (size_t i = 0; i < s->lane_ctx.size(); i++) {
        bppp_fb_destroy(s->lane_fb[i]);
        bppp_gens_destroy(s->lane_gens[i]);
        bppp_free(s->lane_ctx[i]);
    }
    delete s;
}
const char* bppp_rp_last_error(bppp_rp* s) { return s ? s->err.c_str() : "null setup"; }

// infoRP (TypedReciprocal.hs:292, Binary.hs:120) + optimalWitnessSize: shape of a proof
int bppp_rp_info(bppp_rp* s, size_t* n_inputs, size_t* num_rp_coms, size_t* nrm_len, size_t* lin_len, size_t* rounds,
                 size_t* fin_norm, size_t* fin_lin) {
    if (!s) return BPPP_ERR_ARG;
    if (n_inputs) *n_inputs = s->n_inputs;
    if (num_rp_coms) *num_rp_coms = s->num_rp_coms;
    if (nrm_len) *nrm_len = s->nrm_len;
    if (lin_len) *lin_len = s->lin_len;
    if (rounds) *rounds = s->prover_rounds;
    if (fin_norm) *fin_norm = s->prover_fin_n;
    if (fin_lin) *fin_lin = s->prover_fin_l;
    return BPPP_OK;
}
// the generator list (h, g, ...) as the setup derived it: count*64 bytes
int bppp_rp_points(bppp_rp* s, size_t count, uint8_t* out) {
    if (!s || !out || count > s->pts.size()) return BPPP_ERR_ARG;
    memcpy(out, s->pts.data(), count * 64);
    return BPPP_OK;
}
// `hashToScalars ("Blinding " <> rn)` position j (1-based)  (app/Main.hs:86-87)
int bppp_input_blind(const char* random_seed, uint64_t j, uint8_t out[32]) {
    if (!h64::host_cpu_ok()) return BPPP_ERR_STATE;
    if (!random_seed || !out) return BPPP_ERR_ARG;
    h64::to_bytes(out, tr::input_blind(random_seed, j));
    return BPPP_OK;
}

// proveTRRPM phases 2-4 (TypedReciprocal.hs:412-444) with the norm-entry arithmetic on the device
// (bppp_trrp_*): the host keeps the transcript, the blinders, the scalar/linear slots and the few
// per-proof constants; reciprocals, error terms, public constants and the combined witness never
// leave the GPU.
static int prove_trrp_device(bppp_rp* s, const Lane& ln, std::vector<Proof>& P, uint8_t* coms, uint8_t* responses, uint8_t* finals,
                             const uint8_t* c1, const uint8_t* n_coms, const char* const* random_seeds, uint8_t* cht) {
    // cht != NULL: device transcript -- the challenges of the last oracle call, [batch][3]; chi: their inverses
    const bool dt = cht != nullptr;
    uint8_t* chi = dt ? lane_buf(s, ln, PB_CHI, P.size() * 3 * 32) : nullptr;
    const size_t B = P.size(), n = s->n_inputs, N = s->nrm_len, M = s->lin_len, NC = s->num_rp_coms + n;
    uint8_t* ch = lane_buf(s, ln, PB_CH, B * 4 * 32);
    uint8_t* sclin = lane_buf(s, ln, PB_SCLIN, B * (1 + M) * 32);
    uint8_t* bln = lane_buf(s, ln, PB_BLN, B * N * 32);
    uint8_t* small = lane_buf(s, ln, PB_SMALL, B * 6 * 32);
    uint8_t* c2 = lane_buf(s, ln, PB_C2, B * 64);
    uint8_t* q_b = lane_buf(s, ln, PB_Q, B * 32);
    uint8_t* sc_b = lane_buf(s, ln, PB_S, B * 32);
    uint8_t* l_b = lane_buf(s, ln, PB_L, B * M * 32);
    uint8_t* c_b = lane_buf(s, ln, PB_C, B * M * 32);
    auto put_sclin = [&](size_t b, const RPW& w) {
        uint8_t* o = sclin + 32 * b * (1 + M);
        memset(o, 0, 32 * (1 + M));
        h64::to_bytes(o, w.sc);
        for (size_t i = 0; i < w.lin.size() && i < M; i++)
            if (!w.lin[i].is_zero()) h64::to_bytes(o + 32 * (1 + i), w.lin[i]);
    };
    const size_t ERR7_SLOT = 4;                                            // blindErrWitness 3 [err7]: lin = [b1, b2, 0, b3, err7, 0]
    // ---------------- phase 2
    parallel_for(B, [&](size_t b) {
        Proof& p = P[b];
        uint8_t* out = coms + 64 * b * NC;
        memcpy(out + 128, &c1[64 * (2 * b)], 64);                           // dmCom
        memcpy(out + 192, &c1[64 * (2 * b + 1)], 64);                       // mCom
        memcpy(out + 256, &n_coms[64 * b * n], 64 * n);
        Fr chs[3];
        Sect sect;
        if (dt) for (int i = 0; i < 3; i++) chs[i] = h64::from_bytes(cht + 32 * (3 * b + i));
        else p.zk.oracle(out + 128, 2 + n, chs, 3);                         // T3 e x r0 <- oracle' (dmCom:mCom:nComs)
        sect.lap(S_ORACLE);
        p.e = chs[0]; p.x = chs[1]; p.r0 = chs[2];
        if (dt) {
            p.e_inv = h64::from_bytes(chi + 32 * (3 * b)); p.r0_inv = h64::from_bytes(chi + 32 * (3 * b + 2));
        } else {
            Fr iv[2] = {p.e, p.r0};
            h64::batch_inv(iv, 2);
            p.e_inv = iv[0]; p.r0_inv = iv[1];
        }
        p.base_map = make_base_map(s, p.x);
        p.dm.nrm.clear(); p.m.nrm.clear(); p.ph1s.clear();                  // the norm parts live on the device
        p.r = blind_err_witness(p.zk, 3, {h64::zero()}, {}, {});            // err7 is filled in by the device
        sect.lap(S_RANDOM);
        put_sclin(b, p.r);
        h64::to_bytes(ch + 32 * (4 * b), p.e); h64::to_bytes(ch + 32 * (4 * b + 1), p.e_inv);
        h64::to_bytes(ch + 32 * (4 * b + 2), p.x); h64::to_bytes(ch + 32 * (4 * b + 3), p.r0_inv);
        sect.lap(S_SCALARS);
    });
    g_tm.lap("host_phase2");
    if (dt) bppp_trrp_want_inverses(ln.trrp, chi);                             // 1/q, 1/x', 1/r1
    int rc = dt ? bppp_trrp_phase2_tr(ln.trrp, ch, sclin, ERR7_SLOT, c2, small, cht) : bppp_trrp_phase2(ln.trrp, ch, sclin, ERR7_SLOT, c2, small);
    if (rc) return fail(s, rc, std::string("reciprocal commitment: ") + ctx_err(ln));
    // the shared-multiplicity coefficients (255 reciprocals per proof for base 256) come from the device as well
    const uint8_t* shc = nullptr;
    if (dt && s->n_shared) {
        uint8_t* o = lane_buf(s, ln, PB_SHC, B * s->n_shared * 32);
        rc = bppp_trrp_shared_coeffs(ln.trrp, 1, o);
        if (rc) return fail(s, rc, std::string("shared coefficients: ") + ctx_err(ln));
        shc = o;
    }
    g_tm.lap("msm_phase2");
    // ---------------- phase 3
    bool dev_rnd = dt;
    for (size_t b = 0; dev_rnd && b < B; b++) dev_rnd = strlen(random_seeds[b]) <= 40;
    std::vector<uint64_t> rnd_n0(dev_rnd ? B : 0);
    parallel_for(B, [&](size_t b) {
        Proof& p = P[b];
        uint8_t* out = coms + 64 * b * NC;
        memcpy(out + 64, &c2[64 * b], 64);                                  // rCom
        p.r.lin[ERR7_SLOT] = h64::from_bytes(small + 32 * b);
        Fr chs[3];
        Sect sect;
        if (dt) for (int i = 0; i < 3; i++) chs[i] = h64::from_bytes(cht + 32 * (3 * b + i));
        else p.zk.oracle(out + 64, 1, chs, 3);                              // T3 q x' r1 <- oracle' [rCom]
        sect.lap(S_ORACLE);
        p.q = chs[0]; p.xq = chs[1]; p.r1 = chs[2];
        p.q0 = q0_of(s->arg, p.q);
        if (dt) {
            p.q_inv = h64::from_bytes(chi + 32 * (3 * b)); p.r1_inv = h64::from_bytes(chi + 32 * (3 * b + 2));
            p.q0_inv = q0_of(s->arg, p.q_inv);                              // 1 / (+-q^2) = +-(1/q)^2
        } else {
            Fr iv[3] = {p.q, p.q0, p.r1};
            h64::batch_inv(iv, 3);
            p.q_inv = iv[0]; p.q0_inv = iv[1]; p.r1_inv = iv[2];
        }
        std::vector<U128> mb;
        for (auto& kv : p.base_mss) mb.push_back(kv.first);
        if (shc && mb == s->m_bases) {                                      // Montgomery residues: the host's own form
            p.shared_cs.resize(s->n_shared);
            memcpy((void*)p.shared_cs.data(), shc + 32 * b * s->n_shared, 32 * s->n_shared);
        } else p.shared_cs = make_shared_coeffs(p.e, p.e_inv, mb, p.base_map);
        sect.lap(S_COEFFS);
        p.bls_lin.resize(M > 5 ? M - 5 : 0);
        p.zk.random_fill(p.bls_lin.data(), p.bls_lin.size());
        if (dev_rnd) { rnd_n0[b] = p.zk.n_random; p.zk.n_random += N; }      // the N norm blinders are drawn on the device
        else p.zk.random_fill_canonical(bln + 32 * b * N, N);
        sect.lap(S_RANDOM);
        h64::to_bytes(ch + 32 * (2 * b), p.q0); h64::to_bytes(ch + 32 * (2 * b + 1), p.xq);
        std::vector<Fr> ic = input_coeffs_trrp(s, p.x, p.q0);
        RPW nsum;                                                           // sum_i ic_i * nWit_i without temporaries
        size_t nl = 0;
        for (size_t i = 0; i < n; i++) nl = std::max(nl, p.n_wits[i].lin.size());
        nsum.lin.assign(nl, h64::zero());
        for (size_t i = 0; i < n; i++) {
            const RPW& w = p.n_wits[i];
            if (!w.sc.is_zero()) nsum.sc = h64::add(nsum.sc, h64::mul(w.sc, ic[i]));
            for (size_t j = 0; j < w.lin.size(); j++)
                if (!w.lin[j].is_zero()) nsum.lin[j] = h64::add(nsum.lin[j], h64::mul(w.lin[j], ic[i]));
        }
        p.wit = nsum;                                                       // parked: nWitSum
        sect.lap(S_COMBINE);
    });
    g_tm.lap("host_phase3");
    rc = dev_rnd ? bppp_trrp_phase3_rnd(ln.trrp, ch, random_seeds, rnd_n0.data(), small) : bppp_trrp_phase3(ln.trrp, ch, bln, small);
    if (rc) return fail(s, rc, std::string("error terms: ") + ctx_err(ln));
    g_tm.lap("msm_phase3");
    parallel_for(B, [&](size_t b) {
        Proof& p = P[b];
        Sect sect;
        std::vector<Fr> errs(6);
        for (int i = 0; i < 6; i++) errs[i] = h64::from_bytes(small + 32 * (6 * b + i));
        Fr aug = h64::zero();                                               // 2 sum cs_i bls_i over the shared multiplicities
        for (size_t i = 0; i < p.shared_cs.size() && i + 1 < p.bls_lin.size(); i++)
            aug = h64::add(aug, h64::mul(p.shared_cs[i], p.bls_lin[i + 1]));
        errs[3] = h64::add(errs[3], h64::dbl(aug));
        sect.lap(S_ERRTERMS);
        Fr tC = s->flag ? p.xq : h64::zero();
        Fr input_bl = p.wit.lin.size() > 1 ? p.wit.lin[1] : h64::zero();
        RPW blbl;
        blbl.lin = p.bls_lin;
        std::vector<const RPW*> wits = {&p.m, &p.dm, &p.r};
        p.bl = blind_blinding_term(blbl, tC, p.r0, p.r0_inv, p.r1, p.r1_inv, errs, wits, input_bl);
        sect.lap(S_BLIND);
        put_sclin(b, p.bl);
        sect.lap(S_SCALARS);
    });
    g_tm.lap("host_phase3");
    rc = dt ? bppp_trrp_commit_bl_tr(ln.trrp, sclin, c2, cht) : bppp_trrp_commit_bl(ln.trrp, sclin, c2);
    if (rc) return fail(s, rc, std::string("blinding commitment: ") + ctx_err(ln));
    g_tm.lap("msm_phase3");
    // ---------------- phase 4
    parallel_for(B, [&](size_t b) {
        Proof& p = P[b];
        uint8_t* out = coms + 64 * b * NC;
        memcpy(out, &c2[64 * b], 64);                                       // blCom
        Sect sect;
        if (dt) p.t = h64::from_bytes(cht + 32 * b);
        else p.zk.oracle(out, 1, &p.t, 1);
        sect.lap(S_ORACLE);
        h64::to_bytes(ch + 32 * (2 * b), p.t); h64::to_bytes(ch + 32 * (2 * b + 1), p.q0_inv);
    });
    g_tm.lap("host_phase4");
    rc = bppp_trrp_phase4(ln.trrp, ch, small);
    if (rc) return fail(s, rc, std::string("public constants: ") + ctx_err(ln));
    g_tm.lap("msm_phase4");
    parallel_for(B, [&](size_t b) {
        Proof& p = P[b];
        Sect sect;
        using namespace h64;
        const Fr ts0 = from_bytes(small + 32 * (3 * b)), sum_q2 = from_bytes(small + 32 * (3 * b + 1)), sum_v = from_bytes(small + 32 * (3 * b + 2));
        const Fr t2 = sqr(p.t), t3 = mul(t2, p.t), t5 = mul(sqr(t2), p.t), two_t5 = dbl(t5);
        RPW wit;                                                            // scalar + linear slots only
        wit.sc = add(public_consts_z_trrp(s, p.e, p.x, two_t5), add(ts0, mul(two_t5, add(sum_q2, mul(p.e_inv, sum_v)))));
        sect.lap(S_PUB);
        rpw_add(wit, p.bl);
        rpw_add(wit, rpw_scale(p.m, p.t));
        rpw_add(wit, rpw_scale(p.dm, t2));
        rpw_add(wit, rpw_scale(p.r, t3));
        rpw_add(wit, rpw_scale(p.wit, two_t5));
        p.cs = make_bp_coeffs(s->flag, p.xq, p.r0, p.r1, p.t, p.shared_cs);
        sect.lap(S_COMBINE);
        memset(&l_b[32 * b * M], 0, 32 * M);
        memset(&c_b[32 * b * M], 0, 32 * M);
        to_bytes(&q_b[32 * b], p.q);
        to_bytes(&sc_b[32 * b], wit.sc);
        for (size_t i = 0; i < wit.lin.size() && i < M; i++) to_bytes(&l_b[32 * (b * M + i)], wit.lin[i]);
        for (size_t i = 0; i < p.cs.size() && i < M; i++) to_bytes(&c_b[32 * (b * M + i)], p.cs[i]);
        sect.lap(S_TOBYTES);
    });
    g_tm.lap("host_phase4");
    return run_argument(s, ln, P, s->prover_rounds, q_b, sc_b, nullptr, l_b, c_b, responses, finals, s->prover_fin_n, s->prover_fin_l, true, dt);
}

// RangeProof.proveM (src/RangeProof.hs:95-97) for `batch` independent proofs.
//   values/types/blinds: [batch][n_inputs] 32-byte scalars (types ignored for binary proofs;
//   blinds == NULL derives them from the proof's randomSeed like app/Main.hs:275-276);
//   random_seeds: [batch] NUL-terminated randomSeed strings.
// Outputs: coms [batch][num_rp_coms + n_inputs] points in the reference's order
//   (blCom : rCom : dmCom : mCom : nComs  /  blCom : dCom : nComs); responses [batch][rounds][2]
//   points NEWEST FIRST; finals [batch][fin_norm + fin_lin] scalars (getWitness order).
static int prove_impl(bppp_rp* s, const Lane& ln, size_t batch, const uint8_t* values, const uint8_t* types, const uint8_t* blinds,
                      const char* const* random_seeds, uint8_t* coms, uint8_t* responses, uint8_t* finals) {
    const size_t B = batch, n = s->n_inputs, N = s->nrm_len, M = s->lin_len, P0 = 1 + N + M, NC = s->num_rp_coms + n;
    std::vector<Proof> P(B);
    std::atomic<int> bad(0);
    g_tm.start();
    const size_t in_terms = s->binary ? 2 : 3;
    const bool dev_mode = s->dev_phases && ln.trrp && !s->binary;
    uint8_t* in_sc = lane_buf(s, ln, PB_IN, B * n * in_terms * 32);
    // ---------------- phase 1 (host): witnesses, input openings, digit commitments
    uint8_t* sc1 = lane_buf(s, ln, PB_SC1, B * (s->binary ? 1 : 2) * P0 * 32);
    parallel_for(B, [&](size_t b) {
        Proof& p = P[b];
        Sect sect;
        p.zk.fmt = s->fmt;
        p.zk.seed = random_seeds[b];
        std::vector<Fr> vals(n), tys(n), bls(n);
        for (size_t i = 0; i < n; i++) {
            vals[i] = h64::from_bytes(values + 32 * (b * n + i));
            tys[i] = (types && !s->binary) ? h64::from_bytes(types + 32 * (b * n + i)) : h64::zero();
            if (blinds) bls[i] = h64::from_bytes(blinds + 32 * (b * n + i));
        }
        if (!blinds) {
            // hashToScalars ("Blinding " <> rn), positions 1.. (app/Main.hs:86-87, 275-276): one-block messages, hashed
            // two at a time when the seed is short enough
            const std::string pre = "Blinding " + p.zk.seed;
            size_t i = 0;
            if (pre.size() + 20 <= 55) {
                uint8_t m0[64], m1[64], d0[32], d1[32];
                memcpy(m0, pre.data(), pre.size());
                memcpy(m1, pre.data(), pre.size());
                for (; i + 1 < n; i += 2) {
                    const size_t l0 = pre.size() + tr::format_uint((char*)m0 + pre.size(), i + 1);
                    const size_t l1 = pre.size() + tr::format_uint((char*)m1 + pre.size(), i + 2);
                    sha::digest_short_x2(d0, m0, l0, d1, m1, l1);
                    uint64_t w[4];
                    tr::digest_to_words(w, d0);
                    bls[i] = h64::from_wide(w);
                    tr::digest_to_words(w, d1);
                    bls[i + 1] = h64::from_wide(w);
                }
            }
            for (; i < n; i++) bls[i] = tr::input_blind(p.zk.seed, i + 1);
        }
        if (s->binary) {
            // witnessBRP (Binary.hs:161-168): Nothing unless conserved and balanced
            Fr vsum = h64::from_i128(s->net_pub);
            for (size_t i = 0; i < n; i++) vsum = s->rds[i].is_output ? h64::sub(vsum, vals[i]) : h64::add(vsum, vals[i]);
            if (!s->flag || !vsum.is_zero()) { p.ok = false; bad++; return; }
            std::vector<Fr> ds;
            for (size_t i = 0; i < n; i++) {
                const Range& rd = s->rds[i];
                if (rd.is_assumed) continue;
                U128 n_adj;
                if (!adjust_value(rd, values + 32 * (b * n + i), n_adj)) { p.ok = false; bad++; return; }
                U128 bn = rd.coeffs[0];
                int n1 = (int)rd.coeffs.size() - 1;
                U128 dn = 0, n2 = n_adj;
                if (n_adj > bn) { dn = 1; n2 = n_adj - bn; }
                std::vector<Fr> bits;                                       // baseDigits 2, MSB first
                while (n2) { bits.insert(bits.begin(), h64::from_u64((uint64_t)(n2 & 1))); n2 >>= 1; }
                ds.push_back(h64::from_u64((uint64_t)dn));
                for (int k = (int)bits.size(); k < n1; k++) ds.push_back(h64::zero());   // padLeft
                ds.insert(ds.end(), bits.begin(), bits.end());
            }
            for (size_t i = 0; i < n; i++) {
                RPW w; w.sc = vals[i]; w.lin = {bls[i]};
                p.n_wits.push_back(w);
                h64::to_bytes(&in_sc[32 * ((b * n + i) * 2)], vals[i]);
                h64::to_bytes(&in_sc[32 * ((b * n + i) * 2 + 1)], bls[i]);
            }
            p.s_bl = p.zk.random();
            Fr l_bl0 = p.zk.random();
            p.d.sc = p.s_bl; p.d.lin = {l_bl0, h64::zero()}; p.d.nrm = ds;
            commit_scalars(s, p.d, &sc1[32 * b * P0]);
            return;
        }
        // witnessTRRP (TypedReciprocal.hs:373-388)
        if (s->flag) {
            std::vector<std::pair<Fr, Fr>> sums;                            // (type, net amount)
            auto addto = [&](const Fr& t, const Fr& v, bool negate) {
                for (auto& kv : sums)
                    if (kv.first == t) { kv.second = negate ? h64::sub(kv.second, v) : h64::add(kv.second, v); return; }
                sums.push_back({t, negate ? h64::neg(v) : v});
            };
            for (auto& pb : s->pubs) addto(h64::from_i128(pb.type), h64::from_i128(pb.amount), pb.is_output);
            for (size_t i = 0; i < n; i++) addto(tys[i], vals[i], s->rds[i].is_output);
            for (auto& kv : sums)
                if (!kv.second.is_zero()) { p.ok = false; bad++; return; }
        }
        std::vector<Ph1> digits_ph1;
        std::map<U128, std::vector<U128>> bm;                              // multiplicities summed as integers
        if (dev_mode) {
            // The norm part of the witness goes to the device: digits (small integers), types and inline
            // multiplicities are written straight into the commitment scalar rows [sc | nrm | lin] of the
            // dm and m witnesses -- no Phase-1 records, no field conversions for the digits.
            uint8_t* dm_row = &sc1[32 * (2 * b) * P0];
            uint8_t* m_row = &sc1[32 * (2 * b + 1) * P0];
            memset(dm_row, 0, 32 * P0);
            memset(m_row, 0, 32 * P0);
            size_t ent = 0;
            auto put_int = [](uint8_t* dst, U128 v) { memcpy(dst, &v, 16); };
            if (s->flag)
                for (size_t i = 0; i < n; i++, ent++)
                    if (types) memcpy(dm_row + 32 * (1 + ent), types + 32 * (b * n + i), 32);
            for (size_t i = 0; i < n; i++) {
                const Range& rdi = s->rds[i];
                if (rdi.is_shared && !rdi.is_assumed) {
                    U128 left;
                    if (!adjust_value(rdi, values + 32 * (b * n + i), left)) { p.ok = false; bad++; return; }
                    std::vector<U128>& arr = bm[rdi.base];
                    if (arr.empty()) arr.assign((size_t)rdi.base - 1, 0);
                    for (size_t k = 0; k < rdi.coeffs.size(); k++, ent++) {
                        const bool bit = rdi.has_bit && k == 0;
                        const U128 basek = bit ? (U128)2 : rdi.base, cf = rdi.coeffs[k];
                        U128 d;
                        if (!cf) d = basek - 1;
                        else if (!(uint64_t)(left >> 64) && !(uint64_t)(cf >> 64)) {      // 64-bit operands: one hardware division
                            const uint64_t qd = (uint64_t)left / (uint64_t)cf;
                            d = std::min<U128>(basek - 1, qd);
                        } else d = std::min<U128>(basek - 1, left / cf);
                        left -= d * cf;
                        put_int(dm_row + 32 * (1 + ent), d);
                        if (bit) {
                            std::vector<U128>& a2 = bm[2];
                            if (a2.empty()) a2.assign(1, 0);
                            a2[0] += d;
                        } else if (d >= 1 && d < rdi.base) arr[(size_t)d - 1] += 1;
                    }
                    continue;
                }
                bool has_ms;
                std::vector<U128> ms;
                std::vector<Ph1> tmp;
                if (!make_phase1s((int)i, rdi, values + 32 * (b * n + i), true, tmp, has_ms, ms)) { p.ok = false; bad++; return; }
                for (auto& q : tmp) {
                    h64::to_bytes(dm_row + 32 * (1 + ent), q.d);
                    if (q.kind == 'I') h64::to_bytes(m_row + 32 * (1 + ent), q.m);
                    ent++;
                }
                if (has_ms) {                                               // baseMss (:363-367)
                    auto merge = [&](U128 base, const U128* v, size_t cnt) {
                        auto it = bm.find(base);
                        if (it == bm.end()) bm[base] = std::vector<U128>(v, v + cnt);
                        else for (size_t k = 0; k < cnt && k < it->second.size(); k++) it->second[k] += v[k];
                    };
                    if (rdi.has_bit) {
                        merge(2, ms.data(), 1);
                        merge(rdi.base, ms.data() + 1, ms.size() - 1);
                    } else merge(rdi.base, ms.data(), ms.size());
                }
            }
            if (ent != N) { p.ok = false; bad++; return; }
            sect.lap(S_WITNESS);
            std::vector<Fr> ms_shared;
            for (auto& kv : bm) {
                std::vector<Fr> v;
                for (auto m : kv.second) v.push_back(fr_small(m));
                ms_shared.insert(ms_shared.end(), v.begin(), v.end());
                p.base_mss.push_back({kv.first, v});
            }
            for (size_t i = 0; i < n; i++) {
                RPW w; w.sc = vals[i]; w.lin = {tys[i], bls[i]};
                p.n_wits.push_back(w);
                memcpy(&in_sc[32 * ((b * n + i) * 3)], values + 32 * (b * n + i), 32);       // already canonical
                h64::to_bytes(&in_sc[32 * ((b * n + i) * 3 + 1)], tys[i]);
                h64::to_bytes(&in_sc[32 * ((b * n + i) * 3 + 2)], bls[i]);
            }
            sect.lap(S_TOBYTES);
            p.dm = blind_witness(p.zk, 3, 2, ms_shared, {});                // scalar + linear slots (same blinder draws)
            p.m = blind_witness(p.zk, 3, 1, {}, {});
            sect.lap(S_RANDOM);
            auto put_sclin = [&](uint8_t* row, const RPW& w) {
                h64::to_bytes(row, w.sc);
                for (size_t j = 0; j < w.lin.size() && j < M; j++)
                    if (!w.lin[j].is_zero()) h64::to_bytes(row + 32 * (1 + N + j), w.lin[j]);
            };
            put_sclin(dm_row, p.dm);
            put_sclin(m_row, p.m);
            sect.lap(S_SCALARS);
            return;
        }
        digits_ph1.reserve(s->nrm_len);
        for (size_t i = 0; i < n; i++) {
            const Range& rdi = s->rds[i];
            if (rdi.is_shared && !rdi.is_assumed) {
                // fast path of makePhase1s for shared digits: template entries + digits, multiplicities
                // counted straight into the per-base integer arrays (baseMss, :363-367)
                U128 left;
                if (!adjust_value(rdi, values + 32 * (b * n + i), left)) { p.ok = false; bad++; return; }
                const size_t off = digits_ph1.size();
                digits_ph1.insert(digits_ph1.end(), s->ph1_tmpl[i].begin(), s->ph1_tmpl[i].end());
                std::vector<U128>& arr = bm[rdi.base];
                if (arr.empty()) arr.assign((size_t)rdi.base - 1, 0);
                for (size_t k = 0; k < rdi.coeffs.size(); k++) {
                    const bool bit = rdi.has_bit && k == 0;
                    const U128 basek = bit ? (U128)2 : rdi.base, cf = rdi.coeffs[k];
                    U128 d = cf ? std::min<U128>(basek - 1, left / cf) : (basek - 1);
                    left -= d * cf;
                    Ph1& e1 = digits_ph1[off + k];
                    e1.d = fr_small(d);
                    e1.di = d < 2048 ? (int)d : -1;
                    if (bit) {
                        std::vector<U128>& a2 = bm[2];
                        if (a2.empty()) a2.assign(1, 0);
                        a2[0] += d;
                    } else if (d >= 1 && d < rdi.base) arr[(size_t)d - 1] += 1;
                }
                continue;
            }
            bool has_ms;
            std::vector<U128> ms;
            if (!make_phase1s((int)i, s->rds[i], values + 32 * (b * n + i), true, digits_ph1, has_ms, ms)) { p.ok = false; bad++; return; }
            if (has_ms) {                                                   // baseMss (:363-367)
                const Range& rd = s->rds[i];
                auto merge = [&](U128 base, const U128* v, size_t cnt) {
                    auto it = bm.find(base);
                    if (it == bm.end()) bm[base] = std::vector<U128>(v, v + cnt);
                    else for (size_t k = 0; k < cnt && k < it->second.size(); k++) it->second[k] += v[k];
                };
                if (rd.has_bit) {
                    merge(2, ms.data(), 1);
                    merge(rd.base, ms.data() + 1, ms.size() - 1);
                } else merge(rd.base, ms.data(), ms.size());
            }
        }
        if (s->flag)
            for (size_t i = 0; i < n; i++) {
                Ph1 t;
                t.kind = 'T'; t.ind = (int)i; t.io = s->rds[i].is_output; t.ia = s->rds[i].is_assumed;
                t.m = vals[i]; t.d = tys[i];
                p.ph1s.push_back(t);
            }
        p.ph1s.insert(p.ph1s.end(), digits_ph1.begin(), digits_ph1.end());
        sect.lap(S_WITNESS);
        for (auto& kv : bm) {
            std::vector<Fr> v;
            for (auto m : kv.second) v.push_back(fr_small(m));
            p.base_mss.push_back({kv.first, v});
        }
        // proveTRRPM phase 1 (TypedReciprocal.hs:399-410)
        std::vector<Fr> ms_shared, ds, ms_inline;
        for (auto& kv : p.base_mss) ms_shared.insert(ms_shared.end(), kv.second.begin(), kv.second.end());
        for (auto& q : p.ph1s) { ds.push_back(q.d); ms_inline.push_back(q.kind == 'I' ? q.m : h64::zero()); }
        for (size_t i = 0; i < n; i++) {
            RPW w; w.sc = vals[i]; w.lin = {tys[i], bls[i]};
            p.n_wits.push_back(w);
            h64::to_bytes(&in_sc[32 * ((b * n + i) * 3)], vals[i]);
            h64::to_bytes(&in_sc[32 * ((b * n + i) * 3 + 1)], tys[i]);
            h64::to_bytes(&in_sc[32 * ((b * n + i) * 3 + 2)], bls[i]);
        }
        sect.lap(S_TOBYTES);
        p.dm = blind_witness(p.zk, 3, 2, ms_shared, ds);
        p.m = blind_witness(p.zk, 3, 1, {}, ms_inline);
        sect.lap(S_RANDOM);
        commit_scalars(s, p.dm, &sc1[32 * (2 * b) * P0]);
        commit_scalars(s, p.m, &sc1[32 * (2 * b + 1) * P0]);
        sect.lap(S_SCALARS);
    });
    if (bad.load()) return fail(s, BPPP_ERR_RANGE, "invalid witness (out of range / unbalanced)");
    g_tm.lap("host_phase1");
    // ---------------- device: input commitments + digit commitments
    uint8_t* n_coms = lane_buf(s, ln, PB_NCOMS, B * n * 64);
    uint8_t* c1 = lane_buf(s, ln, PB_C1, B * (s->binary ? 1 : 2) * 64);
    int rc;
    if (n) {
        rc = bppp_fb_msm_batch(ln.fb, B * n, in_sc, n_coms);
        if (rc) return fail(s, rc, std::string("input commitments: ") + ctx_err(ln));
    }
    const bool dev = s->dev_phases && ln.trrp;
    const bool dt = dev && s->dev_transcript;
    uint8_t* cht = dt ? lane_buf(s, ln, PB_CHT, B * 3 * 32) : nullptr;
    if (dt) {
        bppp_trrp_set_transcript(ln.trrp, s->fmt);
        bppp_trrp_want_inverses(ln.trrp, lane_buf(s, ln, PB_CHI, B * 3 * 32));      // 1/e, 1/x, 1/r0 come back inverted
    } else if (dev) bppp_trrp_set_transcript(ln.trrp, -1);
    rc = dt ? bppp_trrp_phase1_tr(ln.trrp, B, sc1, values, n, n_coms, c1, cht)
            : dev ? bppp_trrp_phase1(ln.trrp, B, sc1, values, c1) : bppp_gens_msm_batch(ln.gens, B * (s->binary ? 1 : 2), P0, sc1, c1);
    if (rc) return fail(s, rc, std::string("digit commitments: ") + ctx_err(ln));
    g_tm.lap("msm_phase1");
    if (dev) return prove_trrp_device(s, ln, P, coms, responses, finals, c1, n_coms, random_seeds, cht);
    uint8_t* q_b = lane_buf(s, ln, PB_Q, B * 32);
    uint8_t* sc_b = lane_buf(s, ln, PB_S, B * 32);
    uint8_t* w_b = lane_buf(s, ln, PB_W, B * N * 32);
    uint8_t* l_b = lane_buf(s, ln, PB_L, B * M * 32);
    uint8_t* c_b = lane_buf(s, ln, PB_C, B * M * 32);
    uint8_t* sc2 = lane_buf(s, ln, PB_SC2, B * P0 * 32);
    uint8_t* c2 = lane_buf(s, ln, PB_C2, B * 64);

    if (s->binary) {
        // ---------------- proveBRPM (Binary.hs:171-203)
        parallel_for(B, [&](size_t b) {
            Proof& p = P[b];
            uint8_t* out = coms + 64 * b * NC;
            memcpy(out + 64, &c1[64 * b], 64);                              // dCom
            memcpy(out + 128, &n_coms[64 * b * n], 64 * n);
            Fr ch[3];
            p.zk.oracle(out + 64, 1 + n, ch, 3);                            // T3 q x r <- oracle' (dCom:nComs)
            p.q = ch[0]; p.x = ch[1]; p.r0 = ch[2];
            Fr r_inv = h64::inv(p.r0);
            p.q0 = q0_of(s->arg, p.q);                                      // head (qPowers q)
            p.q0_inv = h64::inv(p.q0);
            p.pub = public_consts_brp(s, p.x, p.q0, p.q0_inv);
            p.bls_nrm.clear();
            { size_t n0 = p.bls_nrm.size(); p.bls_nrm.resize(n0 + N); p.zk.random_fill(p.bls_nrm.data() + n0, N); }
            Fr bl_bl = p.zk.random();
            // makePolyTerms (qPowers q) [blsNrm, nrm (dWit + pubWit)]  (Binary.hs:188, Internal.hs:65-75)
            std::vector<Fr> dn = p.d.nrm;
            vadd(dn, p.pub.nrm);
            std::vector<Fr> ws = q_powers(s->arg, p.q, N);
            Fr bl0 = h64::zero(), bl1 = h64::zero();
            for (size_t i = 0; i < N; i++) bl0 = h64::add(bl0, h64::mul(ws[i], h64::sqr(p.bls_nrm[i])));
            for (size_t i = 0; i < N && i < dn.size(); i++) bl1 = h64::add(bl1, h64::mul(ws[i], h64::mul(p.bls_nrm[i], dn[i])));
            bl1 = h64::dbl(bl1);
            p.bl.sc = bl0;
            p.bl.lin = {bl_bl, h64::mul(r_inv, h64::sub(p.s_bl, bl1))};
            p.bl.nrm = p.bls_nrm;
            commit_scalars(s, p.bl, &sc2[32 * b * P0]);
        });
        rc = bppp_gens_msm_batch(ln.gens, B, P0, sc2, c2);
        if (rc) return fail(s, rc, std::string("blinding commitment: ") + ctx_err(ln));
        parallel_for(B, [&](size_t b) {
            Proof& p = P[b];
            uint8_t* out = coms + 64 * b * NC;
            memcpy(out, &c2[64 * b], 64);                                   // blCom
            p.zk.oracle(out, 1, &p.t, 1);
            // wit' = pub' + dWit + 2t * sum coeff_i * nWit_i ; bpWit = blWit + t * wit'
            RPW pub1 = p.pub;
            pub1.sc = h64::mul(p.t, p.pub.sc);
            RPW w1 = pub1;
            rpw_add(w1, p.d);
            std::vector<Fr> ic = input_coeffs_brp(s, p.x);
            RPW nsum;
            for (size_t i = 0; i < n; i++) rpw_add(nsum, rpw_scale(p.n_wits[i], ic[i]));
            rpw_add(w1, rpw_scale(nsum, h64::dbl(p.t)));
            RPW wit = p.bl;
            rpw_add(wit, rpw_scale(w1, p.t));
            memset(&w_b[32 * b * N], 0, 32 * N);
            memset(&l_b[32 * b * M], 0, 32 * M);
            memset(&c_b[32 * b * M], 0, 32 * M);
            h64::to_bytes(&q_b[32 * b], p.q);
            h64::to_bytes(&sc_b[32 * b], wit.sc);
            for (size_t i = 0; i < wit.nrm.size() && i < N; i++) h64::to_bytes(&w_b[32 * (b * N + i)], wit.nrm[i]);
            for (size_t i = 0; i < wit.lin.size() && i < M; i++) h64::to_bytes(&l_b[32 * (b * M + i)], wit.lin[i]);
            h64::to_bytes(&c_b[32 * (b * M + 1)], h64::mul(p.r0, p.t));     // cs' = [0, r*t]  (Binary.hs:151)
        });
    } else {
        // ---------------- proveTRRPM phase 2 (TypedReciprocal.hs:412-419)
        parallel_for(B, [&](size_t b) {
            Proof& p = P[b];
            uint8_t* out = coms + 64 * b * NC;
            memcpy(out + 128, &c1[64 * (2 * b)], 64);                       // dmCom
            memcpy(out + 192, &c1[64 * (2 * b + 1)], 64);                   // mCom
            memcpy(out + 256, &n_coms[64 * b * n], 64 * n);
            Fr ch[3];
            Sect sect;
            p.zk.oracle(out + 128, 2 + n, ch, 3);                           // T3 e x r0 <- oracle' (dmCom:mCom:nComs)
            sect.lap(S_ORACLE);
            p.e = ch[0]; p.x = ch[1]; p.r0 = ch[2];
            Fr iv[2] = {p.e, p.r0};
            h64::batch_inv(iv, 2);
            p.e_inv = iv[0]; p.r0_inv = iv[1];
            p.base_map = make_base_map(s, p.x);
            p.ph2s = make_phase2s(true, p.e, p.e_inv, p.x, p.base_map, p.ph1s);
            Fr e7 = h64::zero();
            std::vector<Fr> rs;
            for (auto& o : p.ph2s) { e7 = h64::add(e7, h64::dbl(h64::mul(o.r, o.c))); rs.push_back(o.r); }
            Fr err7 = h64::mul(p.r0_inv, h64::neg(e7));
            sect.lap(S_PHASE2);
            p.r = blind_err_witness(p.zk, 3, {err7}, {}, rs);
            sect.lap(S_RANDOM);
            commit_scalars(s, p.r, &sc2[32 * b * P0]);
            sect.lap(S_SCALARS);
        });
        g_tm.lap("host_phase2");
        rc = bppp_gens_msm_batch(ln.gens, B, P0, sc2, c2);
        if (rc) return fail(s, rc, std::string("reciprocal commitment: ") + ctx_err(ln));
        g_tm.lap("msm_phase2");
        // ---------------- phase 3 (TypedReciprocal.hs:421-434)
        parallel_for(B, [&](size_t b) {
            Proof& p = P[b];
            uint8_t* out = coms + 64 * b * NC;
            memcpy(out + 64, &c2[64 * b], 64);                              // rCom
            Fr ch[3];
            Sect sect;
            p.zk.oracle(out + 64, 1, ch, 3);                                // T3 q x' r1 <- oracle' [rCom]
            sect.lap(S_ORACLE);
            p.q = ch[0]; p.xq = ch[1]; p.r1 = ch[2];
            p.q0 = q0_of(s->arg, p.q);
            Fr iv[3] = {p.q, p.q0, p.r1};
            h64::batch_inv(iv, 3);
            p.q_inv = iv[0]; p.q0_inv = iv[1]; p.r1_inv = iv[2];
            std::vector<U128> mb;
            for (auto& kv : p.base_mss) mb.push_back(kv.first);
            p.shared_cs = make_shared_coeffs(p.e, p.e_inv, mb, p.base_map);
            Fr tC = s->flag ? p.xq : h64::zero();
            sect.lap(S_COEFFS);
            p.bls_lin.clear(); p.bls_nrm.clear();
            p.bls_lin.resize(M > 5 ? M - 5 : 0);
            p.bls_nrm.resize(N);
            p.zk.random_fill(p.bls_lin.data(), p.bls_lin.size());
            p.zk.random_fill(p.bls_nrm.data(), N);
            sect.lap(S_RANDOM);
            std::vector<Fr> bls_ms(p.bls_lin.begin() + 1, p.bls_lin.end());
            std::vector<Fr> ic = input_coeffs_trrp(s, p.x, p.q0);
            RPW nsum;
            for (size_t i = 0; i < n; i++) rpw_add(nsum, rpw_scale(p.n_wits[i], ic[i]));
            Fr input_bl = nsum.lin.size() > 1 ? nsum.lin[1] : h64::zero();
            std::vector<Fr> q2s = q_powers(s->arg, p.q, p.ph2s.size());
            sect.lap(S_COMBINE);
            std::vector<Fr> errs = make_error_terms(p.e, p.xq, p.shared_cs, bls_ms, p.ph2s, q2s, p.bls_nrm);
            sect.lap(S_ERRTERMS);
            RPW blbl;
            blbl.lin = p.bls_lin; blbl.nrm = p.bls_nrm;
            std::vector<const RPW*> wits = {&p.m, &p.dm, &p.r};
            p.bl = blind_blinding_term(blbl, tC, p.r0, p.r0_inv, p.r1, p.r1_inv, errs, wits, input_bl);
            p.wit = nsum;                                                    // parked: nWitSum
            sect.lap(S_BLIND);
            commit_scalars(s, p.bl, &sc2[32 * b * P0]);
            sect.lap(S_SCALARS);
        });
        g_tm.lap("host_phase3");
        rc = bppp_gens_msm_batch(ln.gens, B, P0, sc2, c2);
        if (rc) return fail(s, rc, std::string("blinding commitment: ") + ctx_err(ln));
        g_tm.lap("msm_phase3");
        // ---------------- phase 4 (TypedReciprocal.hs:435-444)
        parallel_for(B, [&](size_t b) {
            Proof& p = P[b];
            uint8_t* out = coms + 64 * b * NC;
            memcpy(out, &c2[64 * b], 64);                                   // blCom
            Sect sect;
            p.zk.oracle(out, 1, &p.t, 1);
            sect.lap(S_ORACLE);
            p.pub = make_public_consts_trrp(s, p.e, p.e_inv, p.x, p.xq, p.q0, p.q0_inv, p.t, p.ph2s);
            sect.lap(S_PUB);
            Fr t2 = h64::sqr(p.t), t3 = h64::mul(t2, p.t), t5 = h64::mul(h64::sqr(t2), p.t);
            RPW nsum = p.wit;
            RPW wit = p.pub;
            rpw_add(wit, p.bl);
            rpw_add(wit, rpw_scale(p.m, p.t));
            rpw_add(wit, rpw_scale(p.dm, t2));
            rpw_add(wit, rpw_scale(p.r, t3));
            rpw_add(wit, rpw_scale(nsum, h64::dbl(t5)));
            p.cs = make_bp_coeffs(s->flag, p.xq, p.r0, p.r1, p.t, p.shared_cs);
            sect.lap(S_COMBINE);
            memset(&w_b[32 * b * N], 0, 32 * N);
            memset(&l_b[32 * b * M], 0, 32 * M);
            memset(&c_b[32 * b * M], 0, 32 * M);
            h64::to_bytes(&q_b[32 * b], p.q);
            h64::to_bytes(&sc_b[32 * b], wit.sc);
            for (size_t i = 0; i < wit.nrm.size() && i < N; i++) h64::to_bytes(&w_b[32 * (b * N + i)], wit.nrm[i]);
            for (size_t i = 0; i < wit.lin.size() && i < M; i++) h64::to_bytes(&l_b[32 * (b * M + i)], wit.lin[i]);
            for (size_t i = 0; i < p.cs.size() && i < M; i++) h64::to_bytes(&c_b[32 * (b * M + i)], p.cs[i]);
            sect.lap(S_TOBYTES);
        });
    }
    g_tm.lap("host_phase4");
    return run_argument(s, ln, P, s->prover_rounds, q_b, sc_b, w_b, l_b, c_b, responses, finals, s->prover_fin_n, s->prover_fin_l);
}

// RangeProof.verifyM (src/RangeProof.hs:99-101) for `batch` proofs; `rounds`, n_norm, n_lin
// describe the proofs as encoded (for binary proofs the prover's round rule may differ from
// optimalWitnessSize, src/RangeProof/Binary.hs:195 vs :218; verifyBPM ignores `rounds`).
// weights of a random linear combination over `n` proofs: 128 random bits each from the OS entropy source, drawn
// after the proofs are fixed (they are this call's inputs); canonical 32-byte scalars, the first one is 1
static std::vector<uint8_t> batch_weights(size_t n) {
    std::vector<uint8_t> w(32 * n, 0);
    std::vector<uint8_t> rnd(16 * n);
    size_t got = 0;
    while (got < rnd.size()) {
        ssize_t r = getrandom(rnd.data() + got, std::min<size_t>(rnd.size() - got, 256), 0);
        if (r <= 0) return std::vector<uint8_t>();                  // no entropy source: never predictable weights -- the caller verifies per proof
        got += (size_t)r;
    }
    for (size_t i = 0; i < n; i++) memcpy(&w[32 * i], &rnd[16 * i], 16);
    memset(&w[0], 0, 32);
    w[0] = 1;
    return w;
}
static int verify_impl(bppp_rp* s, const Lane& ln, size_t batch, size_t rounds, size_t n_norm, size_t n_lin, const uint8_t* coms,
                       const uint8_t* responses, const uint8_t* finals, int* ok) {
    const size_t B = batch, n = s->n_inputs, N = s->nrm_len, M = s->lin_len, NC = s->num_rp_coms + n, k = rounds;
    uint8_t* q_b = lane_buf(s, ln, PB_V0, B * 32);
    uint8_t* sp_b = lane_buf(s, ln, PB_V1, B * 32);
    uint8_t* pw_b = lane_buf(s, ln, PB_V2, B * N * 32);
    uint8_t* c_b = lane_buf(s, ln, PB_V3, B * M * 32);
    uint8_t* es_b = lane_buf(s, ln, PB_V4, B * k * 32);
    uint8_t* fw_b = lane_buf(s, ln, PB_V5, B * n_norm * 32);
    uint8_t* fl_b = lane_buf(s, ln, PB_V6, B * n_lin * 32);
    uint8_t* is_b = lane_buf(s, ln, PB_V7, B * NC * 32);
    uint8_t* ip_b = lane_buf(s, ln, PB_V8, B * NC * 64);
    g_tm.start();
    if (s->dev_phases && ln.trrp) {
        // verifyTRRPM (TypedReciprocal.hs:447-467) with the norm part of the public constants on the device
        struct VP { Fr q, e, e_inv, x, xq, r0, r1, q0, q0_inv, t; };
        std::vector<VP> V(B);
        uint8_t* ch = lane_buf(s, ln, PB_CH, B * 8 * 32);
        uint8_t* small = lane_buf(s, ln, PB_SMALL, B * 6 * 32);
        // device transcript: every commitment of the proofs is rendered once, then ONE launch squeezes the
        // challenges of all stages (each stage's transcript is a suffix of the final one)
        const uint8_t* dch = nullptr;
        const uint8_t* dchi = nullptr;
        const size_t n_ch = 7 + k;
        if (s->dev_transcript && k + 3 <= 64) {
            if ((size_t)ln.index >= s->lane_vtr.size()) return fail(s, BPPP_ERR_STATE, "lane without a transcript slot");
            bppp_dtr*& t = s->lane_vtr[ln.index];
            if (t && !bppp_dtr_fits(t, B, NC + 2 * k, s->fmt)) { bppp_dtr_destroy(t); t = nullptr; }
            int rc = t ? BPPP_OK : bppp_dtr_create(ln.ctx, B, NC + 2 * k, s->fmt, &t);
            if (!rc) rc = bppp_dtr_reset(t);
            if (!rc) rc = bppp_dtr_absorb(t, coms + 128, NC, 2 + n);        // dmCom : mCom : nComs
            if (!rc) rc = bppp_dtr_absorb(t, coms + 64, NC, 1);              // rCom
            if (!rc) rc = bppp_dtr_absorb(t, coms, NC, 1);                   // blCom
            for (size_t r = 0; !rc && r < k; r++) rc = bppp_dtr_absorb(t, responses + 128 * (k - 1 - r), 2 * k, 2);   // oldest round first
            std::vector<uint8_t> idx(n_ch, 1), st(n_ch);
            idx[1] = 2; idx[2] = 3; idx[4] = 2; idx[5] = 3;
            for (size_t j = 0; j < n_ch; j++) st[j] = (uint8_t)(j < 3 ? 1 : j < 6 ? 2 : j == 6 ? 3 : 4 + (j - 7));
            uint8_t* out = lane_buf(s, ln, PB_CHT, B * n_ch * 32);
            uint8_t* outi = lane_buf(s, ln, PB_CHI, B * n_ch * 32);
            if (!rc) rc = bppp_dtr_squeeze_inv(t, n_ch, idx.data(), st.data(), out, outi);
            if (rc) return fail(s, rc, std::string("device transcript: ") + ctx_err(ln));
            dch = out;
            dchi = outi;
            g_tm.lap("verify_transcript");
        }
        parallel_for(B, [&](size_t b) {
            tr::Zkpt zk;
            Sect sect;
            zk.fmt = s->fmt;
            zk.no_random = true;
            const uint8_t* cm = coms + 64 * b * NC;
            VP& v = V[b];
            if (dch) {
                const uint8_t* c = dch + 32 * b * n_ch;
                v.e = h64::from_bytes(c); v.x = h64::from_bytes(c + 32); v.r0 = h64::from_bytes(c + 64);
                v.q = h64::from_bytes(c + 96); v.xq = h64::from_bytes(c + 128); v.r1 = h64::from_bytes(c + 160);
                v.q0 = q0_of(s->arg, v.q);
                v.t = h64::from_bytes(c + 192);
                for (size_t r = 0; r < k; r++) memcpy(&es_b[32 * (b * k + (k - 1 - r))], c + 32 * (7 + r), 32);
            } else {
            Fr c1[3], c2[3];
            zk.oracle(cm + 128, 2 + n, c1, 3);
            v.e = c1[0]; v.x = c1[1]; v.r0 = c1[2];
            zk.oracle(cm + 64, 1, c2, 3);
            v.q = c2[0]; v.xq = c2[1]; v.r1 = c2[2];
            v.q0 = q0_of(s->arg, v.q);
            zk.oracle(cm, 1, &v.t, 1);
            // challenges of the argument: oldest round hashed first, list newest first (Bulletproof.hs:374)
            std::vector<const uint8_t*> rp(k);
            std::vector<Fr> es(k);
            for (size_t r = 0; r < k; r++) rp[r] = responses + 128 * (b * k + (k - 1 - r));
            zk.oracle_rounds(rp.data(), k, es.data());
            for (size_t r = 0; r < k; r++) h64::to_bytes(&es_b[32 * (b * k + (k - 1 - r))], es[r]);
            }
            sect.lap(S_V_ORACLE);
            if (dchi) {
                v.e_inv = h64::from_bytes(dchi + 32 * b * n_ch);
                v.q0_inv = q0_of(s->arg, h64::from_bytes(dchi + 32 * (b * n_ch + 3)));    // 1 / (+-q^2) = +-(1/q)^2
            } else {
                Fr iv[2] = {v.e, v.q0};
                h64::batch_inv(iv, 2);
                v.e_inv = iv[0]; v.q0_inv = iv[1];
            }
            uint8_t* o = ch + 32 * 8 * b;
            h64::to_bytes(o, v.e); h64::to_bytes(o + 32, v.e_inv); h64::to_bytes(o + 64, v.x); h64::to_bytes(o + 96, v.xq);
            h64::to_bytes(o + 128, v.q0); h64::to_bytes(o + 160, v.q0_inv); h64::to_bytes(o + 192, v.t);
            memset(o + 224, 0, 32);
            sect.lap(S_V_MISC);
        });
        g_tm.lap("verify_host");
        int rc = bppp_trrp_verify_pub(ln.trrp, B, ch, small);
        if (rc) return fail(s, rc, std::string("bppp_trrp_verify_pub: ") + ctx_err(ln));
        const uint8_t* shc = nullptr;
        if (dch && s->n_shared) {
            uint8_t* o = lane_buf(s, ln, PB_SHC, B * s->n_shared * 32);
            rc = bppp_trrp_shared_coeffs(ln.trrp, 1, o);
            if (rc) return fail(s, rc, std::string("shared coefficients: ") + ctx_err(ln));
            shc = o;
        }
        g_tm.lap("verify_pub");
        parallel_for(B, [&](size_t b) {
            using namespace h64;
            Sect sect;
            const VP& v = V[b];
            const uint8_t* cm = coms + 64 * b * NC;
            const Fr ts0 = from_bytes(small + 32 * (3 * b)), sum_q2 = from_bytes(small + 32 * (3 * b + 1)), sum_v = from_bytes(small + 32 * (3 * b + 2));
            const Fr t2 = sqr(v.t), t3 = mul(t2, v.t), t5 = mul(sqr(t2), v.t), two_t5 = dbl(t5);
            const Fr sp = add(public_consts_z_trrp(s, v.e, v.x, two_t5), add(ts0, mul(two_t5, add(sum_q2, mul(v.e_inv, sum_v)))));
            sect.lap(S_V_PUB);
            std::vector<Fr> shared;
            if (shc) {
                shared.resize(s->n_shared);
                memcpy((void*)shared.data(), shc + 32 * b * s->n_shared, 32 * s->n_shared);
            } else {
                std::map<U128, Fr> bm = make_base_map(s, v.x);
                shared = make_shared_coeffs(v.e, v.e_inv, s->m_bases, bm);
            }
            std::vector<Fr> cs = make_bp_coeffs(s->flag, v.xq, v.r0, v.r1, v.t, shared);
            // TranscriptTRRP.openWith (TypedReciprocal.hs:279-282): [1,t,t^2,t^3] on [bl,m,dm,r]
            std::vector<Fr> init_s = {one(), t3, t2, v.t};                  // coms order: bl, r, dm, m
            for (auto& c : input_coeffs_trrp(s, v.x, v.q0)) init_s.push_back(mul(two_t5, c));
            sect.lap(S_V_MISC);
            memset(&c_b[32 * b * M], 0, 32 * M);
            to_bytes(&q_b[32 * b], v.q);
            to_bytes(&sp_b[32 * b], sp);
            for (size_t i = 0; i < cs.size() && i < M; i++) to_bytes(&c_b[32 * (b * M + i)], cs[i]);
            for (size_t i = 0; i < NC; i++) to_bytes(&is_b[32 * (b * NC + i)], init_s[i]);
            memcpy(&ip_b[64 * b * NC], cm, 64 * NC);
            memcpy(&fw_b[32 * b * n_norm], finals + 32 * b * (n_norm + n_lin), 32 * n_norm);
            memcpy(&fl_b[32 * b * n_lin], finals + 32 * (b * (n_norm + n_lin) + n_norm), 32 * n_lin);
            sect.lap(S_TOBYTES);
        });
        g_tm.lap("verify_host");
        std::vector<uint8_t> wts;
        if (s->batch_verify && B > 1) wts = batch_weights(B);
        if (!wts.empty()) {
            rc = bppp_nl_verify_trrp_rlc(ln.trrp, k, q_b, sp_b, c_b, es_b, responses, n_norm, n_lin, fw_b, fl_b, NC, is_b, ip_b, wts.data(), ok);
        } else
        rc = bppp_nl_verify_trrp(ln.trrp, k, q_b, sp_b, c_b, es_b, responses, n_norm, n_lin, fw_b, fl_b, NC, is_b, ip_b, ok);
        g_tm.lap("nl_verify");
        g_tm.dump("verify");
        if (t_lane_threads_is_main()) dump_sections("verify, all lanes", B);
        if (rc) return fail(s, rc, std::string("bppp_nl_verify: ") + ctx_err(ln));
        return BPPP_OK;
    }
    std::vector<Ph1> ph1v;
    if (!s->binary) ph1v = ph1s_verifier(s);
    parallel_for(B, [&](size_t b) {
        tr::Zkpt zk;
        Sect sect;
        zk.fmt = s->fmt;
        zk.no_random = true;
        const uint8_t* cm = coms + 64 * b * NC;
        Fr q;
        RPW pub;
        std::vector<Fr> cs, init_s;
        if (s->binary) {                                                    // verifyBRPM (Binary.hs:205-220)
            Fr ch[3], t;
            zk.oracle(cm + 64, 1 + n, ch, 3);
            q = ch[0];
            Fr x = ch[1], r = ch[2];
            Fr q0 = q0_of(s->arg, q), q0_inv = h64::inv(q0);
            zk.oracle(cm, 1, &t, 1);
            RPW pw = public_consts_brp(s, x, q0, q0_inv);
            // pub = t *^ RPW (t * pubSc) [] pubNrm
            pub.sc = h64::mul(h64::sqr(t), pw.sc);
            for (auto& v : pw.nrm) pub.nrm.push_back(h64::mul(t, v));
            cs = {h64::zero(), h64::mul(r, t)};
            Fr two_t2 = h64::dbl(h64::sqr(t));                              // TranscriptBRP.openWith (Binary.hs:106-110)
            for (auto& c : input_coeffs_brp(s, x)) init_s.push_back(h64::mul(two_t2, c));
            // opening order here: [blCom (1), dCom (t), nComs ...] matching `coms`
            init_s.insert(init_s.begin(), t);
            init_s.insert(init_s.begin(), h64::one());
        } else {                                                            // verifyTRRPM (TypedReciprocal.hs:447-467)
            Fr ch[3], ch2[3], t;
            zk.oracle(cm + 128, 2 + n, ch, 3);
            Fr e = ch[0], x = ch[1], r0 = ch[2];
            zk.oracle(cm + 64, 1, ch2, 3);
            q = ch2[0];
            Fr xq = ch2[1], r1 = ch2[2];
            Fr q0 = q0_of(s->arg, q);
            zk.oracle(cm, 1, &t, 1);
            sect.lap(S_V_ORACLE);
            Fr iv[3] = {e, q, q0};
            h64::batch_inv(iv, 3);
            Fr e_inv = iv[0], q0_inv = iv[2];
            std::map<U128, Fr> bm = make_base_map(s, x);
            std::vector<Ph2> ph2s = make_phase2s(false, e, e_inv, x, bm, ph1v);
            sect.lap(S_V_PHASE2);
            pub = make_public_consts_trrp(s, e, e_inv, x, xq, q0, q0_inv, t, ph2s);
            sect.lap(S_V_PUB);
            cs = make_bp_coeffs(s->flag, xq, r0, r1, t, make_shared_coeffs(e, e_inv, s->m_bases, bm));
            // TranscriptTRRP.openWith (TypedReciprocal.hs:279-282): [1,t,t^2,t^3] on [bl,m,dm,r]
            Fr t2 = h64::sqr(t), t3 = h64::mul(t2, t), t5 = h64::mul(h64::sqr(t2), t);
            init_s = {h64::one(), t3, t2, t};                               // coms order: bl, r, dm, m
            Fr two_t5 = h64::dbl(t5);
            for (auto& c : input_coeffs_trrp(s, x, q0)) init_s.push_back(h64::mul(two_t5, c));
        }
        sect.lap(S_V_MISC);
        // challenges of the argument: oldest round hashed first, list newest first (Bulletproof.hs:374)
        {
            std::vector<const uint8_t*> rp(k);
            std::vector<Fr> es(k);
            for (size_t r = 0; r < k; r++) rp[r] = responses + 128 * (b * k + (k - 1 - r));   // oldest round sits last
            zk.oracle_rounds(rp.data(), k, es.data());
            for (size_t r = 0; r < k; r++) h64::to_bytes(&es_b[32 * (b * k + (k - 1 - r))], es[r]);
        }
        sect.lap(S_V_ORACLE);
        memset(&pw_b[32 * b * N], 0, 32 * N);
        memset(&c_b[32 * b * M], 0, 32 * M);
        h64::to_bytes(&q_b[32 * b], q);
        h64::to_bytes(&sp_b[32 * b], pub.sc);
        for (size_t i = 0; i < pub.nrm.size() && i < N; i++) h64::to_bytes(&pw_b[32 * (b * N + i)], pub.nrm[i]);
        for (size_t i = 0; i < cs.size() && i < M; i++) h64::to_bytes(&c_b[32 * (b * M + i)], cs[i]);
        for (size_t i = 0; i < NC; i++) h64::to_bytes(&is_b[32 * (b * NC + i)], init_s[i]);
        memcpy(&ip_b[64 * b * NC], cm, 64 * NC);
        memcpy(&fw_b[32 * b * n_norm], finals + 32 * b * (n_norm + n_lin), 32 * n_norm);
        memcpy(&fl_b[32 * b * n_lin], finals + 32 * (b * (n_norm + n_lin) + n_norm), 32 * n_lin);
        sect.lap(S_TOBYTES);
    });
    g_tm.lap("verify_host");
    int rc;
    std::vector<uint8_t> wts;
    if (s->batch_verify && B > 1) wts = batch_weights(B);
    if (!wts.empty()) {
        rc = bppp_nl_verify_gens_rlc(ln.gens, s->arg, B, k, q_b, sp_b, pw_b, c_b, es_b, responses,
                                     n_norm, n_lin, fw_b, fl_b, NC, is_b, ip_b, wts.data(), ok);
    } else
    rc = bppp_nl_verify_gens(ln.gens, s->arg, B, k, q_b, sp_b, pw_b, c_b, es_b, responses,
                             n_norm, n_lin, fw_b, fl_b, NC, is_b, ip_b, ok);
    g_tm.lap("nl_verify");
    g_tm.dump("verify");
    if (t_lane_threads_is_main()) dump_sections("verify, all lanes", B);
    if (rc) return fail(s, rc, std::string("bppp_nl_verify: ") + ctx_err(ln));
    return BPPP_OK;
}

}  // extern "C"
namespace {
std::vector<Lane> make_lanes(bppp_rp* s, size_t batch) {
    std::vector<Lane> L;
    L.push_back({s->ctx, s->fb, s->gens, 0, 0, s->trrp});
    for (size_t i = 0; i < s->lane_ctx.size(); i++)
        L.push_back({s->lane_ctx[i], s->lane_fb[i], s->lane_gens[i], 0, (int)i + 1, i < s->lane_trrp.size() ? s->lane_trrp[i] : nullptr});
    if (s->pinned.size() < L.size()) s->pinned.resize(L.size(), std::vector<bppp_rp::Pinned>(PB_COUNT));
    if (s->lane_vtr.size() < L.size()) s->lane_vtr.resize(L.size(), nullptr);
    size_t want = std::max<size_t>(1, std::min(L.size(), batch / 32));      // tiny batches: one lane
    L.resize(want);
    int total = g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency();
    if (total < 1) total = 4;
    // every lane may use all host threads: its phases are jobs on the shared worker pool
    const char* ev = getenv("BPPP_LANE_THREADS");
    int per = ev ? atoi(ev) : total;
    for (auto& l : L) l.threads = std::max(1, per);
    return L;
}
template <class F>
int run_lanes(bppp_rp* s, size_t batch, F fn) {
    std::vector<Lane> L = make_lanes(s, batch);
    if (L.size() == 1) return fn(L[0], (size_t)0, batch);
    std::vector<int> rcs(L.size(), 0);
    std::vector<std::thread> th;
    size_t per = (batch + L.size() - 1) / L.size();
    for (size_t i = 0; i < L.size(); i++) {
        size_t b0 = std::min(batch, i * per), nb = std::min(per, batch - b0);
        if (!nb) continue;
        th.emplace_back([&, i, b0, nb]() {
            t_lane_threads = L[i].threads;
            t_lane_id = (int)i;
            t_is_lane0 = (i == 0);
            // the short per-proof loops inside the device calls (round constants, rationalReduceScalar):
            // the lanes already run side by side, so each gets its share of the cores, not all of them
            bppp_set_thread_host_threads(std::max(1, L[i].threads / (int)L.size()));
            rcs[i] = fn(L[i], b0, nb);
        });
    }
    for (auto& t : th) t.join();
    g_tm.dump_trace("lanes");
    for (int rc : rcs)
        if (rc) return rc;
    return BPPP_OK;
}
}  // namespace
extern "C" {

// proveBPM (src/Bulletproof.hs:357-359) over a device-resident argument: `rounds` times
//   (X, R) <- the two commitments; e <- head <$> oracle [X, R] (Bulletproof.hs:351); collapse e
// with the reference's Fiat-Shamir transcript (shaOracle over the commitment list, app/Main.hs:75-80,
// src/ZKP.hs:96-101) kept on the host.  init_pts = [batch][n_init] commitments already in the transcript,
// NEWEST FIRST (e.g. the range-proof commitments, or the initial commitment of a bare argument).
// responses = [batch][rounds][2] points and es = [batch][rounds] challenges (may be NULL), newest first.
int bppp_nl_prove(bppp_nl* h, size_t batch, int show_format, size_t n_init, const uint8_t* init_pts, size_t rounds,
                  uint8_t* responses, uint8_t* es) {
    if (!h || !responses || batch == 0 || (n_init && !init_pts)) return BPPP_ERR_ARG;
    std::vector<tr::Zkpt> zk(batch);
    std::vector<uint8_t> X(batch * 64), R(batch * 64), E(batch * 32);
    for (size_t b = 0; b < batch; b++) {
        zk[b].fmt = show_format;
        zk[b].no_random = true;
        if (n_init) zk[b].absorb(init_pts + 64 * b * n_init, n_init);
    }
    for (size_t r = 0; r < rounds; r++) {
        int rc = bppp_nl_round_commit(h, X.data(), R.data());
        if (rc) return rc;
        parallel_for((batch + 1) / 2, [&](size_t pi) {                  // two transcripts per task (two-stream SHA)
            const size_t b0 = 2 * pi, nb = std::min<size_t>(2, batch - b0);
            uint8_t xr[2][128];
            Fr e[2];
            for (size_t j = 0; j < nb; j++) {
                memcpy(xr[j], &X[64 * (b0 + j)], 64);
                memcpy(xr[j] + 64, &R[64 * (b0 + j)], 64);
            }
            if (nb == 2) tr::Zkpt::oracle_pair(zk[b0], xr[0], zk[b0 + 1], xr[1], 2, &e[0], &e[1]);
            else zk[b0].oracle(xr[0], 2, &e[0], 1);
            for (size_t j = 0; j < nb; j++) {
                h64::to_bytes(&E[32 * (b0 + j)], e[j]);
                memcpy(responses + 128 * ((b0 + j) * rounds + (rounds - 1 - r)), xr[j], 128);
                if (es) h64::to_bytes(es + 32 * ((b0 + j) * rounds + (rounds - 1 - r)), e[j]);
            }
        });
        rc = bppp_nl_round_fold(h, E.data());
        if (rc) return rc;
    }
    return BPPP_OK;
}
// The verifier's half of the same transcript (verifyBPM, src/Bulletproof.hs:370-378): the challenges of
// `rounds` responses (newest first), es = [batch][rounds] newest first.
int bppp_nl_challenges(size_t batch, int show_format, size_t n_init, const uint8_t* init_pts, size_t rounds,
                       const uint8_t* responses, uint8_t* es) {
    if (!es || batch == 0 || (n_init && !init_pts) || (rounds && !responses)) return BPPP_ERR_ARG;
    parallel_for(batch, [&](size_t b) {
        tr::Zkpt zk;
        zk.fmt = show_format;
        zk.no_random = true;
        if (n_init) zk.absorb(init_pts + 64 * b * n_init, n_init);
        std::vector<const uint8_t*> rp(rounds);
        std::vector<Fr> e(rounds);
        for (size_t r = 0; r < rounds; r++) rp[r] = responses + 128 * (b * rounds + (rounds - 1 - r));   // oldest round sits last
        zk.oracle_rounds(rp.data(), rounds, e.data());
        for (size_t r = 0; r < rounds; r++) h64::to_bytes(es + 32 * (b * rounds + (rounds - 1 - r)), e[r]);
    });
    return BPPP_OK;
}

int bppp_rp_prove_batch(bppp_rp* s, size_t batch, const uint8_t* values, const uint8_t* types, const uint8_t* blinds,
                        const char* const* random_seeds, uint8_t* coms, uint8_t* responses, uint8_t* finals) {
    if (!s) return BPPP_ERR_ARG;
    if (!values || !random_seeds || !coms || !responses || !finals || batch == 0) return fail(s, BPPP_ERR_ARG, "null/empty argument");
    const size_t n = s->n_inputs, NC = s->num_rp_coms + n, nf = s->prover_fin_n + s->prover_fin_l, k = s->prover_rounds;
    return run_lanes(s, batch, [&](const Lane& ln, size_t b0, size_t nb) {
        return prove_impl(s, ln, nb, values + 32 * b0 * n, types ? types + 32 * b0 * n : nullptr,
                          blinds ? blinds + 32 * b0 * n : nullptr, random_seeds + b0, coms + 64 * b0 * NC,
                          responses + 128 * b0 * k, finals + 32 * b0 * nf);
    });
}
int bppp_rp_verify_batch(bppp_rp* s, size_t batch, size_t rounds, size_t n_norm, size_t n_lin, const uint8_t* coms,
                         const uint8_t* responses, const uint8_t* finals, int* ok) {
    if (!s) return BPPP_ERR_ARG;
    if (!coms || !finals || !ok || batch == 0 || (rounds && !responses)) return fail(s, BPPP_ERR_ARG, "null/empty argument");
    // The shape of a proof is fixed by the setup (decodeProof' derives it with optimalWitnessSize,
    // src/RangeProof.hs:70-71); the Binary prover's own round rule (Binary.hs:195) is accepted as well.
    // Anything else is refused: extra rounds / final scalars would not be bound by the check.
    const bool shape_prover = rounds == s->prover_rounds && n_norm == s->prover_fin_n && n_lin == s->prover_fin_l;
    const bool shape_decoder = rounds == s->rounds && n_norm == s->fin_n && n_lin == s->fin_l;
    if (!shape_prover && !shape_decoder) return fail(s, BPPP_ERR_ARG, "proof shape (rounds, final witness lengths) differs from the setup's");
    const size_t NC = s->num_rp_coms + s->n_inputs;
    return run_lanes(s, batch, [&](const Lane& ln, size_t b0, size_t nb) {
        return verify_impl(s, ln, nb, rounds, n_norm, n_lin, coms + 64 * b0 * NC, responses + 128 * b0 * rounds,
                           finals + 32 * b0 * (n_norm + n_lin), ok + b0);
    });
}
// the contexts of all lanes (lane 0 = the setup's own): for profile / launch-count aggregation
int bppp_rp_contexts(bppp_rp* s, bppp_ctx** out, size_t cap, size_t* count) {
    if (!s || !count) return BPPP_ERR_ARG;
    *count = 1 + s->lane_ctx.size();
    if (out) {
        if (cap < *count) return BPPP_ERR_ARG;
        out[0] = s->ctx;
        for (size_t i = 0; i < s->lane_ctx.size(); i++) out[1 + i] = s->lane_ctx[i];
    }
    return BPPP_OK;
}

// ---- wire format (src/Encoding.hs:75-134, src/RangeProof.hs:60-85)
namespace {
void put_field_be(uint8_t* out, const uint8_t le[32]) {          // four big-endian Word64, least significant word first
    for (int w = 0; w < 4; w++)
        for (int k = 0; k < 8; k++) out[8 * w + k] = le[8 * w + 7 - k];
}
bool y_is_larger(const uint8_t y_le[32]) {                        // y > q - y  (getXAndSign, Encoding.hs:117-122)
    u256 y = host::from_bytes(y_le), ny = fq::neg(y);
    return !u256_geq(ny, y);
}
size_t put_commitments(uint8_t* out, const uint8_t* const* pts, size_t n) {
    size_t ns = (n + 7) / 8;
    memset(out, 0, ns);
    for (size_t i = 0; i < n; i++) {
        if (y_is_larger(pts[i] + 32)) out[i / 8] |= (uint8_t)(1u << (i % 8));
        put_field_be(out + ns + 32 * i, pts[i]);
    }
    return ns + 32 * n;
}
}  // namespace

// size in bytes of proof.bin / commits.bin for this setup
int bppp_rp_encoded_sizes(bppp_rp* s, size_t* proof_bytes, size_t* commits_bytes) {
    if (!s) return BPPP_ERR_ARG;
    size_t npts = s->num_rp_coms + 2 * s->prover_rounds, nsc = s->prover_fin_n + s->prover_fin_l;
    if (proof_bytes) *proof_bytes = 32 * nsc + (npts + 7) / 8 + 32 * npts;
    if (commits_bytes) *commits_bytes = (s->n_inputs + 7) / 8 + 32 * s->n_inputs;
    return BPPP_OK;
}
// encodeProof' (src/RangeProof.hs:60-66): proof.bin = scalars ++ signs ++ xs of (rpComs ++ bpComs),
// commits.bin = signs ++ xs of the input commitments.  Inputs in the layout of bppp_rp_prove_batch.
int bppp_rp_encode_batch(bppp_rp* s, size_t batch, const uint8_t* coms, const uint8_t* responses, const uint8_t* finals,
                         uint8_t* proof_bin, uint8_t* commits_bin) {
    if (!s || !coms || !responses || !finals || !proof_bin || !commits_bin) return BPPP_ERR_ARG;
    const size_t n = s->n_inputs, k = s->num_rp_coms, NC = k + n, rounds = s->prover_rounds;
    const size_t nsc = s->prover_fin_n + s->prover_fin_l, npts = k + 2 * rounds;
    size_t pb, cb;
    bppp_rp_encoded_sizes(s, &pb, &cb);
    parallel_for(batch, [&](size_t b) {
        uint8_t* out = proof_bin + b * pb;
        for (size_t i = 0; i < nsc; i++) put_field_be(out + 32 * i, finals + 32 * (b * nsc + i));
        std::vector<const uint8_t*> pts;
        for (size_t i = 0; i < k; i++) pts.push_back(coms + 64 * (b * NC + i));
        for (size_t i = 0; i < 2 * rounds; i++) pts.push_back(responses + 64 * (b * 2 * rounds + i));
        put_commitments(out + 32 * nsc, pts.data(), npts);
        pts.clear();
        for (size_t i = 0; i < n; i++) pts.push_back(coms + 64 * (b * NC + k + i));
        put_commitments(commits_bin + b * cb, pts.data(), n);
    });
    return BPPP_OK;
}
// decodeProof' (src/RangeProof.hs:68-85) + decodeCommitments / fromXWithSign (Encoding.hs:97-128):
// x-only points are decompressed with one Fq square root each (host; batchable on the device later).
// ok[b] = 0 when some x is not on the curve; scalars are reduced mod r like the reference's toP (Encoding.hs:75-79).
int bppp_rp_decode_batch(bppp_rp* s, size_t batch, const uint8_t* proof_bin, const uint8_t* commits_bin, uint8_t* coms,
                         uint8_t* responses, uint8_t* finals, int* ok) {
    if (!s || !proof_bin || !commits_bin || !coms || !responses || !finals || !ok) return BPPP_ERR_ARG;
    const size_t n = s->n_inputs, k = s->num_rp_coms, NC = k + n, rounds = s->prover_rounds;
    const size_t nsc = s->prover_fin_n + s->prover_fin_l, npts = k + 2 * rounds;
    size_t pb, cb;
    bppp_rp_encoded_sizes(s, &pb, &cb);
    u256 e4 = fq::modulus();                                       // (q + 1) / 4
    {
        u256 t;
        u256_add(t, e4, u256_one());
        for (int i = 0; i < 8; i++) e4.v[i] = (t.v[i] >> 2) | (i < 7 ? t.v[i + 1] << 30 : 0);
    }
    auto get_field = [](const uint8_t* in, uint8_t le[32]) {
        for (int w = 0; w < 4; w++)
            for (int kk = 0; kk < 8; kk++) le[8 * w + 7 - kk] = in[8 * w + kk];
    };
    auto decompress = [&](const uint8_t* xin, bool larger, uint8_t* out) -> bool {
        uint8_t le[32];
        get_field(xin, le);
        u256 x = host::from_bytes(le);
        x = fq::cond_sub(x, 0);                                    // toP
        u256 seven = u256_zero();
        seven.v[0] = 7;
        u256 rhs = fq::add(fq::mul(fq::sqr(x), x), seven), y = u256_one();
        for (int i = 255; i >= 0; i--) { y = fq::sqr(y); if (u256_bit(e4, i)) y = fq::mul(y, rhs); }
        if (!u256_eq(fq::sqr(y), rhs)) return false;
        u256 ny = fq::neg(y);
        bool y_larger = !u256_geq(ny, y);
        if (y_larger != larger) y = ny;
        host::to_bytes(out, x);
        host::to_bytes(out + 32, y);
        return true;
    };
    parallel_for(batch, [&](size_t b) {
        const uint8_t* in = proof_bin + b * pb;
        bool good = true;
        for (size_t i = 0; i < nsc; i++) {
            uint8_t le[32];
            get_field(in + 32 * i, le);
            uint64_t c[4];
            memcpy(c, le, 32);
            h64::Fr v = h64::from_wide(c);                          // toP (mod r)
            h64::to_bytes(finals + 32 * (b * nsc + i), v);
        }
        const uint8_t* sg = in + 32 * nsc;
        const uint8_t* xs = sg + (npts + 7) / 8;
        for (size_t i = 0; i < npts; i++) {
            uint8_t* dst = i < k ? coms + 64 * (b * NC + i) : responses + 64 * (b * 2 * rounds + (i - k));
            good &= decompress(xs + 32 * i, (sg[i / 8] >> (i % 8)) & 1, dst);
        }
        const uint8_t* cs = commits_bin + b * cb;
        const uint8_t* cx = cs + (n + 7) / 8;
        for (size_t i = 0; i < n; i++) good &= decompress(cx + 32 * i, (cs[i / 8] >> (i % 8)) & 1, coms + 64 * (b * NC + k + i));
        ok[b] = good ? 1 : 0;
    });
    return BPPP_OK;
}

// self-test hooks for the CPU test-suite (no device needed)
int bppp_host_sha256(const uint8_t* data, size_t n, uint8_t out[32]) {
    sha::digest3(out, data, n, nullptr, 0, nullptr, 0);
    return 0;
}
int bppp_host_oracle(const uint8_t* pts, size_t npts, int count, int show_format, uint8_t* out) {
    if (!h64::host_cpu_ok()) return BPPP_ERR_STATE;
    tr::Zkpt zk;
    zk.fmt = show_format;
    std::vector<Fr> o(count);
    zk.oracle(pts, npts, o.data(), count);
    for (int i = 0; i < count; i++) h64::to_bytes(out + 32 * i, o[i]);
    return 0;
}
// A prover-style transcript run through the batched paths the range-proof layer uses: `n_random`
// values of `random` (two-stream one-block hashing; canonical form), then the commitments `pts`,
// then `rounds` (X, R) pairs whose challenges come from oracle_rounds (verifier style) and, on a
// second transcript, from oracle_pair (prover style, two transcripts at a time).
// out: [n_random] randoms | [rounds] challenges (oracle_rounds) | [rounds] challenges (oracle_pair); 32 B each
int bppp_host_transcript(const char* seed, int show_format, size_t n_random, const uint8_t* pts, size_t npts,
                         const uint8_t* round_pts, size_t rounds, uint8_t* out) {
    if (!h64::host_cpu_ok()) return BPPP_ERR_STATE;
    if (!seed || !out || (npts && !pts) || (rounds && !round_pts)) return BPPP_ERR_ARG;
    tr::Zkpt a, b, c;
    a.fmt = b.fmt = c.fmt = show_format;
    a.seed = b.seed = c.seed = seed;
    std::vector<Fr> rnd(n_random);
    a.random_fill(rnd.data(), n_random);
    std::vector<uint8_t> canon(32 * n_random);
    b.random_fill_canonical(canon.data(), n_random);
    for (size_t i = 0; i < n_random; i++) {
        h64::to_bytes(out + 32 * i, rnd[i]);
        if (memcmp(out + 32 * i, &canon[32 * i], 32)) return BPPP_ERR_STATE;
    }
    Fr ch;
    if (npts) { a.oracle(pts, npts, &ch, 1); b.oracle(pts, npts, &ch, 1); c.oracle(pts, npts, &ch, 1); }
    std::vector<const uint8_t*> rp(rounds);
    for (size_t r = 0; r < rounds; r++) rp[r] = round_pts + 128 * r;
    std::vector<Fr> es(rounds);
    a.oracle_rounds(rp.data(), rounds, es.data());
    for (size_t r = 0; r < rounds; r++) h64::to_bytes(out + 32 * (n_random + r), es[r]);
    for (size_t r = 0; r < rounds; r++) {
        Fr eb, ec;
        tr::Zkpt::oracle_pair(b, rp[r], c, rp[r], 2, &eb, &ec);
        if (!(eb == ec)) return BPPP_ERR_STATE;
        h64::to_bytes(out + 32 * (n_random + rounds + r), eb);
    }
    return BPPP_OK;
}
// The host job scheduler under load: `lanes` threads each submit `jobs` jobs of `items` items with
// rising priorities (like lanes advancing through their phases).  Every item must run exactly once
// and every submitter must come back.  Returns 0, or the number of items with a wrong count.
int bppp_host_scheduler_selftest(int lanes, int jobs, int items) {
    if (lanes <= 0 || jobs <= 0 || items <= 0) return -1;
    std::vector<std::atomic<int>> hits((size_t)lanes * jobs * items);
    for (auto& h : hits) h.store(0);
    std::vector<std::thread> th;
    for (int l = 0; l < lanes; l++)
        th.emplace_back([&, l] {
            for (int j = 0; j < jobs; j++) {
                std::function<void(size_t)> f = [&, l, j](size_t i) { hits[((size_t)l * jobs + j) * items + i]++; };
                pool().run((size_t)items, (int64_t)j * 1024 - l, f);
            }
        });
    for (auto& t : th) t.join();
    int bad = 0;
    for (auto& h : hits) bad += h.load() != 1;
    return bad;
}
int bppp_host_fr(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    if (!h64::host_cpu_ok()) return BPPP_ERR_STATE;
    Fr x = h64::from_bytes(a), y = h64::from_bytes(b), r;
    switch (op) {
        case 0: r = h64::mul(x, y); break;
        case 1: r = h64::add(x, y); break;
        case 2: r = h64::sub(x, y); break;
        case 3: r = h64::inv(x); break;
        case 4: r = h64::neg(x); break;
        default: return 1;
    }
    h64::to_bytes(out, r);
    return 0;
}
int bppp_host_get_points(const char* seed, size_t count, int root_policy, uint8_t* out) {
    std::vector<Affine> p = tr::get_points(seed, count, root_policy);
    memcpy(out, p.data(), count * 64);
    return 0;
}

}  // extern "C"
