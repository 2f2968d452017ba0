// ldpc_b200.cu -- C ABI (include/ldpc_b200.h) over the sm_100a kernels.  GPU-only: every entry point
// fails loudly (negative status + ldpc_b200_last_error()) when it cannot run on the device.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdarg.h>
#include <mutex>
#include <atomic>
#include <string>
#include <vector>
#include <unordered_map>
#include <algorithm>
#include <utility>
#include <new>
#include <exception>
#include <thread>
#include <chrono>
#include <memory>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/ldpc_b200.h"
#include "e1_kernels.cuh"
#include "relabel.hpp"
#include "e2_kernels.cuh"
#include "e1_staged.cuh"
#include "e1_overlap.cuh"

using namespace ldpc;

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(LDPC_B200_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                    \
    } while (0)

extern "C" const char* ldpc_b200_last_error(void) { return g_last_error.c_str(); }

// No C++ exception may cross the C ABI (a std::bad_alloc from a host-side planner would otherwise terminate the
// caller's process): every entry point that allocates runs its body through this.
template <class F>
static int guarded(const char* what, F&& body) {
    try { return body(); }
    catch (const std::bad_alloc&) { return fail(LDPC_B200_ENOMEM, "%s: out of host memory", what); }
    catch (const std::exception& e) { return fail(LDPC_B200_EINVAL, "%s: %s", what, e.what()); }
    catch (...) { return fail(LDPC_B200_EINVAL, "%s: unknown C++ exception", what); }
}

extern "C" int ldpc_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int ldpc_b200_counters_len(int iterations) { return iterations + 6; }

// ------------------------------------------------------------------------------------------------
static inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

struct E1Config {
    int words = 0, teams = 0, team_warps = 0;
    bool values = false;
    bool v2c_global = false;
    uint32_t smem = 0;
    uint32_t sm_c2v, sm_v2c, sm_team0, sm_team_stride, sm_x, sm_t, sm_v, sm_p, sm_hbuf, sm_scr;
    int v2c_res = 0; uint32_t sm_vres = 0;   // v2c_global: resident part of the variable-side lists
    bool tm = false; int tm_cpw = 0, tm_trips_c = 0, tm_trips_v = 0;   // socket lists in tensor memory (e1_tmem.cuh)
    // overlapped stage 0 (e1_overlap.cuh): team_warps consumer warps + ov_np producer warps
    int ov_np = 0; uint32_t sm_stage = 0, sm_q = 0, sm_pscr = 0, sm_mbox = 0;
};

struct StagePlan {
    int nstages = 1;
    std::vector<int> cuts;   // [npoints][nstages+1]
    double predicted_inflation = 0;
    unsigned long long pilot_frames = 0;   // frames per point the pilot decoded (a much larger run draws a better one)
    double pilot_fer = 0;                  // largest block error rate over the points, as the pilot saw it
    int* d_cuts = nullptr;                 // device copy of `cuts`, uploaded once when the plan is made
    StagePlan() = default;
    StagePlan(const StagePlan&) = delete;
    StagePlan(StagePlan&& o) noexcept { *this = std::move(o); }
    StagePlan& operator=(StagePlan&& o) noexcept {
        nstages = o.nstages; cuts = std::move(o.cuts); predicted_inflation = o.predicted_inflation;
        pilot_frames = o.pilot_frames; pilot_fer = o.pilot_fer; std::swap(d_cuts, o.d_cuts);
        return *this;
    }
    ~StagePlan() { if (d_cuts) cudaFree(d_cuts); }
};

struct ldpc_b200_graph {
    int n = 0, k = 0, m = 0, dv = 0, dc = 0, device = 0;
    int n_pad = 0, m_pad = 0;
    std::vector<int> v2c, c2v;     // host copies, reference layouts
    uint16_t* d_c2v16 = nullptr;   // packed socket lists (socklist.hpp), dc*m_pad uint16
    uint16_t* d_v2c16 = nullptr;   // packed socket lists, dv*n_pad uint16
    int sm_count = 0, smem_optin = 0;
    // bank-conflict-aware placements for the Monte-Carlo kernel, one per words-per-group (index = words)
    struct Placed {
        bool ready = false;            // device arrays below are valid
        int effort = 0;                // effort of the placement on the device
        uint16_t* d_c2v = nullptr; uint16_t* d_v2c = nullptr; uint16_t* d_var_orig = nullptr;
        double wf_natural = 0, wf_placed = 0, wf_ideal = 0;
        // background search (relabel.cpp runs on host threads while the GPU already decodes on the natural layout)
        std::thread worker;
        std::atomic<int> state{0};     // 0 idle, 1 searching, 2 result waiting to be uploaded
        int search_effort = 0;
        int source = 0;                // 1 = searched in this process, 2 = read from the disk cache
        std::unique_ptr<Placement> result;
    } placed[5];
    uint64_t graph_hash = 0;           // content hash (dims + lookups), keys the placement cache on disk
    cudaEvent_t chunk_ev[2] = {nullptr, nullptr};   // throttle of the chunk loop while a placement search is pending
    int place_effort = -1;  // -1: LDPC_B200_PLACE_EFFORT or default
    // E2 (HBM/L2-streamed engine): int32 lookups in the reference layouts + frame-minor state rows
    int* d_c2v32 = nullptr; int* d_v2c32 = nullptr;
    void* e2_X = nullptr; void* e2_T = nullptr; void* e2_V = nullptr; void* e2_PAR = nullptr;
    size_t e2_xbytes = 0, e2_vbytes = 0;
    uint32_t* e2_acc = nullptr; uint32_t* e2_framecnt = nullptr; unsigned long long* e2_tilectr = nullptr;
    unsigned long long* e2_hbuf = nullptr; size_t e2_hbuf_cap = 0;
    int engine = 0;  // 0 auto, 1 force E1, 2 force E2
    // staged E1 Monte-Carlo: cut points per (eps list, cap, words), spill buffers, per-stage record counts
    std::unordered_map<uint64_t, StagePlan> plans;
    void* d_spill[2] = {nullptr, nullptr}; size_t spill_bytes = 0; int spill_budget_mb = 0;
    uint32_t* d_idlist = nullptr; size_t idlist_bytes = 0;   // expurgated runs: frames listed for the re-count pass
    unsigned long long* d_stage_cnt = nullptr; size_t stage_cnt_cap = 0;
    // batched decode buffers (cached between calls)
    // single-group decode (<= 32 frames: the message_passing drop-in): one pinned host block and one device block
    // [symbols | errors | ret], one upload, ONE kernel, one download
    void* fast_pin = nullptr; void* fast_dev = nullptr; size_t fast_cap = 0;
    std::mutex mu;   // serialises calls on this handle (the drop-in may be called from several host threads)
    void* dec_mvc = nullptr; void* dec_x = nullptr; void* dec_v = nullptr; void* dec_err = nullptr; void* dec_ret = nullptr;
    size_t dec_mvc_cap = 0, dec_x_cap = 0, dec_v_cap = 0, dec_err_cap = 0, dec_ret_cap = 0;
    int launches = 0;      // kernels launched since the last stats read
    int last_stages = 1;
    // workspace
    unsigned long long* d_work = nullptr;
    unsigned long long* d_raw = nullptr; size_t raw_cap = 0;
    BecParams* d_bec = nullptr; int bec_cap = 0;
    // erasure-point parameters go up through a small ring of pinned host slots (no synchronisation in mc_launch:
    // a slot is reused only after the copy that read it has completed)
    static constexpr int BEC_RING = 8;
    BecParams* h_bec[BEC_RING] = {}; int h_bec_cap[BEC_RING] = {}; cudaEvent_t h_bec_ev[BEC_RING] = {}; int h_bec_next = 0;
    unsigned long long* d_out = nullptr; size_t out_cap = 0;  // for mc_run
    // tuning overrides
    int tune_words = 0, tune_teams = 0, tune_team_warps = 0;
    // statistics of launches since last read
    unsigned long long* d_stats = nullptr;  // [8]: phys frame-iterations, second-pass groups, spilled records, re-counted frames,
                                            //      overlapped stage 0: consumer wait / producer wait / draw / consumer busy cycles (thread 0 of SM-resident CTAs)
    ldpc_b200_mc_stats last = {};
    // kernel timing ring
    static constexpr int RING = 256;
    cudaEvent_t ev0[RING], ev1[RING];
    int ev_count = 0;
    bool ev_init = false;
};

static double host_seconds() {   // since the first call (LDPC_B200_TRACE time stamps)
    static const auto t0 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

static int tuning_env(const char* name) {
    const char* s = getenv(name);
    return s ? atoi(s) : 0;
}

// Plan with the socket lists in tensor memory (e1_tmem.cuh): shared memory holds state only; every thread's gather
// addresses must fit its warp's share of the 512 TMEM columns (4 lane quarters x 512 columns x 32 lanes).
static bool plan_e1_tm(const ldpc_b200_graph* g, int cap, E1Config& cfg, int force_words = 0) {
    if (!((g->dv == 3 && g->dc == 6) || (g->dv == 4 && g->dc == 8))) return false;   // degree pairs the kernels are specialised for
    if (g->n > 65535 || tuning_env("LDPC_B200_TMEM") == 2) return false;
    const uint32_t limit = (uint32_t)g->smem_optin - 64;   // the kernels' static shared memory (the TMEM base slot)
    const int want_words = force_words ? force_words : (g->tune_words ? g->tune_words : tuning_env("LDPC_B200_WORDS"));
    const int want_teams = g->tune_teams ? g->tune_teams : tuning_env("LDPC_B200_TEAMS");
    const int want_tw = g->tune_team_warps ? g->tune_team_warps : tuning_env("LDPC_B200_TEAM_WARPS");
    const int cand[3] = {4, 2, 1};
    for (int ci = 0; ci < 3; ci++) {
        const int words = cand[ci];
        if (want_words && words != want_words) continue;
        E1Config c;
        uint32_t off = 0;
        c.sm_x = off; off += align_up((uint32_t)g->n * 4 * words, 16);
        c.sm_t = off; off += align_up((uint32_t)g->m * 4 * words, 16);
        c.sm_v = off; c.sm_p = off;
        c.sm_hbuf = off; off += align_up((uint32_t)(cap + 1) * 4, 16);
        c.sm_scr = off; off += align_up((uint32_t)e1_scr_words(words) * 4, 16);
        const uint32_t tb = off;
        if (tb > limit) continue;
        const int max_warps = e1_threads(words) / 32;
        const int fit = (int)(limit / tb);
        // most warps first; among equals the fewest teams (wider teams amortise the per-group work better)
        int best_teams = 0, best_tw = 0;
        for (int teams = std::min(fit, std::min(max_warps, 15)); teams >= 1; teams--) {
            if (want_teams && teams != std::min(want_teams, fit)) continue;
            for (int tw = max_warps / teams; tw >= 1; tw--) {
                if (want_tw && tw != want_tw) continue;
                const int nthr = tw * 32, wpq = (teams * tw + 3) / 4;
                const int tc = (g->m + nthr - 1) / nthr, tv = (g->n + nthr - 1) / nthr;
                if ((tc * g->dc + tv * g->dv) * wpq > 512) continue;
                if (teams * tw > best_teams * best_tw) { best_teams = teams; best_tw = tw; }
                break;
            }
        }
        if (!best_teams) continue;
        const int nthr = best_tw * 32;
        c.words = words; c.teams = best_teams; c.team_warps = best_tw; c.values = false; c.v2c_global = false;
        c.sm_c2v = 0; c.sm_v2c = 0; c.sm_team0 = 0; c.sm_team_stride = tb; c.smem = tb * best_teams;
        c.tm = true; c.tm_trips_c = (g->m + nthr - 1) / nthr; c.tm_trips_v = (g->n + nthr - 1) / nthr;
        c.tm_cpw = c.tm_trips_c * g->dc + c.tm_trips_v * g->dv;
        cfg = c;
        return true;
    }
    return false;
}

// Plan of the overlapped stage-0 kernel (e1_overlap.cuh) for the word width the staged plan `base` chose: ONE consumer
// team + producer warps, the state plus a staging pattern, the producers' draw queue and scratch in shared memory.
// Used when one team owns the SM anyway (large graphs: nothing else would overlap the ALU-only channel draw with the
// shared-memory-bound decoding).  LDPC_B200_OVERLAP: 0 auto, 1 force where it fits, 2 off; LDPC_B200_PRODUCERS = warps.
static bool plan_e1_overlap(const ldpc_b200_graph* g, int cap, const E1Config& base, E1Config& out) {
    const int mode = tuning_env("LDPC_B200_OVERLAP");
    if (mode == 2 || !base.tm) return false;
    if (mode != 1 && base.teams != 1) return false;
    const int words = base.words;
    int max_warps = ov_threads(words) / 32;
    if (tuning_env("LDPC_B200_OV_THREADS") >= 64) max_warps = std::min(max_warps, tuning_env("LDPC_B200_OV_THREADS") / 32);
    int np = tuning_env("LDPC_B200_PRODUCERS");
    if (np <= 0) np = 8;   // measured at n = 10^4 (profiles/r2b_overlap_sweep.log): 2: 5.5, 4: 7.0, 6: 7.8, 8: 8.1 e7 frames/s (off: 7.1)
    if (np >= max_warps) return false;
    const uint32_t limit = (uint32_t)g->smem_optin - 64;
    E1Config c;
    uint32_t off = 0;
    c.sm_x = off; off += align_up((uint32_t)g->n * 4 * words, 16);
    c.sm_t = off; off += align_up((uint32_t)g->m * 4 * words, 16);
    c.sm_v = off; c.sm_p = off;
    c.sm_hbuf = off; off += align_up((uint32_t)(cap + 1) * 4, 16);
    c.sm_scr = off; off += align_up((uint32_t)e1_scr_words(words) * 4, 16);
    c.sm_stage = off; off += align_up((uint32_t)g->n * 4 * words, 16);
    c.sm_pscr = off; off += align_up((uint32_t)e1_scr_words(words) * 4, 16);
    c.sm_mbox = off; off += 64;
    c.sm_q = off; off += (uint32_t)np * ov_wq_bytes();   // one private draw queue per producer warp
    if (off > limit) return false;
    // consumer warps: as many as the thread budget and the tensor-memory columns allow
    int nc = 0;
    for (int tw = max_warps - np; tw >= 1; tw--) {
        const int nthr = tw * 32, wpq = (tw + 3) / 4;
        const int tc = (g->m + nthr - 1) / nthr, tv = (g->n + nthr - 1) / nthr;
        if ((tc * g->dc + tv * g->dv) * wpq > 512) continue;
        nc = tw; break;
    }
    if (nc < 4) return false;
    const int nthr = nc * 32;
    c.words = words; c.teams = 1; c.team_warps = nc; c.values = false; c.v2c_global = false;
    c.sm_c2v = 0; c.sm_v2c = 0; c.sm_team0 = 0; c.sm_team_stride = off; c.smem = off;
    c.tm = true; c.tm_trips_c = (g->m + nthr - 1) / nthr; c.tm_trips_v = (g->n + nthr - 1) / nthr;
    c.tm_cpw = c.tm_trips_c * g->dc + c.tm_trips_v * g->dv;
    c.ov_np = np;
    out = c;
    return true;
}

// Shared-memory plan for E1.  Returns false if even one W=32 team does not fit.
static bool plan_e1(const ldpc_b200_graph* g, int cap, bool values, bool decode, E1Config& cfg) {
    if (!values && !decode && plan_e1_tm(g, cap, cfg)) return true;
    const uint32_t limit = (uint32_t)g->smem_optin;
    const uint32_t graph_bytes = align_up((uint32_t)g->dc * g->m_pad * 2, 16) + align_up((uint32_t)g->dv * g->n_pad * 2, 16);
    if (g->n > 65535) return false;
    auto team_bytes = [&](int words, E1Config& c) {
        uint32_t off = 0;
        c.sm_x = off; off += align_up((uint32_t)g->n * 4 * words, 16);
        c.sm_t = off; off += align_up((uint32_t)g->m * 4 * words, 16);
        c.sm_v = off; if (values) off += align_up((uint32_t)g->n * 4 * words, 16);
        c.sm_p = off; if (values) off += align_up((uint32_t)g->m * 4 * words, 16);
        c.sm_hbuf = off; off += align_up((uint32_t)(cap + 1) * 4, 16);
        c.sm_scr = off; off += align_up((uint32_t)e1_scr_words(words) * 4, 16);
        return off;
    };
    int want_words = g->tune_words ? g->tune_words : tuning_env("LDPC_B200_WORDS");
    int want_teams = g->tune_teams ? g->tune_teams : tuning_env("LDPC_B200_TEAMS");
    int want_tw = g->tune_team_warps ? g->tune_team_warps : tuning_env("LDPC_B200_TEAM_WARPS");

    // candidates in order of preference: (words, variable-side lookups streamed from L2?)
    // A wider group amortises index traffic and instructions over more frames; streaming v2c frees
    // dv*n*2 bytes of shared memory (n = 10^4: W = 64 fits only that way).
    struct Cand { int words; bool vg; };
    const Cand cand[5] = {{4, false}, {2, false}, {2, true}, {1, false}, {1, true}};
    const int want_vg = tuning_env("LDPC_B200_V2C_GLOBAL");   // 0 auto, 1 force on, 2 force off
    const uint32_t c2v_bytes = align_up((uint32_t)g->dc * g->m_pad * 2, 16);
    for (int ci = 0; ci < 5; ci++) {
        const int words = cand[ci].words;
        const bool vg = cand[ci].vg;
        if (vg && (decode || want_vg == 2)) continue;
        if (!vg && want_vg == 1 && !decode) continue;
        const int max_warps = (decode ? E1_MAX_THREADS : e1_threads(words)) / 32;
        if (decode && words != 1) continue;
        if (want_words && words != want_words) continue;
        const uint32_t gbytes = vg ? c2v_bytes : graph_bytes;
        E1Config c;
        uint32_t tb = team_bytes(words, c);
        if (gbytes + tb > limit) continue;
        int fit = (int)((limit - gbytes) / tb);
        const int min_teams = want_words ? 1 : (words == 4 ? 4 : (words == 2 && !vg ? 2 : 1));
        if (fit < min_teams) continue;
        int teams = std::min(fit, max_warps);
        if (want_teams) teams = std::min(want_teams, fit);
        int tw = want_tw ? want_tw : std::max(1, max_warps / teams);
        if (tw > 1 && teams > 15) teams = 15;  // named barriers 1..15
        while (teams * tw > max_warps && tw > 1) tw--;
        if (teams * tw > max_warps) teams = max_warps / tw;
        if (teams < 1) continue;
        c.words = words; c.teams = teams; c.team_warps = tw; c.values = values; c.v2c_global = vg;
        c.sm_c2v = 0;
        c.sm_v2c = c2v_bytes;
        c.sm_team0 = gbytes;
        c.sm_team_stride = tb;
        c.smem = gbytes + tb * teams;
        c.v2c_res = 0; c.sm_vres = 0;
        if (vg && tuning_env("LDPC_B200_V2C_RES") != 2) {
            // whatever shared memory is left keeps the lists of the first variables resident (whole trips of the team)
            const uint32_t off = align_up(c.smem, 16);
            const uint32_t nthr = (uint32_t)tw * 32u;
            if (limit > off) {
                uint32_t res = (limit - off) / ((uint32_t)g->dv * 2u);
                res = res / nthr * nthr;
                if (res >= (uint32_t)g->n) res = (uint32_t)g->n / nthr * nthr;
                if (res > 0) { c.v2c_res = (int)res; c.sm_vres = off; c.smem = off + res * (uint32_t)g->dv * 2u; }
            }
        }
        cfg = c;
        return true;
    }
    return false;
}

// geometry of a plan (threads, shared-memory map, tensor-memory map) -> kernel parameters; leaves the run fields alone
static void set_geometry(const ldpc_b200_graph* g, const E1Config& c, E1Params& P) {
    P.vec_bytes = 4u * (uint32_t)c.words; P.k16 = 65536u; P.k16s = 65536u * P.vec_bytes;
    P.team_warps = c.team_warps; P.teams = c.teams; P.v2c_global = c.v2c_global ? 1 : 0; P.v2c_res = c.v2c_res; P.sm_vres = c.sm_vres;
    P.sm_c2v = c.sm_c2v; P.sm_v2c = c.sm_v2c; P.sm_team0 = c.sm_team0; P.sm_team_stride = c.sm_team_stride;
    P.sm_x = c.sm_x; P.sm_t = c.sm_t; P.sm_v = c.sm_v; P.sm_p = c.sm_p; P.sm_hbuf = c.sm_hbuf; P.sm_scr = c.sm_scr;
    P.tm = c.tm ? 1 : 0; P.tm_cpw = c.tm_cpw; P.tm_trips_c = c.tm_trips_c; P.tm_trips_v = c.tm_trips_v;
    P.tm_full_c = g->m / (c.team_warps * 32); P.tm_full_v = g->n / (c.team_warps * 32);
    P.ov_np = c.ov_np; P.sm_stage = c.sm_stage; P.sm_q = c.sm_q; P.sm_pscr = c.sm_pscr; P.sm_mbox = c.sm_mbox;
}

static void fill_params(const ldpc_b200_graph* g, const E1Config& c, int cap, E1Params& P) {
    memset(&P, 0, sizeof P);
    P.c2v = g->d_c2v16; P.v2c = g->d_v2c16;
    P.n = g->n; P.m = g->m; P.dv = g->dv; P.dc = g->dc; P.n_pad = g->n_pad; P.m_pad = g->m_pad;
    P.draw_qcap = (uint32_t)std::max(0, tuning_env("LDPC_B200_DRAW_QCAP"));
    set_geometry(g, c, P);
    P.cap = cap;
}

// ------------------------------------------------------------------------------------------------
static int ldpc_b200_graph_create_impl(const int* v2c, const int* c2v, int n, int k, int dv, int dc, int device,
                                      ldpc_b200_graph** out) {
    if (!out) return fail(LDPC_B200_EINVAL, "graph_create: out is NULL");
    *out = nullptr;
    if (!v2c || !c2v) return fail(LDPC_B200_EINVAL, "graph_create: NULL lookup");
    if (n <= 0 || dv <= 0 || dc <= 0 || k < 0 || k >= n) return fail(LDPC_B200_EINVAL, "graph_create: bad sizes n=%d k=%d dv=%d dc=%d", n, k, dv, dc);
    const int m = n - k;
    if ((long long)m * dc != (long long)n * dv)
        return fail(LDPC_B200_EINVAL, "graph_create: (n-k)*dc != n*dv (n=%d k=%d dv=%d dc=%d)", n, k, dv, dc);
    if (dv > 64 || dc > 64) return fail(LDPC_B200_EUNSUPPORTED, "graph_create: degree > 64");
    const long long E = (long long)n * dv;
    for (long long e = 0; e < E; e++) {
        if (v2c[e] < 0 || v2c[e] >= m) return fail(LDPC_B200_EINVAL, "graph_create: variable_to_check_list[%lld]=%d out of range", e, v2c[e]);
        if (c2v[e] < 0 || c2v[e] >= n) return fail(LDPC_B200_EINVAL, "graph_create: check_to_variable_list[%lld]=%d out of range", e, c2v[e]);
    }
    // every check a variable lists must hold that variable in one of its sockets (the kernels use
    // "check c resolves i" <=> "c has exactly one erased socket"; message_passing.c:58 searches the socket)
    for (int i = 0; i < n; i++)
        for (int j = 0; j < dv; j++) {
            const int c = v2c[(long long)i * dv + j];
            bool found = false;
            for (int l = 0; l < dc; l++) found |= (c2v[(long long)c * dc + l] == i);
            if (!found) return fail(LDPC_B200_EINVAL, "graph_create: variable %d lists check %d which has no socket for it", i, c);
        }
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(LDPC_B200_ECUDA, "graph_create: device %d not present (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));

    ldpc_b200_graph* g = new ldpc_b200_graph();
    g->n = n; g->k = k; g->m = m; g->dv = dv; g->dc = dc; g->device = device;
    g->n_pad = (int)align_up((uint32_t)n, 8); g->m_pad = (int)align_up((uint32_t)m, 8);
    g->v2c.assign(v2c, v2c + E); g->c2v.assign(c2v, c2v + E);
    cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&g->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (n <= 65535) {
        std::vector<uint16_t> h_c((size_t)dc * g->m_pad, 0), h_v((size_t)dv * g->n_pad, 0);
        for (int c = 0; c < m; c++) for (int j = 0; j < dc; j++) h_c[sock_off(dc, g->m_pad, j, c)] = (uint16_t)c2v[(size_t)c * dc + j];
        for (int i = 0; i < n; i++) for (int j = 0; j < dv; j++) h_v[sock_off(dv, g->n_pad, j, i)] = (uint16_t)v2c[(size_t)i * dv + j];
        const size_t bc = align_up((uint32_t)(h_c.size() * 2), 16), bv = align_up((uint32_t)(h_v.size() * 2), 16);
        cudaError_t e = cudaMalloc(&g->d_c2v16, bc);
        if (e == cudaSuccess) e = cudaMalloc(&g->d_v2c16, bv);
        if (e == cudaSuccess) e = cudaMemset(g->d_c2v16, 0, bc);
        if (e == cudaSuccess) e = cudaMemset(g->d_v2c16, 0, bv);
        if (e == cudaSuccess) e = cudaMemcpy(g->d_c2v16, h_c.data(), h_c.size() * 2, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(g->d_v2c16, h_v.data(), h_v.size() * 2, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { ldpc_b200_graph_destroy(g); return fail(LDPC_B200_ECUDA, "graph_create: %s", cudaGetErrorString(e)); }
    }
    {
        cudaError_t e = cudaMalloc(&g->d_work, sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMalloc(&g->d_stats, 8 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemset(g->d_stats, 0, 8 * sizeof(unsigned long long));
        if (e != cudaSuccess) { ldpc_b200_graph_destroy(g); return fail(LDPC_B200_ECUDA, "graph_create: %s", cudaGetErrorString(e)); }
    }
    *out = g;
    return LDPC_B200_OK;
}
extern "C" int ldpc_b200_graph_create(const int* v2c, const int* c2v, int n, int k, int dv, int dc, int device,
                                      ldpc_b200_graph** out) {
    return guarded("graph_create", [&] { return ldpc_b200_graph_create_impl(v2c, c2v, n, k, dv, dc, device, out); });
}


extern "C" void ldpc_b200_graph_destroy(ldpc_b200_graph* g) {
    if (!g) return;
    for (auto& pl : g->placed) if (pl.worker.joinable()) pl.worker.join();
    cudaSetDevice(g->device);
    for (auto& e : g->chunk_ev) if (e) cudaEventDestroy(e);
    cudaFree(g->d_c2v16); cudaFree(g->d_v2c16); cudaFree(g->d_work); cudaFree(g->d_raw);
    cudaFree(g->d_bec); cudaFree(g->d_out); cudaFree(g->d_stats);
    for (int i = 0; i < ldpc_b200_graph::BEC_RING; i++) { if (g->h_bec[i]) cudaFreeHost(g->h_bec[i]); if (g->h_bec_ev[i]) cudaEventDestroy(g->h_bec_ev[i]); }
    for (auto& pl : g->placed) { cudaFree(pl.d_c2v); cudaFree(pl.d_v2c); cudaFree(pl.d_var_orig); }
    cudaFree(g->d_c2v32); cudaFree(g->d_v2c32); cudaFree(g->e2_X); cudaFree(g->e2_T); cudaFree(g->e2_V); cudaFree(g->e2_PAR);
    cudaFree(g->e2_acc); cudaFree(g->e2_framecnt); cudaFree(g->e2_hbuf); cudaFree(g->e2_tilectr);
    cudaFree(g->d_spill[0]); cudaFree(g->d_spill[1]); cudaFree(g->d_stage_cnt); cudaFree(g->d_idlist);
    cudaFree(g->dec_mvc); cudaFree(g->dec_x); cudaFree(g->dec_v); cudaFree(g->dec_err); cudaFree(g->dec_ret);
    if (g->fast_pin) cudaFreeHost(g->fast_pin);
    cudaFree(g->fast_dev);
    if (g->ev_init) for (int i = 0; i < ldpc_b200_graph::RING; i++) { cudaEventDestroy(g->ev0[i]); cudaEventDestroy(g->ev1[i]); }
    delete g;
}

extern "C" int ldpc_b200_set_place_effort(ldpc_b200_graph* g, int effort) {
    if (!g) return fail(LDPC_B200_EINVAL, "set_place_effort: NULL graph");
    g->place_effort = effort;
    for (auto& pl : g->placed) {
        if (pl.worker.joinable()) pl.worker.join();
        cudaFree(pl.d_c2v); cudaFree(pl.d_v2c); cudaFree(pl.d_var_orig);
        pl.d_c2v = pl.d_v2c = pl.d_var_orig = nullptr;
        pl.ready = false; pl.effort = 0; pl.state.store(0); pl.result.reset();
    }
    return LDPC_B200_OK;
}

static int ensure_placement(ldpc_b200_graph* g, int words, double frames, bool wait);

extern "C" int ldpc_b200_placement_info(ldpc_b200_graph* g, int words, double* natural, double* placed, double* ideal) {
    if (!g || words < 1 || words > 4) return fail(LDPC_B200_EINVAL, "placement_info: bad argument");
    if (g->placed[words].state.load() != 0) {   // a background search is running or has just finished: take its result
        CUDA_TRY(cudaSetDevice(g->device));
        const int rc = ensure_placement(g, words, 0.0, true);
        if (rc) return rc;
    }
    if (!g->placed[words].ready) return fail(LDPC_B200_EINVAL, "placement_info: no placement for words=%d yet", words);
    if (natural) *natural = g->placed[words].wf_natural;
    if (placed) *placed = g->placed[words].wf_placed;
    if (ideal) *ideal = g->placed[words].wf_ideal;
    return LDPC_B200_OK;
}

static int want_engine(const ldpc_b200_graph* g) {
    if (g->engine) return g->engine;
    const int e = tuning_env("LDPC_B200_ENGINE");
    return (e == 1 || e == 2) ? e : 0;
}

extern "C" int ldpc_b200_set_engine(ldpc_b200_graph* g, int engine) {
    if (!g || engine < 0 || engine > 2) return fail(LDPC_B200_EINVAL, "set_engine: engine must be 0 (auto), 1 (shared memory) or 2 (streamed)");
    g->engine = engine;
    return LDPC_B200_OK;
}

extern "C" int ldpc_b200_set_tuning(ldpc_b200_graph* g, int words, int teams, int team_warps) {
    if (!g) return fail(LDPC_B200_EINVAL, "set_tuning: NULL graph");
    if (words != 0 && words != 1 && words != 2 && words != 4) return fail(LDPC_B200_EINVAL, "set_tuning: words must be 0,1,2,4");
    g->tune_words = words; g->tune_teams = teams; g->tune_team_warps = team_warps;
    return LDPC_B200_OK;
}

// ------------------------------------------------------------------------------------------------
static BecParams make_bec(double eps, uint64_t seed) {
    BecParams bp;
    bp.seed_lo = (uint32_t)seed; bp.seed_hi = (uint32_t)(seed >> 32);
    // u < eps for u = U / 2^32, U uniform 32-bit  <=>  U < T, T = ceil(eps * 2^32) (channels.py:26 is strict)
    if (!(eps > 0.0)) { bp.mode = 0; bp.threshold = 0; }
    else if (eps >= 1.0) { bp.mode = 2; bp.threshold = 0; }
    else {
        double t = ceil(eps * 4294967296.0);
        if (t >= 4294967296.0) { bp.mode = 2; bp.threshold = 0; }
        else { bp.mode = 1; bp.threshold = (uint32_t)t; }
    }
    return bp;
}

template <int WORDS>
static cudaError_t launch_mc_words(const E1Params& P, int grid, int block, uint32_t smem, cudaStream_t st) {
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, block, smem, st>>>(P);
        return cudaGetLastError();
    };
    if (P.tm && P.dv == 3 && P.dc == 6) return go(e1_mc_kernel<WORDS, 3, 6, true>);
    if (P.tm && P.dv == 4 && P.dc == 8) return go(e1_mc_kernel<WORDS, 4, 8, true>);
    if (P.dv == 3 && P.dc == 6) return go(e1_mc_kernel<WORDS, 3, 6>);
    if (P.dv == 4 && P.dc == 8) return go(e1_mc_kernel<WORDS, 4, 8>);
    return go(e1_mc_kernel<WORDS, 0, 0>);
}

__global__ void mc_compact_kernel(const unsigned long long* raw, unsigned long long* out, unsigned long long* stats,
                                  int npoints, int cap) {
    // raw has been finalized; add the public counter block into `out`, perf counters into `stats`
    const int p = blockIdx.x;
    const unsigned long long* C = raw + (size_t)p * raw_stride(cap);
    for (int b = threadIdx.x; b < cap + 6; b += blockDim.x) out[(size_t)p * (cap + 6) + b] += C[b];
    if (threadIdx.x == 0) {
        atomicAdd(stats + 0, C[2 * cap + 8]);
        atomicAdd(stats + 1, C[2 * cap + 9]);
    }
}

static int ensure_events(ldpc_b200_graph* g) {
    if (g->ev_init) return LDPC_B200_OK;
    for (int i = 0; i < ldpc_b200_graph::RING; i++) {
        CUDA_TRY(cudaEventCreate(&g->ev0[i]));
        CUDA_TRY(cudaEventCreate(&g->ev1[i]));
    }
    g->ev_init = true;
    return LDPC_B200_OK;
}

// ---- shared-memory placement of the graph for `words` words per node (relabel.hpp) --------------------------------
// The search costs 0.5 s (n = 1000) to a few seconds (n = 10^4) of host threads.  It no longer delays the first call:
//   * a finished placement is kept on disk, keyed by the graph's content hash, word width and effort
//     ($LDPC_B200_CACHE_DIR, else $XDG_CACHE_HOME/ldpc_b200, else ~/.cache/ldpc_b200; LDPC_B200_CACHE=0 disables it),
//     and verified against the graph when it is read back (a relabelling that does not match is ignored);
//   * otherwise the search runs on a background thread while the GPU decodes on the NATURAL layout (same results by
//     construction: the channel generator is keyed by original variable ids); launches pick the placed layout up as
//     soon as it is there -- between the frame chunks of a staged run, or at the next call.
// LDPC_B200_PLACE_SYNC=1 restores the blocking behaviour (tests that measure the placed kernels).
// `frames` = frames the caller is about to decode (all points): only a very long run pays for a longer placement
// search by itself (n = 1000: effort 24 takes 1.9 s on the host instead of 0.5 s and removes about a fifth of the
// remaining conflict replays, ~1 % of the run time); callers that reuse a handle set it with ldpc_b200_set_place_effort.
static uint64_t fnv1a64(const void* data, size_t bytes, uint64_t h) {   // 8 bytes per step: this hashes megabytes
    const unsigned char* p = (const unsigned char*)data;
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) { uint64_t w; memcpy(&w, p + i, 8); h ^= w; h *= 1099511628211ull; h ^= h >> 29; }
    for (; i < bytes; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

static uint64_t graph_content_hash(ldpc_b200_graph* g) {
    if (!g->graph_hash) {
        uint64_t h = 1469598103934665603ull;
        const int dims[4] = {g->n, g->k, g->dv, g->dc};
        h = fnv1a64(dims, sizeof dims, h);
        h = fnv1a64(g->v2c.data(), g->v2c.size() * sizeof(int), h);
        h = fnv1a64(g->c2v.data(), g->c2v.size() * sizeof(int), h);
        g->graph_hash = h ? h : 1;
    }
    return g->graph_hash;
}

static std::string place_cache_path(ldpc_b200_graph* g, int words, int effort) {
    const char* off = getenv("LDPC_B200_CACHE");
    if (off && atoi(off) == 0 && off[0] == '0') return "";
    std::string dir;
    if (const char* d = getenv("LDPC_B200_CACHE_DIR")) dir = d;
    else if (const char* x = getenv("XDG_CACHE_HOME")) dir = std::string(x) + "/ldpc_b200";
    else if (const char* h = getenv("HOME")) dir = std::string(h) + "/.cache/ldpc_b200";
    else return "";
    char name[128];
    snprintf(name, sizeof name, "/place_%016llx_w%d_e%d.bin", (unsigned long long)graph_content_hash(g), words, effort);
    return dir + name;
}

static void mkdirs(const std::string& dir) {
    for (size_t i = 1; i <= dir.size(); i++)
        if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0755);
}

struct PlaceFileHeader { uint32_t magic, version; int32_t n, m, dv, dc, B, effort; uint64_t hash; double wf[3]; };
static constexpr uint32_t kPlaceMagic = 0x4C50424Cu;   // "LBPL"

static void place_cache_store(ldpc_b200_graph* g, int words, int effort, const Placement& P) {
    const std::string path = place_cache_path(g, words, effort);
    if (path.empty()) return;
    mkdirs(path.substr(0, path.rfind('/')));
    const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    PlaceFileHeader H = {kPlaceMagic, 1u, g->n, g->m, g->dv, g->dc, 32 / words, effort, graph_content_hash(g),
                         {P.wavefronts_natural, P.wavefronts_placed, P.wavefronts_ideal}};
    bool ok = fwrite(&H, sizeof H, 1, f) == 1;
    auto put = [&](const std::vector<int>& v) { ok = ok && fwrite(v.data(), sizeof(int), v.size(), f) == v.size(); };
    put(P.var_slot); put(P.chk_slot); put(P.c2v); put(P.v2c); put(P.var_orig);
    ok = (fclose(f) == 0) && ok;
    if (ok) rename(tmp.c_str(), path.c_str()); else remove(tmp.c_str());
}

// reads a cached placement and VERIFIES that it is a relabelling of this graph (nothing on disk is trusted)
static bool place_cache_load(ldpc_b200_graph* g, int words, int effort, Placement& P) {
    const std::string path = place_cache_path(g, words, effort);
    if (path.empty()) return false;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    PlaceFileHeader H;
    bool ok = fread(&H, sizeof H, 1, f) == 1 && H.magic == kPlaceMagic && H.version == 1u && H.n == g->n && H.m == g->m &&
              H.dv == g->dv && H.dc == g->dc && H.B == 32 / words && H.effort == effort && H.hash == graph_content_hash(g);
    const size_t n = (size_t)g->n, m = (size_t)g->m, E = n * g->dv;
    auto get = [&](std::vector<int>& v, size_t cnt) { v.resize(cnt); ok = ok && fread(v.data(), sizeof(int), cnt, f) == cnt; };
    if (ok) { get(P.var_slot, n); get(P.chk_slot, m); get(P.c2v, E); get(P.v2c, E); get(P.var_orig, n); }
    fclose(f);
    if (!ok) return false;
    // var_slot / chk_slot are permutations, var_orig inverts var_slot, and every node's neighbour multiset maps over
    std::vector<char> seen(std::max(n, m), 0);
    for (size_t i = 0; i < n; i++) { const int s = P.var_slot[i]; if (s < 0 || (size_t)s >= n || seen[s] || P.var_orig[s] != (int)i) return false; seen[s] = 1; }
    std::fill(seen.begin(), seen.end(), 0);
    for (size_t c = 0; c < m; c++) { const int s = P.chk_slot[c]; if (s < 0 || (size_t)s >= m || seen[s]) return false; seen[s] = 1; }
    std::vector<int> a, b;
    for (size_t c = 0; c < m; c++) {
        a.clear(); b.clear();
        for (int j = 0; j < g->dc; j++) { a.push_back(P.var_slot[g->c2v[c * g->dc + j]]); b.push_back(P.c2v[(size_t)P.chk_slot[c] * g->dc + j]); }
        std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
        if (a != b) return false;
    }
    for (size_t i = 0; i < n; i++) {
        a.clear(); b.clear();
        for (int j = 0; j < g->dv; j++) { a.push_back(P.chk_slot[g->v2c[i * g->dv + j]]); b.push_back(P.v2c[(size_t)P.var_slot[i] * g->dv + j]); }
        std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
        if (a != b) return false;
    }
    P.n = g->n; P.m = g->m; P.dv = g->dv; P.dc = g->dc; P.B = 32 / words;
    P.wavefronts_natural = H.wf[0]; P.wavefronts_placed = H.wf[1]; P.wavefronts_ideal = H.wf[2];
    return true;
}

static int placement_upload(ldpc_b200_graph* g, int words, int effort, const Placement& P) {
    ldpc_b200_graph::Placed& pl = g->placed[words];
    std::vector<uint16_t> h_c((size_t)g->dc * g->m_pad, 0), h_v((size_t)g->dv * g->n_pad, 0), h_o((size_t)g->n_pad, 0);
    for (int c = 0; c < g->m; c++) for (int j = 0; j < g->dc; j++) h_c[sock_off(g->dc, g->m_pad, j, c)] = (uint16_t)P.c2v[(size_t)c * g->dc + j];
    for (int i = 0; i < g->n; i++) for (int j = 0; j < g->dv; j++) h_v[sock_off(g->dv, g->n_pad, j, i)] = (uint16_t)P.v2c[(size_t)i * g->dv + j];
    for (int i = 0; i < g->n; i++) h_o[i] = (uint16_t)P.var_orig[i];
    const size_t bc = align_up((uint32_t)(h_c.size() * 2), 16), bv = align_up((uint32_t)(h_v.size() * 2), 16), bo = align_up((uint32_t)(h_o.size() * 2), 16);
    // launches already in flight may still read the previous arrays: the old ones are released only after the device is idle
    uint16_t *nc = nullptr, *nv = nullptr, *no = nullptr;
    CUDA_TRY(cudaMalloc(&nc, bc)); CUDA_TRY(cudaMalloc(&nv, bv)); CUDA_TRY(cudaMalloc(&no, bo));
    CUDA_TRY(cudaMemset(nc, 0, bc)); CUDA_TRY(cudaMemset(nv, 0, bv));
    CUDA_TRY(cudaMemcpy(nc, h_c.data(), h_c.size() * 2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(nv, h_v.data(), h_v.size() * 2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(no, h_o.data(), h_o.size() * 2, cudaMemcpyHostToDevice));
    if (pl.ready) { CUDA_TRY(cudaDeviceSynchronize()); cudaFree(pl.d_c2v); cudaFree(pl.d_v2c); cudaFree(pl.d_var_orig); }
    pl.d_c2v = nc; pl.d_v2c = nv; pl.d_var_orig = no;
    pl.wf_natural = P.wavefronts_natural; pl.wf_placed = P.wavefronts_placed; pl.wf_ideal = P.wavefronts_ideal;
    pl.ready = true; pl.effort = effort;
    return LDPC_B200_OK;
}

static int placement_effort(const ldpc_b200_graph* g, double frames) {
    int effort = g->place_effort;
    if (effort < 0) {
        const char* s = getenv("LDPC_B200_PLACE_EFFORT");
        if (s) effort = atoi(s);
        else {
            effort = 6;
            if (frames * (double)g->n >= 2e13) effort = std::max(6, std::min(24, 24000 / std::max(1, g->n)));
        }
    }
    return effort;
}

// Makes progress towards a placement of at least the wanted effort without blocking (unless wait is set): takes a
// finished background search or a cached file, or starts the search.  Afterwards g->placed[words].ready says whether
// a placed layout (of any effort) is on the device; callers fall back to the natural layout when it is not.
static int ensure_placement(ldpc_b200_graph* g, int words, double frames, bool wait) {
    ldpc_b200_graph::Placed& pl = g->placed[words];
    const int effort = placement_effort(g, frames);
    if (tuning_env("LDPC_B200_PLACE_SYNC") == 1) wait = true;
    for (;;) {
        if (pl.ready && pl.effort >= effort && pl.state.load() == 0) return LDPC_B200_OK;
        const int st = pl.state.load(std::memory_order_acquire);
        if (st == 2) {   // a search has finished: upload its result
            if (pl.worker.joinable()) pl.worker.join();
            std::unique_ptr<Placement> R = std::move(pl.result);
            pl.state.store(0);
            if (R && (!pl.ready || pl.search_effort > pl.effort)) {
                const int rc = placement_upload(g, words, pl.search_effort, *R);
                if (rc) return rc;
                pl.source = 1;
            }
            continue;
        }
        if (st == 1) {   // searching
            if (!wait) return LDPC_B200_OK;
            if (pl.worker.joinable()) pl.worker.join();
            continue;
        }
        // idle and nothing good enough yet: disk cache first, else start a search
        if (effort <= 0) {   // effort 0 = natural layout, "placed" trivially
            Placement P;
            place_graph(g->n, g->m, g->dv, g->dc, g->v2c.data(), g->c2v.data(), 32 / words, 0, 0x5EEDull, P);
            return placement_upload(g, words, 0, P);
        }
        {
            Placement P;
            if (place_cache_load(g, words, effort, P)) { const int rc = placement_upload(g, words, effort, P); if (rc) return rc; pl.source = 2; continue; }
        }
        graph_content_hash(g);   // computed on this thread: the worker only reads it
        // Progressive: the FIRST search of a graph is a quick one (a quarter of effort 1 = 1/24 of the default's time, yet
        // most of the gain: the natural layout costs 3.1x the conflict-free wavefront count, effort 1 1.17x, effort 6
        // 1.11x), so that a cold run leaves the natural layout early; the next call then finds a placement of lower effort
        // than wanted and starts the full search behind it.
        const int search_effort = (!pl.ready && effort > 1 && tuning_env("LDPC_B200_PLACE_PROGRESSIVE") != 2) ? 1 : effort;
        pl.search_effort = search_effort;
        pl.state.store(1);
        try {
            pl.worker = std::thread([g, words, search_effort, effort, &pl]() {
                std::unique_ptr<Placement> R(new Placement());
                place_graph(g->n, g->m, g->dv, g->dc, g->v2c.data(), g->c2v.data(), 32 / words, search_effort, 0x5EEDull, *R,
                            search_effort == 1 && search_effort != effort ? 0.25 : 1.0);
                if (search_effort > 1) place_cache_store(g, words, search_effort, *R);   // the quick first placement is not worth a file
                pl.result = std::move(R);
                pl.state.store(2, std::memory_order_release);
            });
        } catch (...) {   // no thread to be had: search here
            pl.state.store(0);
            Placement P;
            place_graph(g->n, g->m, g->dv, g->dc, g->v2c.data(), g->c2v.data(), 32 / words, effort, 0x5EEDull, P);
            place_cache_store(g, words, effort, P);
            pl.source = 1;
            return placement_upload(g, words, effort, P);
        }
        if (!wait) return LDPC_B200_OK;
    }
}

// Blocks until no placement search of this handle is running and every finished one is on the device.  source (may be
// NULL) receives, for `words` (1, 2 or 4; 0 = skip): 0 = no placement, 1 = searched in this process, 2 = disk cache.
extern "C" int ldpc_b200_placement_wait(ldpc_b200_graph* g, int words, int* source) {
    if (!g || words < 0 || words > 4) return fail(LDPC_B200_EINVAL, "placement_wait: bad argument");
    return guarded("placement_wait", [&] {
        CUDA_TRY(cudaSetDevice(g->device));
        for (int w = 1; w <= 4; w++)   // every word width that was asked for: drive it to the full effort
            if (g->placed[w].state.load() != 0 || g->placed[w].ready) { const int rc = ensure_placement(g, w, 0.0, true); if (rc) return rc; }
        if (source) *source = words ? (g->placed[words].ready ? g->placed[words].source : 0) : 0;
        return (int)LDPC_B200_OK;
    });
}

// kernel parameters of the layout to decode on right now: the placed one if it is on the device, else the natural one
static void use_layout(const ldpc_b200_graph* g, int words, E1Params& P) {
    const ldpc_b200_graph::Placed& pl = g->placed[words];
    if (pl.ready) { P.c2v = pl.d_c2v; P.v2c = pl.d_v2c; P.var_orig = pl.d_var_orig; }
    else { P.c2v = g->d_c2v16; P.v2c = g->d_v2c16; P.var_orig = nullptr; }
}

// Scratch device buffers of the handle-less entry points.  cudaMalloc / cudaFree cost milliseconds (and now and
// then tens of milliseconds) per call, so released buffers go back to a small per-device pool instead.
struct DevPool {
    std::mutex mu;
    struct Item { int dev; size_t bytes; void* p; };
    std::vector<Item> free_list;
    size_t pooled = 0;
    static constexpr size_t kMaxPooled = (size_t)1 << 30;
    cudaError_t get(size_t bytes, void** out, size_t* got) {
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> lk(mu);
            int best = -1;
            for (int i = 0; i < (int)free_list.size(); i++)
                if (free_list[i].dev == dev && free_list[i].bytes >= bytes && free_list[i].bytes <= 2 * bytes + 4096 &&
                    (best < 0 || free_list[i].bytes < free_list[best].bytes)) best = i;
            if (best >= 0) {
                *out = free_list[best].p; *got = free_list[best].bytes;
                pooled -= free_list[best].bytes;
                free_list.erase(free_list.begin() + best);
                return cudaSuccess;
            }
        }
        *got = bytes;
        return cudaMalloc(out, bytes);
    }
    void put(void* p, size_t bytes) {
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> lk(mu);
            if (pooled + bytes <= kMaxPooled && free_list.size() < 64) {
                free_list.push_back({dev, bytes, p});
                pooled += bytes;
                return;
            }
        }
        cudaFree(p);
    }
};
static DevPool g_dev_pool;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    // the device is idle before a buffer goes back to the pool: on an error path work on the caller's stream may still
    // be using it (on the normal paths the callers have synchronised already, so this costs nothing)
    ~DevBuf() { if (p) { cudaDeviceSynchronize(); g_dev_pool.put(p, bytes); } }
    cudaError_t alloc(size_t want) { return g_dev_pool.get(want ? want : 1, &p, &bytes); }
};

#include "e2_host.inc"
#include "e1_staged_host.inc"

static int ldpc_b200_mc_launch_impl(ldpc_b200_graph* g, const double* eps, int npoints, uint32_t first_point, uint64_t seed,
                                   uint64_t first_frame, uint64_t num_frames, int iterations, int expurgation,
                                   uint64_t* d_counters, void* stream) {
    if (!g || !eps || !d_counters) return fail(LDPC_B200_EINVAL, "mc_launch: NULL argument");
    if (npoints <= 0 || npoints >= (1 << 24) || (unsigned long long)first_point + (unsigned)npoints > (1u << 24))
        return fail(LDPC_B200_EINVAL, "mc_launch: npoints / first_point out of range (the generator keys 2^24 points)");
    if (iterations < 0 || iterations > 65535) return fail(LDPC_B200_EINVAL, "mc_launch: iterations out of range");
    if (first_frame % 128 != 0) return fail(LDPC_B200_EINVAL, "mc_launch: first_frame must be a multiple of 128");
    cudaStream_t st = (cudaStream_t)stream;
    if (tuning_env("LDPC_B200_TRACE")) fprintf(stderr, "[ldpc_b200] %.3f s: mc_launch enter (%llu frames/point)\n", host_seconds(), (unsigned long long)num_frames);
    CUDA_TRY(cudaSetDevice(g->device));
    if (num_frames == 0) return LDPC_B200_OK;

    E1Config cfg;
    const int weng = want_engine(g);
    const bool fits_e1 = plan_e1(g, iterations, false, false, cfg);
    if (weng == 1 && !fits_e1)
        return fail(LDPC_B200_EUNSUPPORTED, "mc_launch: graph (n=%d dv=%d dc=%d) does not fit the shared-memory engine", g->n, g->dv, g->dc);
    const bool use_e2 = weng == 2 || !fits_e1;

    const int cap = iterations;
    // One W = 128 team owns the SM but leaves no room for a staged pattern (n = 6000 .. 9000): a W = 64 team with the
    // channel draw overlapped (e1_overlap.cuh) is faster (n = 8000: 4.95 -> 5.23e13 edge updates/s, n = 6000: +2 %) --
    // with the socket lists in tensor memory the word width no longer changes the shared-memory traffic per frame.
    if (!use_e2 && cfg.tm && cfg.words == 4 && cfg.teams == 1 && num_frames >= 8192 && !g->tune_words && !tuning_env("LDPC_B200_WORDS")) {
        E1Config ov, alt;
        if (!plan_e1_overlap(g, cap, cfg, ov) && plan_e1_tm(g, cap, alt, 2) && alt.teams == 1 && plan_e1_overlap(g, cap, alt, ov)) cfg = alt;
    }
    const size_t raw_words = (size_t)npoints * raw_stride(cap);
    if (g->raw_cap < raw_words) {
        cudaFree(g->d_raw); g->d_raw = nullptr; g->raw_cap = 0;
        CUDA_TRY(cudaMalloc(&g->d_raw, raw_words * 8));
        g->raw_cap = raw_words;
    }
    if (g->bec_cap < npoints) {
        cudaFree(g->d_bec); g->d_bec = nullptr; g->bec_cap = 0;
        CUDA_TRY(cudaMalloc(&g->d_bec, sizeof(BecParams) * (size_t)npoints));
        g->bec_cap = npoints;
    }
    {
        const int slot = g->h_bec_next;
        g->h_bec_next = (slot + 1) % ldpc_b200_graph::BEC_RING;
        if (!g->h_bec_ev[slot]) CUDA_TRY(cudaEventCreateWithFlags(&g->h_bec_ev[slot], cudaEventDisableTiming));
        else CUDA_TRY(cudaEventSynchronize(g->h_bec_ev[slot]));   // the copy that last used this slot (8 launches ago)
        if (g->h_bec_cap[slot] < npoints) {
            if (g->h_bec[slot]) cudaFreeHost(g->h_bec[slot]);
            g->h_bec[slot] = nullptr; g->h_bec_cap[slot] = 0;
            CUDA_TRY(cudaMallocHost(&g->h_bec[slot], sizeof(BecParams) * (size_t)npoints));
            g->h_bec_cap[slot] = npoints;
        }
        for (int p = 0; p < npoints; p++) g->h_bec[slot][p] = make_bec(eps[p], seed);
        CUDA_TRY(cudaMemcpyAsync(g->d_bec, g->h_bec[slot], sizeof(BecParams) * (size_t)npoints, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(g->h_bec_ev[slot], st));
    }
    CUDA_TRY(cudaMemsetAsync(g->d_raw, 0, raw_words * 8, st));
    CUDA_TRY(cudaMemsetAsync(g->d_work, 0, sizeof(unsigned long long), st));

    if (use_e2) {
        int rc2 = ensure_events(g);
        if (rc2) return rc2;
        const int slot2 = g->ev_count < ldpc_b200_graph::RING ? g->ev_count : -1;
        if (slot2 >= 0) CUDA_TRY(cudaEventRecord(g->ev0[slot2], st));
        int grid2 = 0, R2 = 0;
        rc2 = e2_mc(g, npoints, first_point, first_frame, num_frames, cap, expurgation, st, &grid2, &R2);
        if (rc2) return rc2;
        if (slot2 >= 0) { CUDA_TRY(cudaEventRecord(g->ev1[slot2], st)); g->ev_count++; }
        e1_finalize_kernel<<<(npoints + 127) / 128, 128, 0, st>>>(g->d_raw, npoints, cap);
        mc_compact_kernel<<<npoints, 128, 0, st>>>(g->d_raw, (unsigned long long*)d_counters, g->d_stats, npoints, cap);
        CUDA_TRY(cudaGetLastError());
        g->launches += 3; g->last_stages = 1;
        g->last.engine = 2; g->last.words = R2; g->last.teams = 1; g->last.team_warps = E2_THREADS / 32;
        g->last.grid = grid2; g->last.smem_bytes = 0;
        return LDPC_B200_OK;
    }
    {
        int rcp = ensure_placement(g, cfg.words, (double)num_frames * npoints, false);
        if (rcp) return rcp;
    }
    E1Params P;
    fill_params(g, cfg, cap, P);
    use_layout(g, cfg.words, P);
    const unsigned long long W = 32ull * cfg.words;
    P.bec = g->d_bec; P.seed_lo = (uint32_t)seed; P.seed_hi = (uint32_t)(seed >> 32); P.npoints = npoints; P.expurgation = expurgation;
    philox_round_keys(P.seed_lo, P.seed_hi, P.rk);
    P.point0 = first_point;
    P.first_word = first_frame / 32; P.num_frames = num_frames;
    P.groups_per_point = (num_frames + W - 1) / W;
    P.counters = g->d_raw; P.work_counter = g->d_work;
    const unsigned long long items = P.groups_per_point * (unsigned long long)npoints;
    const unsigned long long ctas_needed = (items + cfg.teams - 1) / cfg.teams;
    const int grid = (int)std::min<unsigned long long>((unsigned long long)g->sm_count, ctas_needed);
    const int block = cfg.teams * cfg.team_warps * 32;

    int rc = ensure_events(g);
    if (rc) return rc;
    // staged path: survivors are re-packed between stages (e1_staged.cuh); exact, used when every frame
    // counts (no expurgation mask) and the run is long enough to amortise the pilot
    // Expurgated runs (expurgation >= 0) take the same pipeline as a listing pass + a re-count of the listed frames
    // (e1_staged.cuh) unless the pilot saw a block error rate at which re-drawing the failed frames one by one
    // would cost more than decoding mixed groups twice (the unstaged kernel below).
    StagePlan* plan = nullptr;
    bool two_pass = false;
    if (num_frames >= 8192 && tuning_env("LDPC_B200_STAGES") != 1 && (expurgation < 0 || tuning_env("LDPC_B200_EXP2PASS") != 2)) {
        rc = get_stage_plan(g, cfg, P, eps, npoints, seed, cap, num_frames, st, &plan);
        if (rc) return rc;
        two_pass = expurgation >= 0 && (plan->pilot_fer <= 0.10 || tuning_env("LDPC_B200_EXP2PASS") == 1);
        if (expurgation >= 0 && !two_pass) plan = nullptr;
        else if (plan->nstages <= 1 && !two_pass) plan = nullptr;
        // the pilot (first call per plan key) pulls its groups from the same work counter: start from zero again
        CUDA_TRY(cudaMemsetAsync(g->d_work, 0, sizeof(unsigned long long), st));
    }
    const int slot = g->ev_count < ldpc_b200_graph::RING ? g->ev_count : -1;
    if (slot >= 0) CUDA_TRY(cudaEventRecord(g->ev0[slot], st));
    if (tuning_env("LDPC_B200_TRACE")) fprintf(stderr, "[ldpc_b200] %.3f s: plan ready (pilot done), launching\n", host_seconds());
    int grid_used = grid;
    g->last.overlap_producers = 0;
    if (plan) {
        rc = e1_mc_staged(g, cfg, P, *plan, npoints, first_frame, num_frames, two_pass, expurgation, st, &grid_used);
        if (rc) return rc;
        g->last_stages = plan->nstages;
    } else {
        cudaError_t e;
        if (cfg.words == 4) e = launch_mc_words<4>(P, grid, block, cfg.smem, st);
        else if (cfg.words == 2) e = launch_mc_words<2>(P, grid, block, cfg.smem, st);
        else e = launch_mc_words<1>(P, grid, block, cfg.smem, st);
        if (e != cudaSuccess) return fail(LDPC_B200_ECUDA, "mc_launch: kernel launch failed: %s", cudaGetErrorString(e));
        g->launches += 1;
        g->last_stages = 1;
    }
    if (slot >= 0) { CUDA_TRY(cudaEventRecord(g->ev1[slot], st)); g->ev_count++; }
    e1_finalize_kernel<<<(npoints + 127) / 128, 128, 0, st>>>(g->d_raw, npoints, cap);
    mc_compact_kernel<<<npoints, 128, 0, st>>>(g->d_raw, (unsigned long long*)d_counters, g->d_stats, npoints, cap);
    CUDA_TRY(cudaGetLastError());

    g->launches += 2;
    g->last.engine = 1; g->last.words = cfg.words; g->last.teams = cfg.teams; g->last.team_warps = cfg.team_warps;
    g->last.grid = grid_used; g->last.smem_bytes = (int)cfg.smem; g->last.tmem_lists = cfg.tm ? 1 : 0;
    g->last.record_bytes = plan ? 4 * (int)align_up((uint32_t)((g->n + 31) / 32) + (two_pass ? 1u : 0u), 4) : 0;
    return LDPC_B200_OK;
}
extern "C" int ldpc_b200_mc_launch(ldpc_b200_graph* g, const double* eps, int npoints, uint64_t seed,
                                   uint64_t first_frame, uint64_t num_frames, int iterations, int expurgation,
                                   uint64_t* d_counters, void* stream) {
    return guarded("mc_launch", [&] { return ldpc_b200_mc_launch_impl(g, eps, npoints, 0u, seed, first_frame, num_frames, iterations, expurgation, d_counters, stream); });
}

// The same launch for the erasure points [first_point, first_point + npoints) of a longer list: the channel generator
// keys point p of this call as first_point + p, so a run may be split by POINT across GPUs or calls and
// still decode exactly the frames of the unsplit run.  d_counters: [npoints][iterations+6] rows of THESE points.
extern "C" int ldpc_b200_mc_launch_points(ldpc_b200_graph* g, const double* eps, int npoints, uint32_t first_point, uint64_t seed,
                                          uint64_t first_frame, uint64_t num_frames, int iterations, int expurgation,
                                          uint64_t* d_counters, void* stream) {
    return guarded("mc_launch_points", [&] { return ldpc_b200_mc_launch_impl(g, eps, npoints, first_point, seed, first_frame, num_frames, iterations, expurgation, d_counters, stream); });
}


#include "e1_ens_host.inc"

extern "C" int ldpc_b200_mc_stats_read(ldpc_b200_graph* g, ldpc_b200_mc_stats* stats) {
    if (!g || !stats) return fail(LDPC_B200_EINVAL, "mc_stats_read: NULL argument");
    CUDA_TRY(cudaSetDevice(g->device));
    unsigned long long h[8];
    CUDA_TRY(cudaMemcpy(h, g->d_stats, sizeof h, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemset(g->d_stats, 0, sizeof h));
    *stats = g->last;
    for (int i = 0; i < 4; i++) stats->overlap_cycles[i] = h[4 + i];
    stats->phys_frame_iterations = h[0];
    stats->second_pass_groups = h[1];
    stats->spilled_records = h[2];
    stats->recounted_frames = h[3];
    stats->stages = g->last_stages;
    stats->kernel_launches = g->launches;
    g->launches = 0;
    return LDPC_B200_OK;
}

extern "C" int ldpc_b200_kernel_times(ldpc_b200_graph* g, float* ms, int max) {
    if (!g || !ms) return fail(LDPC_B200_EINVAL, "kernel_times: NULL argument");
    CUDA_TRY(cudaSetDevice(g->device));
    int n = std::min(g->ev_count, max);
    for (int i = 0; i < n; i++) {
        CUDA_TRY(cudaEventSynchronize(g->ev1[i]));
        CUDA_TRY(cudaEventElapsedTime(&ms[i], g->ev0[i], g->ev1[i]));
    }
    g->ev_count = 0;
    return n;
}

static int ldpc_b200_mc_run_impl(ldpc_b200_graph* g, const double* eps, int npoints, uint64_t seed,
                                uint64_t first_frame, uint64_t num_frames, int iterations, int expurgation,
                                uint64_t* counters, ldpc_b200_mc_stats* stats) {
    if (!g || !eps || !counters) return fail(LDPC_B200_EINVAL, "mc_run: NULL argument");
    if (npoints <= 0) return fail(LDPC_B200_EINVAL, "mc_run: npoints <= 0");
    if (iterations < 0) return fail(LDPC_B200_EINVAL, "mc_run: iterations < 0");
    CUDA_TRY(cudaSetDevice(g->device));
    const size_t words = (size_t)npoints * (iterations + 6);
    if (g->out_cap < words) {
        cudaFree(g->d_out); g->d_out = nullptr; g->out_cap = 0;
        CUDA_TRY(cudaMalloc(&g->d_out, words * 8));
        g->out_cap = words;
    }
    CUDA_TRY(cudaMemsetAsync(g->d_out, 0, words * 8, 0));
    int rc = ldpc_b200_mc_launch(g, eps, npoints, seed, first_frame, num_frames, iterations, expurgation,
                                 (uint64_t*)g->d_out, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(counters, g->d_out, words * 8, cudaMemcpyDeviceToHost));
    if (stats) return ldpc_b200_mc_stats_read(g, stats);
    return LDPC_B200_OK;
}
extern "C" int ldpc_b200_mc_run(ldpc_b200_graph* g, const double* eps, int npoints, uint64_t seed,
                                uint64_t first_frame, uint64_t num_frames, int iterations, int expurgation,
                                uint64_t* counters, ldpc_b200_mc_stats* stats) {
    return guarded("mc_run", [&] { return ldpc_b200_mc_run_impl(g, eps, npoints, seed, first_frame, num_frames, iterations, expurgation, counters, stats); });
}


// ------------------------------------------------------------------------------------------------
template <bool VALUES>
static cudaError_t launch_decode(const E1Params& P, int grid, int block, uint32_t smem, cudaStream_t st) {
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, block, smem, st>>>(P);
        return cudaGetLastError();
    };
    if (P.dv == 3 && P.dc == 6) return go(e1_decode_kernel<3, 6, VALUES>);
    if (P.dv == 4 && P.dc == 8) return go(e1_decode_kernel<4, 8, VALUES>);
    return go(e1_decode_kernel<0, 0, VALUES>);
}

// flags |= 1 if some symbol == 1 (value plane needed), |= 2 if some symbol is outside {0,1,2}
__global__ void scan_symbols_kernel(const int8_t* __restrict__ mvc, size_t total, unsigned int* flags) {
    unsigned int f = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int8_t s = mvc[i];
        f |= (s == 1) ? 1u : 0u;
        f |= (s < 0 || s > 2) ? 2u : 0u;
    }
    f = __reduce_or_sync(0xFFFFFFFFu, f);
    if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

static int grow(void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return LDPC_B200_OK;
    cudaFree(*p); *p = nullptr; *cap = 0;
    CUDA_TRY(cudaMalloc(p, bytes));
    *cap = bytes;
    return LDPC_B200_OK;
}

// Decode of at most 32 frames (one frame group) in ONE kernel launch: symbols go up and results come down through one
// pinned block each way, nothing else synchronises.  This is what a message_passing() call costs on the GPU.
static int decode_small(ldpc_b200_graph* g, const E1Config& cfg, bool values, int8_t* Mvc, int frames, int iterations,
                        int* errors, int* ret) {
    const size_t nsym = (size_t)frames * g->n;
    const size_t off_err = (nsym + 15) / 16 * 16, off_ret = off_err + (size_t)frames * std::max(iterations, 1) * 4;
    const size_t total = off_ret + (size_t)frames * 4;
    if (g->fast_cap < total) {
        if (g->fast_pin) cudaFreeHost(g->fast_pin);
        cudaFree(g->fast_dev);
        g->fast_pin = g->fast_dev = nullptr; g->fast_cap = 0;
        const size_t cap_bytes = std::max<size_t>(total, 65536);
        CUDA_TRY(cudaMallocHost(&g->fast_pin, cap_bytes));
        CUDA_TRY(cudaMalloc(&g->fast_dev, cap_bytes));
        g->fast_cap = cap_bytes;
    }
    unsigned char* hp = (unsigned char*)g->fast_pin;
    unsigned char* dp = (unsigned char*)g->fast_dev;
    memcpy(hp, Mvc, nsym);
    CUDA_TRY(cudaMemcpyAsync(dp, hp, nsym, cudaMemcpyHostToDevice, 0));
    E1Config one = cfg;   // ONE group: one team of all the warps (same shared-memory map, the other teams' regions stay unused)
    one.teams = 1; one.team_warps = E1_MAX_THREADS / 32;
    E1Params P;
    fill_params(g, one, iterations, P);
    P.num_frames = (unsigned long long)frames; P.groups_per_point = 1;
    P.mvc8 = (const int8_t*)dp;
    P.errors = (int*)(dp + off_err); P.ret = (int*)(dp + off_ret);
    const int block = one.teams * one.team_warps * 32;
    cudaError_t e = values ? launch_decode<true>(P, 1, block, cfg.smem, 0) : launch_decode<false>(P, 1, block, cfg.smem, 0);
    if (e != cudaSuccess) return fail(LDPC_B200_ECUDA, "decode: kernel launch failed: %s", cudaGetErrorString(e));
    CUDA_TRY(cudaMemcpyAsync(hp, dp, total, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    memcpy(Mvc, hp, nsym);
    if (iterations > 0) memcpy(errors, hp + off_err, (size_t)frames * iterations * 4);
    memcpy(ret, hp + off_ret, (size_t)frames * 4);
    g->launches += 1;
    return LDPC_B200_OK;
}

static int ldpc_b200_decode_batch_impl(ldpc_b200_graph* g, int8_t* Mvc, uint64_t frames, int iterations,
                                      int* errors, int* ret) {
    if (!g || !Mvc || !ret || (!errors && iterations > 0)) return fail(LDPC_B200_EINVAL, "decode_batch: NULL argument");
    if (iterations < 0) return fail(LDPC_B200_EINVAL, "decode_batch: iterations < 0");
    if (frames == 0) return LDPC_B200_OK;
    std::lock_guard<std::mutex> handle_lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    if (frames <= 32 && want_engine(g) != 2 && iterations > 0) {
        // one frame group: symbols are checked on the host (a few KB), then one launch does everything
        const size_t nsym = (size_t)frames * g->n;
        bool values = false;
        for (size_t i = 0; i < nsym; i++) {
            const int8_t sym = Mvc[i];
            if (sym < 0 || sym > 2)
                return fail(LDPC_B200_EINVAL, "decode_batch: symbol %d at frame %zu position %zu is outside {0,1,2}", (int)sym, i / g->n, i % g->n);
            values |= (sym == 1);
        }
        E1Config cfg;
        if (plan_e1(g, iterations, values, true, cfg)) return decode_small(g, cfg, values, Mvc, (int)frames, iterations, errors, ret);
    }
    // chunks of at most 2^20 frames (device buffers are cached in the handle and reused)
    const uint64_t chunk_frames = std::min<uint64_t>(frames, std::max<uint64_t>(1024, std::min<uint64_t>(1ull << 20, (1ull << 31) / (uint64_t)std::max(g->n, 1))));
    const size_t cb = (size_t)chunk_frames * g->n;
    const uint64_t cgroups = (chunk_frames + 31) / 32;
    int rc;
    if ((rc = grow(&g->dec_mvc, &g->dec_mvc_cap, cb))) return rc;
    if ((rc = grow(&g->dec_x, &g->dec_x_cap, cgroups * g->n * 4))) return rc;
    if ((rc = grow(&g->dec_v, &g->dec_v_cap, cgroups * g->n * 4))) return rc;
    if ((rc = grow(&g->dec_err, &g->dec_err_cap, (size_t)chunk_frames * std::max(iterations, 1) * 4))) return rc;
    if ((rc = grow(&g->dec_ret, &g->dec_ret_cap, (size_t)chunk_frames * 4 + 16))) return rc;
    unsigned int* d_flags = (unsigned int*)((char*)g->dec_ret + (size_t)chunk_frames * 4);
    const int weng = want_engine(g);

    for (uint64_t f0 = 0; f0 < frames; f0 += chunk_frames) {
        const uint64_t fc = std::min<uint64_t>(chunk_frames, frames - f0);
        const size_t total = (size_t)fc * g->n;
        int8_t* hM = Mvc + (size_t)f0 * g->n;
        CUDA_TRY(cudaMemcpy(g->dec_mvc, hM, total, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemsetAsync(d_flags, 0, 4, 0));
        scan_symbols_kernel<<<g->sm_count * 8, 256>>>((const int8_t*)g->dec_mvc, total, d_flags);
        unsigned int flags = 0;
        CUDA_TRY(cudaMemcpy(&flags, d_flags, 4, cudaMemcpyDeviceToHost));
        if (flags & 2u) {
            for (size_t i = 0; i < total; i++)
                if (hM[i] < 0 || hM[i] > 2)
                    return fail(LDPC_B200_EINVAL, "decode_batch: symbol %d at frame %zu position %zu is outside {0,1,2}",
                                (int)hM[i], (size_t)f0 + i / g->n, i % g->n);
        }
        const bool values = (flags & 1u) != 0;
        E1Config cfg;
        const bool fits_e1 = plan_e1(g, iterations, values, true, cfg);
        if (weng == 1 && !fits_e1)
            return fail(LDPC_B200_EUNSUPPORTED, "decode_batch: graph (n=%d dv=%d dc=%d, value plane %s) does not fit the shared-memory engine",
                        g->n, g->dv, g->dc, values ? "on" : "off");
        int* d_err = (int*)g->dec_err;
        int* d_ret = (int*)g->dec_ret;
        if (weng == 2 || !fits_e1) {
            rc = e2_decode(g, (int8_t*)g->dec_mvc, fc, iterations, values, d_err, d_ret);
            if (rc) return rc;
        } else {
            const uint64_t groups = (fc + 31) / 32;
            dim3 pgrid((g->n + 127) / 128, (unsigned)groups);
            pack_planes_kernel<<<pgrid, 128>>>((const int8_t*)g->dec_mvc, (uint32_t*)g->dec_x, (uint32_t*)g->dec_v, fc, g->n);
            CUDA_TRY(cudaGetLastError());
            E1Params P;
            fill_params(g, cfg, iterations, P);
            P.num_frames = fc; P.groups_per_point = groups;
            P.xin = (const uint32_t*)g->dec_x; P.vin = (const uint32_t*)g->dec_v;
            P.xout = (uint32_t*)g->dec_x; P.vout = (uint32_t*)g->dec_v;
            P.errors = d_err; P.ret = d_ret;
            const int grid = (int)std::min<uint64_t>((uint64_t)g->sm_count, (groups + cfg.teams - 1) / cfg.teams);
            const int block = cfg.teams * cfg.team_warps * 32;
            cudaError_t e = values ? launch_decode<true>(P, grid, block, cfg.smem, 0) : launch_decode<false>(P, grid, block, cfg.smem, 0);
            if (e != cudaSuccess) return fail(LDPC_B200_ECUDA, "decode_batch: kernel launch failed: %s", cudaGetErrorString(e));
            unpack_planes_kernel<<<pgrid, 128>>>((int8_t*)g->dec_mvc, (const uint32_t*)g->dec_x, (const uint32_t*)g->dec_v, fc, g->n);
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaMemcpy(hM, g->dec_mvc, total, cudaMemcpyDeviceToHost));
        if (iterations > 0) CUDA_TRY(cudaMemcpy(errors + (size_t)f0 * iterations, d_err, (size_t)fc * iterations * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(ret + f0, d_ret, (size_t)fc * 4, cudaMemcpyDeviceToHost));
    }
    return LDPC_B200_OK;
}
extern "C" int ldpc_b200_decode_batch(ldpc_b200_graph* g, int8_t* Mvc, uint64_t frames, int iterations,
                                      int* errors, int* ret) {
    return guarded("decode_batch", [&] { return ldpc_b200_decode_batch_impl(g, Mvc, frames, iterations, errors, ret); });
}


static int ldpc_b200_draw_patterns_impl(int device, int n, double eps, uint64_t seed, uint32_t point,
                                       uint64_t first_frame, uint64_t frames, int8_t* out) {
    if (!out || n <= 0) return fail(LDPC_B200_EINVAL, "draw_patterns: bad argument");
    if (first_frame % 32 != 0) return fail(LDPC_B200_EINVAL, "draw_patterns: first_frame must be a multiple of 32");
    if (frames == 0) return LDPC_B200_OK;
    CUDA_TRY(cudaSetDevice(device));
    const uint64_t words = (frames + 31) / 32;
    if (words > 65535) return fail(LDPC_B200_EUNSUPPORTED, "draw_patterns: more than 65535*32 frames per call");
    DevBuf d_x, d_v, d_m;
    CUDA_TRY(d_x.alloc(words * n * 4));
    CUDA_TRY(d_v.alloc(words * n * 4));
    CUDA_TRY(d_m.alloc(frames * n));
    CUDA_TRY(cudaMemset(d_v.p, 0, words * n * 4));
    dim3 grid((n + 127) / 128, (unsigned)words);
    draw_planes_kernel<<<grid, 128>>>((uint32_t*)d_x.p, make_bec(eps, seed), point, first_frame / 32, words, n);
    unpack_planes_kernel<<<grid, 128>>>((int8_t*)d_m.p, (const uint32_t*)d_x.p, (const uint32_t*)d_v.p, frames, n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, d_m.p, frames * n, cudaMemcpyDeviceToHost));
    return LDPC_B200_OK;
}
extern "C" int ldpc_b200_draw_patterns(int device, int n, double eps, uint64_t seed, uint32_t point,
                                       uint64_t first_frame, uint64_t frames, int8_t* out) {
    return guarded("draw_patterns", [&] { return ldpc_b200_draw_patterns_impl(device, n, eps, seed, point, first_frame, frames, out); });
}


// ------------------------------------------------------------------------------------------------
// Shared-memory bandwidth probe: the roofline denominator for the shared-memory-resident engine is
// measured on the box (conflict-free LDS.128 streams from every SM), not taken from a data sheet.
__global__ void __launch_bounds__(1024, 1) smem_probe_kernel(uint4* sink, int iters) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint4* s = reinterpret_cast<uint4*>(sm);
    const int words = 4096;  // 64 KB
    for (int i = threadIdx.x; i < words; i += blockDim.x) s[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    uint4 acc = make_uint4(0, 0, 0, 0);
    // exactly 16 LDS.128 per thread and iteration: volatile asm, so the compiler can neither merge two loads of the same
    // address nor carry one across iterations (a carried load once made this probe report more than the 128 B/clk/SM
    // the hardware has); lane l of a quarter warp reads bank quad l: conflict-free
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 1023u) * 16u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            uint4 v;
            const uint32_t a = base + (uint32_t)(((k * 256 + it * 64) & 3071)) * 16u;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    if (acc.x == 0x12345678u && acc.y == 1u) sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

extern "C" int ldpc_b200_smem_probe(int device, double* gbytes_per_s) {
    if (!gbytes_per_s) return fail(LDPC_B200_EINVAL, "smem_probe: NULL argument");
    CUDA_TRY(cudaSetDevice(device));
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    DevBuf sink;
    CUDA_TRY(sink.alloc((size_t)sms * 1024 * sizeof(uint4)));
    CUDA_TRY(cudaFuncSetAttribute(smem_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 4000;
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        smem_probe_kernel<<<sms, 1024, 65536>>>((uint4*)sink.p, iters);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double bytes = (double)sms * 1024.0 * iters * 16.0 * 16.0;
        if (rep >= 1) best = std::max(best, bytes / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *gbytes_per_s = best;
    return LDPC_B200_OK;
}

// ------------------------------------------------------------------------------------------------
// message_passing.c:7 drop-in.  Device copies of the graph are cached by content because the reference passes the
// same lookups for every frame in fixed-code mode (parallel_simulator.py:360) -- in freshly made numpy arrays, so the
// pointers say nothing.  A call finds its handle by a SAMPLED hash of the lookups (64 probes, ~100 ns) followed by a
// full memcmp against the handle's host copy (24 KB at n = 1000: ~1 us; a byte-wise hash of both arrays cost 50 us per
// call); at most 16 handles are kept, the least recently used one is dropped; the global lock covers the table only,
// GPU work runs under the handle's own lock.
static std::mutex g_cache_mutex;
struct CacheEntry { ldpc_b200_graph* g; uint64_t key; uint64_t tick; int users; };
static std::vector<CacheEntry> g_cache;
static uint64_t g_cache_tick = 0;
static constexpr size_t kCacheMax = 16;

static uint64_t sampled_key(const int* v2c, const int* c2v, size_t E, int n, int k, int dv, int dc) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { h ^= v; h *= 1099511628211ull; h ^= h >> 31; };
    mix((uint64_t)(uint32_t)n << 32 | (uint32_t)k); mix((uint64_t)(uint32_t)dv << 32 | (uint32_t)dc);
    const size_t step = std::max<size_t>(1, E / 32);
    for (size_t e = 0; e < E; e += step) { mix((uint32_t)v2c[e]); mix((uint32_t)c2v[e]); }
    mix((uint32_t)v2c[E - 1]); mix((uint32_t)c2v[E - 1]);
    return h;
}

static int message_passing_impl(int* Mvc, int iterations, int* variable_to_check_list, int* check_to_variable_list,
                               int* errors, int n, int k, int dv, int dc) {
    if (!Mvc || !variable_to_check_list || !check_to_variable_list || (!errors && iterations > 0))
        return fail(LDPC_B200_EINVAL, "message_passing: NULL argument");
    if (iterations <= 0) return 0;  // message_passing.c:15,81: loop does not run, returns it == 0
    if (n <= 0 || dv <= 0) return fail(LDPC_B200_EINVAL, "message_passing: bad sizes n=%d dv=%d", n, dv);
    const size_t E = (size_t)n * dv;
    const uint64_t key = sampled_key(variable_to_check_list, check_to_variable_list, E, n, k, dv, dc);
    ldpc_b200_graph* g = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        for (auto& ce : g_cache)
            if (ce.key == key && ce.g->n == n && ce.g->k == k && ce.g->dv == dv && ce.g->dc == dc &&
                memcmp(ce.g->v2c.data(), variable_to_check_list, E * sizeof(int)) == 0 &&
                memcmp(ce.g->c2v.data(), check_to_variable_list, E * sizeof(int)) == 0) {
                g = ce.g; ce.tick = ++g_cache_tick; ce.users++;
                break;
            }
    }
    bool cached = g != nullptr;
    if (!g) {
        const char* dev_s = getenv("LDPC_B200_DEVICE");
        int rc = ldpc_b200_graph_create(variable_to_check_list, check_to_variable_list, n, k, dv, dc,
                                        dev_s ? atoi(dev_s) : 0, &g);
        if (rc) return rc;
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        if (g_cache.size() >= kCacheMax) {   // drop the least recently used handle nobody is inside
            int victim = -1;
            for (int i = 0; i < (int)g_cache.size(); i++)
                if (g_cache[i].users == 0 && (victim < 0 || g_cache[i].tick < g_cache[victim].tick)) victim = i;
            if (victim >= 0) { ldpc_b200_graph_destroy(g_cache[victim].g); g_cache.erase(g_cache.begin() + victim); }
        }
        if (g_cache.size() < kCacheMax) { g_cache.push_back({g, key, ++g_cache_tick, 1}); cached = true; }
    }
    struct Guard {
        ldpc_b200_graph* g; bool keep;
        ~Guard() {
            if (!keep) { ldpc_b200_graph_destroy(g); return; }
            std::lock_guard<std::mutex> lock(g_cache_mutex);
            for (auto& ce : g_cache) if (ce.g == g) { ce.users--; break; }
        }
    } guard{g, cached};

    static thread_local std::vector<int8_t> frame, input;
    static thread_local std::vector<int> cnt;
    frame.resize(n); cnt.resize(iterations);
    for (int i = 0; i < n; i++) {
        if (Mvc[i] < 0 || Mvc[i] > 2) return fail(LDPC_B200_EINVAL, "message_passing: Mvc[%d]=%d outside {0,1,2}", i, Mvc[i]);
        frame[i] = (int8_t)Mvc[i];
    }
    bool prefilled = false;
    for (int it = 0; it < iterations; it++) prefilled |= (errors[it] != 0);
    if (prefilled) input = frame;
    int ret = 0;
    int rc = ldpc_b200_decode_batch(g, frame.data(), 1, iterations, cnt.data(), &ret);
    if (rc) return rc;

    if (!prefilled) {
        for (int it = 0; it < iterations; it++) errors[it] = cnt[it];
        for (int i = 0; i < n; i++) Mvc[i] = frame[i];
        return ret;
    }
    // Caller did not zero errors[] (the reference accumulates with +=, message_passing.c:73, and its
    // stall test :16 compares the ACCUMULATED values).  cnt[j] is the erased count after j+1 executed
    // flooding iterations; replay the reference's control flow over it.
    int executed = 0, it = 0;
    for (; it < iterations; it++) {
        if (it > 1 && errors[it - 1] == errors[it - 2]) { errors[it] = errors[it - 1]; continue; }
        const int c = cnt[executed];
        executed++;
        errors[it] += c;
        if (c == 0) break;
    }
    // state after `executed` iterations
    if (executed < iterations) {
        frame = input;
        std::vector<int> tmp(std::max(executed, 1));
        int r2 = 0;
        if (executed > 0) {
            rc = ldpc_b200_decode_batch(g, frame.data(), 1, executed, tmp.data(), &r2);
            if (rc) return rc;
        }
    }
    for (int i = 0; i < n; i++) Mvc[i] = frame[i];
    return it;
}
extern "C" int message_passing(int* Mvc, int iterations, int* variable_to_check_list, int* check_to_variable_list,
                               int* errors, int n, int k, int dv, int dc) {
    return guarded("message_passing", [&] { return message_passing_impl(Mvc, iterations, variable_to_check_list, check_to_variable_list, errors, n, k, dv, dc); });
}

