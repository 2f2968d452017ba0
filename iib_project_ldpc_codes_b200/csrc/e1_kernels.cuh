// e1_kernels.cuh -- "E1": shared-memory-resident bit-sliced BEC flooding decoder (sm_100a).
//
// Hot path of /root/reference/message_passing.c:7-82 as driven by the frame loops of
// /root/reference/parallel_simulator_expurgated.py:200-256 and :367-392, redesigned for B200:
//
//   * W = 32*WORDS frames share one machine word per Tanner-graph node ("frame group").
//     The reference keeps ONE state value per variable (Mvc[i] in {0,1,2}, message_passing.c:54-60);
//     its erasure plane obeys   x_i <- x_i AND (for every check c of i: c has >= 2 erased sockets)
//     (SURVEY.md 7.1), so one flooding iteration is, per check, dc gathered words combined with
//         two |= one & w ; one |= w
//     and, per variable, dv gathered `two` words ANDed into x.  No tensor-core shaped work.
//   * The graph (packed uint16 socket lists, socklist.hpp) is staged ONCE per CTA in shared memory;
//     each team of warps owns a frame group whose state (x[n], two[m]) also lives in shared
//     memory, so HBM is touched only by the final counter REDs (Monte-Carlo) or the
//     pattern/result planes (batched decode).
//   * Persistent CTAs (one per SM); teams pull frame groups from a global work counter.
//   * Early termination / stall detection / erasure counts are warp REDUX (__reduce_or_sync,
//     __reduce_add_sync) over the lanes of a team.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "bec_rng.cuh"
#include "socklist.hpp"
#include "e1_tmem.cuh"

namespace ldpc {

constexpr int E1_MAX_THREADS = 512;
#ifndef LDPC_W64_THREADS
#define LDPC_W64_THREADS 768
#endif
#ifndef LDPC_W128_THREADS
#define LDPC_W128_THREADS 768
#endif
#ifndef LDPC_CHECK_UNROLL
#define LDPC_CHECK_UNROLL 1
#endif
#ifndef LDPC_VAR_UNROLL
#define LDPC_VAR_UNROLL 3
#endif
#ifndef LDPC_VAR_UNROLL_VG
#define LDPC_VAR_UNROLL_VG 4
#endif
#ifndef LDPC_TM_PIPE
#define LDPC_TM_PIPE 1   // software-pipeline the tensor-memory phases two trips deep (when the registers allow)
#endif
constexpr int kCheckUnroll = LDPC_CHECK_UNROLL, kVarUnroll = LDPC_VAR_UNROLL, kVarUnrollVG = LDPC_VAR_UNROLL_VG;
// one word per node (W = 32) needs few registers: run 32 warps per SM to hide shared-memory latency
__host__ __device__ constexpr int e1_threads(int words) { return words == 1 ? 1024 : (words == 4 ? LDPC_W128_THREADS : LDPC_W64_THREADS); }

template <int WORDS> struct VecT;
template <> struct alignas(4)  VecT<1> { uint32_t a[1]; };
template <> struct alignas(8)  VecT<2> { uint32_t a[2]; };
template <> struct alignas(16) VecT<4> { uint32_t a[4]; };

#define LDPC_FORW _Pragma("unroll") for (int w = 0; w < WORDS; w++)

// one LOP3 with an explicit truth table (nvcc splits e.g. the 3-input majority into two)
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}

// raw device counter block per erasure point, all uint64:
//   [0 .. cap]            per-iteration erasure sums (bin 0 = initial count), counted frames only
//   [cap+1 .. cap+5]      block errors, bit errors, frames, counted frames, executed iterations
//   [cap+6 .. 2cap+7]     tail difference array: tail[b] is added to every bin >= b (finalize kernel)
//   [2cap+8]              physical frame-iterations (W * iterations the group actually ran)
//   [2cap+9]              groups that needed the masked second pass (expurgation)
__host__ __device__ inline int raw_stride(int cap) { return 2 * cap + 10; }

struct GraphGenParams {
    uint32_t seed_lo, seed_hi;      // graph seed
    unsigned long long first_graph; // global index of graph 0 of this launch
    int nb_log2, id_bits;           // bucket count of the socket sort (log2), bits that hold a socket id
    uint32_t sm_sort, sm_bkt, sm_fill, sm_conf;   // byte offsets relative to the team base (alias the state region)
    unsigned long long* fail_flag;  // incremented if a graph could not be generated
};

struct E1Params {
    // graph (global memory, packed uint16 socket lists: socklist.hpp)
    const uint16_t* c2v;
    const uint16_t* v2c;
    const uint16_t* var_orig;  // slot -> original variable id (keys the erasure generator); NULL = identity
    int n, m, dv, dc, n_pad, m_pad;
    // geometry
    int team_warps, teams;
    int v2c_global;   // 1: variable-side lookups stay in global memory (not staged into shared memory)
    int v2c_res;      // v2c_global: lookups of variables [0, v2c_res) are nevertheless resident in shared memory (as much
                      // as fits; a multiple of the team's threads), only the rest is streamed from L2 every iteration
    uint32_t sm_vres; // byte offset of that resident part (packed lists with pad = v2c_res)
    // socket lists in tensor memory (e1_tmem.cuh): ready-made gather addresses per thread, no lists in shared memory
    int tm;                                 // 1: on (kernels instantiated with TM = true)
    int tm_cpw, tm_trips_c, tm_trips_v;     // columns per warp, check / variable trips per thread
    int tm_full_c, tm_full_v;               // trips in which every thread of a team has a node (m / T, n / T)
    // overlapped stage 0 (e1_overlap.cuh): warps [0, team_warps) decode, warps [team_warps, team_warps + ov_np) draw the
    // next group's erasure pattern into a staging buffer meanwhile
    int ov_np;                              // producer (channel) warps; 0 = not the overlapped kernel
    uint32_t sm_stage, sm_q, sm_pscr, sm_mbox;   // staging pattern | producers' private draw queues | producer scratch | mailbox
    // ensemble mode (parallel_simulator_expurgated.py:200-223: a fresh random graph per frame): one graph
    // per `ens_gpg` frame groups, each team stages the graph of its work item into its own region
    // (sm_c2v / sm_v2c are then relative to the team base); graph g lives at c2v + g*ens_stride_c
    int ens;
    unsigned long long ens_stride_c, ens_stride_v;   // uint16 elements per graph
    unsigned long long ens_gpg;                      // frame groups per graph
    unsigned long long ens_wpg;                      // 32-frame words between the first frames of consecutive graphs
    unsigned long long ens_fpg;                      // frames per graph
    GraphGenParams gg;                               // ens == 2: graphs are generated in the kernel (e1_graphgen.cuh)
    // shared-memory byte offsets
    uint32_t sm_c2v, sm_v2c, sm_team0, sm_team_stride;
    uint32_t sm_x, sm_t, sm_v, sm_p, sm_hbuf, sm_scr;  // relative to the team base
    // Monte-Carlo work
    const BecParams* bec;  // [npoints]
    // 2^16 and 2^16 * sizeof(VecT<WORDS>): multipliers of the index unpack (gather_offs); passed at run time so
    // that ptxas keeps them as IMAD.WIDE on the idle FMA pipe instead of shift/mask chains on the ALU pipe
    uint32_t k16, k16s, vec_bytes;
    uint32_t draw_qcap;   // test hook: cap on the continuation queue of team_draw (0 = as large as `two` allows)
    uint32_t seed_lo, seed_hi;   // run seed (same in every bec[p]); a kernel parameter so the Philox round keys are uniform
    uint32_t rk[20];             // the 20 Philox round keys of that seed (philox_round_keys): constant-bank operands
    int npoints, cap, expurgation;
    uint32_t point0;                    // erasure point p of this launch is keyed as point0 + p by the channel generator
    unsigned long long first_word;      // word index q of the first frame (first_frame / 32)
    unsigned long long num_frames;      // frames per point
    unsigned long long groups_per_point;
    unsigned long long* counters;       // [npoints][raw_stride(cap)]
    unsigned long long* work_counter;
    // expurgation re-count (pass 2 of the two-pass expurgated run, e1_staged.cuh): the work items are groups of the
    // frames the staged pass listed (chunk-relative frame ids); denominators are NOT added again
    int recount;
    const uint32_t* rc_ids;              // [npoints][rc_stride]
    const unsigned long long* rc_cnt;    // [npoints]
    unsigned long long rc_stride;
    // batched decode work (WORDS == 1)
    const int8_t* mvc8;    // != NULL: ONE group (<= 32 frames) read from / written to int8 symbols [frames][n] in {0,1,2}
                           // directly (the single-frame drop-in: no pack / unpack kernels, one launch per call)
    const uint32_t* xin;   // [groups][n] erasure plane
    const uint32_t* vin;   // [groups][n] value plane
    uint32_t* xout;
    uint32_t* vout;
    int* errors;           // [frames][cap]
    int* ret;              // [frames]
};

struct TeamCtx {
    int tid;       // thread index inside the team
    int nthr;      // threads in the team
    int lane, warp;  // warp = warp index inside the team
    int team_warps;
    int bar_id;
    uint32_t* scr;   // team scratch (shared)
    uint32_t sa_x, sa_t;     // shared-window addresses of the team's x[] and two[]
    unsigned char* sm;       // start of the CTA's dynamic shared memory
    uint32_t ta_c, ta_v;     // TM: tensor-memory addresses of this thread's check / variable columns
};

__device__ __forceinline__ void team_sync(const TeamCtx& t) {
    if (t.team_warps == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(t.bar_id), "r"(t.nthr) : "memory");
}

// OR-reduce NOR words and add-reduce one count across the team; result valid in every thread.
// Contains one team_sync (which also orders the shared-memory state written before the call).
template <int NOR>
__device__ __forceinline__ void team_reduce(const TeamCtx& t, uint32_t (&ors)[NOR], uint32_t& sum) {
#pragma unroll
    for (int k = 0; k < NOR; k++) ors[k] = __reduce_or_sync(0xFFFFFFFFu, ors[k]);
    sum = __reduce_add_sync(0xFFFFFFFFu, sum);
    if (t.team_warps == 1) { __syncwarp(); return; }
    // scratch layout: [warp][NOR+1]; a single buffer is enough: the next write happens after the
    // next check-phase barrier, which every warp reaches only after its reads below.
    if (t.lane == 0) {
#pragma unroll
        for (int k = 0; k < NOR; k++) t.scr[t.warp * (NOR + 1) + k] = ors[k];
        t.scr[t.warp * (NOR + 1) + NOR] = sum;
    }
    team_sync(t);
    const bool have = t.lane < t.team_warps;
#pragma unroll
    for (int k = 0; k < NOR; k++) {
        uint32_t v = have ? t.scr[t.lane * (NOR + 1) + k] : 0u;
        ors[k] = __reduce_or_sync(0xFFFFFFFFu, v);
    }
    uint32_t s = have ? t.scr[t.lane * (NOR + 1) + NOR] : 0u;
    sum = __reduce_add_sync(0xFFFFFFFFu, s);
}

}  // namespace ldpc
#include "e1_graphgen.cuh"
namespace ldpc {

// ---- packed socket lists (socklist.hpp): all D neighbour slots of `node` with one or two vector loads ----
template <bool GLOBAL> __device__ __forceinline__ uint2 sl_ld64(const uint16_t* p) {
    return GLOBAL ? __ldg(reinterpret_cast<const uint2*>(p)) : *reinterpret_cast<const uint2*>(p);
}
template <bool GLOBAL> __device__ __forceinline__ uint32_t sl_ld32(const uint16_t* p) {
    return GLOBAL ? __ldg(reinterpret_cast<const uint32_t*>(p)) : *reinterpret_cast<const uint32_t*>(p);
}
template <bool GLOBAL> __device__ __forceinline__ uint32_t sl_ld16(const uint16_t* p) {
    return GLOBAL ? (uint32_t)__ldg(p) : (uint32_t)*p;
}
// Shared-window addresses of the D state vectors a node gathers: addr[j] = base + slot_j * sizeof(VecT).
// A packed pair q = lo | hi << 16 is taken apart by multiply-adds on the FMA pipe (no shifts or masks on the
// ALU pipe, which carries the decoder's LOP3 work); the multipliers are run-time values (E1Params) so that
// ptxas does not turn them back into shift/mask chains:
//     t = q * 2^16                        -> t.hi = hi,  t.lo = lo << 16           (IMAD.WIDE)
//     addr_hi = t.hi * S + base                                                     (IMAD)
//     addr_lo = hi32(t.lo * (2^16 * S) + (base << 32)) = lo * S + base              (IMAD.HI)
// FMA_UNPACK = false: plain shift/mask unpack (ALU pipe) -- better for W <= 64, where the LOP3 work per node is
// small and the extra IMAD.WIDE latency is not.
template <int D, bool GLOBAL, bool FMA_UNPACK>
__device__ __forceinline__ void gather_addrs(const E1Params& P, const uint16_t* __restrict__ list, int pad, int node,
                                             uint32_t base, uint32_t vec_bytes, uint32_t (&addr)[D]) {
    constexpr int NQ = D / 4, REM = D % 4;
    const unsigned long long add = (unsigned long long)base << 32;
    auto pair = [&](uint32_t q, uint32_t& a_lo, uint32_t& a_hi) {
        if (FMA_UNPACK) {
            const unsigned long long t = (unsigned long long)q * P.k16;
            a_hi = (uint32_t)(t >> 32) * P.vec_bytes + base;
            a_lo = (uint32_t)(((unsigned long long)(uint32_t)t * P.k16s + add) >> 32);
        } else {
            a_lo = (q & 0xFFFFu) * vec_bytes + base;
            a_hi = (q >> 16) * vec_bytes + base;
        }
    };
#pragma unroll
    for (int g = 0; g < NQ; g++) {
        const uint2 q = sl_ld64<GLOBAL>(list + ((size_t)g * pad + node) * 4);
        pair(q.x, addr[4 * g], addr[4 * g + 1]);
        pair(q.y, addr[4 * g + 2], addr[4 * g + 3]);
    }
    if (REM >= 2) {
        const uint32_t pr = sl_ld32<GLOBAL>(list + (size_t)4 * NQ * pad + (size_t)node * 2);
        pair(pr, addr[4 * NQ], addr[4 * NQ + 1]);
    }
    if (REM & 1) addr[D - 1] = sl_ld16<GLOBAL>(list + (size_t)(D - 1) * pad + node) * (FMA_UNPACK ? P.vec_bytes : vec_bytes) + base;
}

// state vector at a shared-window address (LDS / LDS.64 / LDS.128).  Deliberately neither volatile nor a
// memory clobber (that would pin the loads behind the L2 index loads of the next unrolled variable): the
// address always derives from index loads issued after the phase's team barrier, and within a phase the
// array gathered from (x or two) is never the array stored to, so no reordering the compiler may do is visible.
template <int WORDS> __device__ __forceinline__ VecT<WORDS> lds_vec(uint32_t a);
template <> __device__ __forceinline__ VecT<1> lds_vec<1>(uint32_t a) {
    VecT<1> v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v.a[0]) : "r"(a));
    return v;
}
template <> __device__ __forceinline__ VecT<2> lds_vec<2>(uint32_t a) {
    VecT<2> v;
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.a[0]), "=r"(v.a[1]) : "r"(a));
    return v;
}
template <> __device__ __forceinline__ VecT<4> lds_vec<4>(uint32_t a) {
    VecT<4> v;
    asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.a[0]), "=r"(v.a[1]), "=r"(v.a[2]), "=r"(v.a[3]) : "r"(a));
    return v;
}

// the same loads as volatile asm: for addresses that do NOT derive from a load issued after the phase barrier
template <int WORDS> __device__ __forceinline__ VecT<WORDS> lds_vec_ordered(uint32_t a);
template <> __device__ __forceinline__ VecT<1> lds_vec_ordered<1>(uint32_t a) {
    VecT<1> v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.a[0]) : "r"(a) : "memory");
    return v;
}
template <> __device__ __forceinline__ VecT<2> lds_vec_ordered<2>(uint32_t a) {
    VecT<2> v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.a[0]), "=r"(v.a[1]) : "r"(a) : "memory");
    return v;
}
template <> __device__ __forceinline__ VecT<4> lds_vec_ordered<4>(uint32_t a) {
    VecT<4> v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.a[0]), "=r"(v.a[1]), "=r"(v.a[2]), "=r"(v.a[3]) : "r"(a) : "memory");
    return v;
}

// ---- check phase:  two[c] = [>= 2 erased sockets]  (message_passing.c:24-45 collapsed) ----
template <int WORDS, int DV, int DC, bool VALUES>
__device__ __forceinline__ void check_phase(const E1Params& P, const TeamCtx& t, const uint16_t* __restrict__ s_c2v,
                                            const VecT<WORDS>* __restrict__ x, VecT<WORDS>* __restrict__ two,
                                            const VecT<WORDS>* __restrict__ val, VecT<WORDS>* __restrict__ par) {
    const int dc = DC ? DC : P.dc;
    const int m = P.m, m_pad = P.m_pad;
#pragma unroll kCheckUnroll
    for (int c = t.tid; c < m; c += t.nthr) {
        VecT<WORDS> one, tw, pr;
        LDPC_FORW { one.a[w] = 0u; tw.a[w] = 0u; pr.a[w] = 0u; }
        if (DC >= 3 && !VALUES) {
            // sockets in triples: "some" = a|b|c and "at least two" = majority(a,b,c) are one LOP3 each,
            // and triples combine with two |= maj | (one & some): 7 logic ops per word for dc = 6 instead
            // of 12, with a dependency depth of 3 instead of 6
            uint32_t addr[DC ? DC : 1];
            gather_addrs<DC ? DC : 1, false, WORDS == 4>(P, s_c2v, m_pad, c, t.sa_x, (uint32_t)sizeof(VecT<WORDS>), addr);
            auto X = [&](int j) { return lds_vec<WORDS>(addr[j]); };
#pragma unroll
            for (int j = 0; j + 2 < DC; j += 3) {
                const VecT<WORDS> a = X(j), b = X(j + 1), d = X(j + 2);
                LDPC_FORW {
                    const uint32_t some = lop3<0xFE>(a.a[w], b.a[w], d.a[w]);   // a | b | c
                    const uint32_t maj = lop3<0xE8>(a.a[w], b.a[w], d.a[w]);    // majority(a, b, c)
                    if (j == 0) { tw.a[w] = maj; one.a[w] = some; }
                    else {
                        tw.a[w] = lop3<0xFE>(tw.a[w], maj, one.a[w] & some);
                        one.a[w] |= some;
                    }
                }
            }
#pragma unroll
            for (int j = DC - DC % 3; j < DC; j++) {   // dc not a multiple of 3: the last one or two sockets
                const VecT<WORDS> xv = X(j);
                LDPC_FORW { tw.a[w] |= one.a[w] & xv.a[w]; one.a[w] |= xv.a[w]; }
            }
        } else {
#pragma unroll
            for (int j = 0; j < dc; j++) {
                const int v = s_c2v[sock_off(dc, m_pad, j, c)];
                const VecT<WORDS> xv = x[v];
                LDPC_FORW { tw.a[w] |= one.a[w] & xv.a[w]; one.a[w] |= xv.a[w]; }
                if (VALUES) {
                    const VecT<WORDS> vv = val[v];  // invariant: value bits are 0 where erased
                    LDPC_FORW pr.a[w] ^= vv.a[w];
                }
            }
        }
        two[c] = tw;
        if (VALUES) par[c] = pr;
    }
}

// ---- variable phase:  x_i &= AND_j two[v2c[i][j]]   (message_passing.c:52-65 collapsed) ----
// VALUES: a variable resolved in this iteration takes the parity of the LAST listed check that
// resolves it (message_passing.c:56-61 has no break: last writer wins).
// VG: the variable-side lookups are streamed from global memory / L2 (coalesced, read-only) instead of
// shared memory, which frees room for a wider frame group when the graph is large (n = 10^4).
template <int WORDS, int DV, int DC, bool VALUES, bool MASKED, bool VG = false>
__device__ __forceinline__ void var_phase(const E1Params& P, const TeamCtx& t, const uint16_t* __restrict__ s_v2c,
                                          VecT<WORDS>* __restrict__ x, const VecT<WORDS>* __restrict__ two,
                                          VecT<WORDS>* __restrict__ val, const VecT<WORDS>* __restrict__ par,
                                          const uint32_t (&mask)[WORDS], uint32_t (&any)[WORDS],
                                          uint32_t (&chg)[WORDS], uint32_t& cnt,
                                          int first = 0, int last = -1, int pad = -1, bool reset = true) {
    // variables [first, last) with lookups in `s_v2c` (packed lists of `pad` nodes); reset = false accumulates
    const int dv = DV ? DV : P.dv;
    const int n = last < 0 ? P.n : last, n_pad = pad < 0 ? P.n_pad : pad;
    if (reset) {
        LDPC_FORW { any[w] = 0u; chg[w] = 0u; }
        cnt = 0u;
    }
#pragma unroll(VG ? kVarUnrollVG : (WORDS == 4 ? 1 : kVarUnroll))   // measured: W = 128 best rolled, W <= 64 unrolled x3
    for (int i = first + t.tid; i < n; i += t.nthr) {
        const VecT<WORDS> xi = x[i];
        VecT<WORDS> xn = xi, nv;
        LDPC_FORW nv.a[w] = 0u;
        uint32_t addr[DV ? DV : 1];
        if (DV) gather_addrs<DV ? DV : 1, VG, WORDS == 4>(P, s_v2c, n_pad, i, t.sa_t, (uint32_t)sizeof(VecT<WORDS>), addr);
#pragma unroll
        for (int j = 0; j < dv; j++) {
            const int c = DV ? (int)((addr[DV ? j : 0] - t.sa_t) / (uint32_t)sizeof(VecT<WORDS>))   // only used by VALUES
                             : (int)sl_ld16<VG>(s_v2c + sock_off(dv, n_pad, j, i));
            const VecT<WORDS> tw = DV ? lds_vec<WORDS>(addr[DV ? j : 0]) : two[c];
            LDPC_FORW xn.a[w] &= tw.a[w];
            if (VALUES) {
                const VecT<WORDS> pr = par[c];
                LDPC_FORW nv.a[w] = (nv.a[w] & tw.a[w]) | (pr.a[w] & ~tw.a[w]);
            }
        }
        bool changed = false;
        LDPC_FORW changed |= (xn.a[w] != xi.a[w]);
        if (!VALUES) {
            x[i] = xn;   // unconditional: a predicated store costs more issue slots than the store saves
        } else if (changed) {
            x[i] = xn;
            if (VALUES) {
                VecT<WORDS> vi = val[i];
                LDPC_FORW vi.a[w] |= nv.a[w] & xi.a[w] & ~xn.a[w];
                val[i] = vi;
            }
        }
        LDPC_FORW {
            any[w] |= xn.a[w];
            chg[w] |= xi.a[w] ^ xn.a[w];
            cnt += __popc(MASKED ? (xn.a[w] & mask[w]) : xn.a[w]);
        }
    }
}

// variable phase of the Monte-Carlo kernels: shared-memory lists, or (large graphs) a resident part + an L2-streamed rest
template <int WORDS, int DV, int DC, bool MASKED>
__device__ __forceinline__ void mc_var_phase(const E1Params& P, const TeamCtx& t, const uint16_t* s_v2c, const uint16_t* s_vres,
                                             VecT<WORDS>* x, const VecT<WORDS>* two, const uint32_t (&mask)[WORDS],
                                             uint32_t (&any)[WORDS], uint32_t (&chg)[WORDS], uint32_t& cnt) {
    if (!P.v2c_global) {
        var_phase<WORDS, DV, DC, false, MASKED, false>(P, t, s_v2c, x, two, nullptr, nullptr, mask, any, chg, cnt);
    } else {
        // the streamed part first: its L2 latency overlaps the resident part of the warps that have none
        var_phase<WORDS, DV, DC, false, MASKED, true>(P, t, P.v2c, x, two, nullptr, nullptr, mask, any, chg, cnt, P.v2c_res, -1, -1, true);
        if (P.v2c_res)
            var_phase<WORDS, DV, DC, false, MASKED, false>(P, t, s_vres, x, two, nullptr, nullptr, mask, any, chg, cnt, 0, P.v2c_res,
                                                           P.v2c_res, false);
    }
}

// ------------------------------------------------------------------------------------------------
// The same two phases with the socket lists in tensor memory (e1_tmem.cuh): the gather addresses of a node come
// ready-made from tcgen05.ld, the addresses of the NEXT trip are requested as soon as this trip's gathers are
// issued, and nothing but state vectors travels through the shared-memory pipe.
// ------------------------------------------------------------------------------------------------
template <int WORDS> __device__ __forceinline__ void sts_vec(uint32_t a, const VecT<WORDS>& v);
template <> __device__ __forceinline__ void sts_vec<1>(uint32_t a, const VecT<1>& v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v.a[0]) : "memory");
}
template <> __device__ __forceinline__ void sts_vec<2>(uint32_t a, const VecT<2>& v) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(v.a[0]), "r"(v.a[1]) : "memory");
}
template <> __device__ __forceinline__ void sts_vec<4>(uint32_t a, const VecT<4>& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.a[0]), "r"(v.a[1]), "r"(v.a[2]), "r"(v.a[3]) : "memory");
}

// two = [>= 2 of the DC gathered words have the bit set]; sockets in triples (see check_phase)
template <int WORDS, int DC>
__device__ __forceinline__ VecT<WORDS> check_two(const VecT<WORDS> (&v)[DC]) {
    VecT<WORDS> one, tw;
    LDPC_FORW { one.a[w] = 0u; tw.a[w] = 0u; }
#pragma unroll
    for (int j = 0; j + 2 < DC; j += 3) {
        LDPC_FORW {
            const uint32_t some = lop3<0xFE>(v[j].a[w], v[j + 1].a[w], v[j + 2].a[w]);
            const uint32_t maj = lop3<0xE8>(v[j].a[w], v[j + 1].a[w], v[j + 2].a[w]);
            if (j == 0) { tw.a[w] = maj; one.a[w] = some; }
            else {
                tw.a[w] = lop3<0xFE>(tw.a[w], maj, one.a[w] & some);
                one.a[w] |= some;
            }
        }
    }
#pragma unroll
    for (int j = DC - DC % 3; j < DC; j++) LDPC_FORW { tw.a[w] |= one.a[w] & v[j].a[w]; one.a[w] |= v[j].a[w]; }
    return tw;
}

// One trip = one node per thread.  The trips are software-pipelined two deep: the gathers of trip k+1 are issued
// (and the addresses of trip k+2 requested from tensor memory, into the registers those gathers have just read)
// BEFORE trip k is combined and stored, so every warp keeps two nodes' worth of shared-memory loads in flight --
// with 24 warps per SM the gather latency was the top stall (ncu short-scoreboard).  No hazard: within a phase the
// array gathered from is never the array stored to, and a thread's own x vectors of different trips are distinct.
// full = trips in which every thread of the team has a node (host-computed: no integer division per phase).
template <int WORDS, int DC>
__device__ __forceinline__ void check_phase_tm(const E1Params& P, const TeamCtx& t) {
    const int T = P.tm_trips_c, F = P.tm_full_c;
    const uint32_t step = (uint32_t)t.nthr * (uint32_t)sizeof(VecT<WORDS>);
    uint32_t dst = t.sa_t + (uint32_t)t.tid * (uint32_t)sizeof(VecT<WORDS>);
    uint32_t ta = t.ta_c;
    uint32_t a[DC];
    tm_ld<DC>(ta, a);
    auto issue = [&](VecT<WORDS> (&v)[DC], bool more) {
        tm_wait_ld<DC>(a);
#pragma unroll
        for (int j = 0; j < DC; j++) v[j] = lds_vec<WORDS>(a[j]);
        ta += DC;
        if (more) tm_ld<DC>(ta, a);   // warp-uniform
    };
    auto finish = [&](const VecT<WORDS> (&v)[DC]) {
        sts_vec<WORDS>(dst, check_two<WORDS, DC>(v));
        dst += step;
    };
    VecT<WORDS> v0[DC], v1[DC];
    constexpr bool PIPE = LDPC_TM_PIPE && DC * WORDS <= 24;   // two sets of gathered vectors must fit the register budget
    if (PIPE && F > 0) {
        issue(v0, 1 < T);
        int k = 0;
#pragma unroll 1
        for (; k + 2 < F; k += 2) {
            issue(v1, true); finish(v0);          // trips k+1 and k+2 are full trips: a further trip follows each
            issue(v0, k + 3 < T); finish(v1);
        }
        if (k + 1 < F) { issue(v1, k + 2 < T); finish(v0); finish(v1); }
        else finish(v0);
    } else {
#pragma unroll 2
        for (int k = 0; k < F; k++) { issue(v0, k + 1 < T); finish(v0); }
    }
    if (F < T) {   // last trip: only the first lanes of the team have a check (the others' columns hold a harmless address)
        tm_wait_ld<DC>(a);
        if (F * t.nthr + t.tid < P.m) {
#pragma unroll
            for (int j = 0; j < DC; j++) v0[j] = lds_vec<WORDS>(a[j]);
            finish(v0);
        }
    }
}

template <int WORDS, int DV, bool MASKED>
__device__ __forceinline__ void var_phase_tm(const E1Params& P, const TeamCtx& t, const uint32_t (&mask)[WORDS],
                                             uint32_t (&any)[WORDS], uint32_t (&chg)[WORDS], uint32_t& cnt) {
    const int T = P.tm_trips_v, F = P.tm_full_v;
    const uint32_t step = (uint32_t)t.nthr * (uint32_t)sizeof(VecT<WORDS>);
    uint32_t xa = t.sa_x + (uint32_t)t.tid * (uint32_t)sizeof(VecT<WORDS>);   // own state vector of the trip being issued
    uint32_t xs = xa;                                                          // ... of the trip being finished
    uint32_t ta = t.ta_v;
    LDPC_FORW { any[w] = 0u; chg[w] = 0u; }
    cnt = 0u;
    uint32_t a[DV];
    tm_ld<DV>(ta, a);
    // v[0 .. DV-1] = gathered `two` vectors, v[DV] = own state (its address does not derive from a post-barrier load:
    // ordered load)
    auto issue = [&](VecT<WORDS> (&v)[DV + 1], bool more) {
        tm_wait_ld<DV>(a);
        v[DV] = lds_vec_ordered<WORDS>(xa);
#pragma unroll
        for (int j = 0; j < DV; j++) v[j] = lds_vec<WORDS>(a[j]);
        xa += step;
        ta += DV;
        if (more) tm_ld<DV>(ta, a);
    };
    auto finish = [&](const VecT<WORDS> (&v)[DV + 1]) {
        VecT<WORDS> xn = v[DV];
#pragma unroll
        for (int j = 0; j < DV; j++) LDPC_FORW xn.a[w] &= v[j].a[w];
        sts_vec<WORDS>(xs, xn);
        xs += step;
        LDPC_FORW {
            any[w] |= xn.a[w];
            chg[w] |= v[DV].a[w] ^ xn.a[w];
            cnt += __popc(MASKED ? (xn.a[w] & mask[w]) : xn.a[w]);
        }
    };
    VecT<WORDS> v0[DV + 1], v1[DV + 1];
    constexpr bool PIPE = LDPC_TM_PIPE && (DV + 1) * WORDS <= 24;
    if (PIPE && F > 0) {
        issue(v0, 1 < T);
        int k = 0;
#pragma unroll 1
        for (; k + 2 < F; k += 2) {
            issue(v1, true); finish(v0);
            issue(v0, k + 3 < T); finish(v1);
        }
        if (k + 1 < F) { issue(v1, k + 2 < T); finish(v0); finish(v1); }
        else finish(v0);
    } else {
#pragma unroll 2
        for (int k = 0; k < F; k++) { issue(v0, k + 1 < T); finish(v0); }
    }
    if (F < T) {
        tm_wait_ld<DV>(a);
        if (F * t.nthr + t.tid < P.n) {
            v0[DV] = lds_vec_ordered<WORDS>(xa);
#pragma unroll
            for (int j = 0; j < DV; j++) v0[j] = lds_vec<WORDS>(a[j]);
            finish(v0);
        }
    }
}

// every thread writes the gather addresses of its nodes into its tensor-memory columns (once per launch)
template <int WORDS, int DV, int DC>
__device__ __forceinline__ void tm_fill_lists(const E1Params& P, TeamCtx& t, uint32_t tbase) {
    const int wcta = (int)(threadIdx.x >> 5);
    const uint32_t ta0 = tbase + ((uint32_t)(32 * (wcta & 3)) << 16) + (uint32_t)((wcta >> 2) * P.tm_cpw);
    t.ta_c = ta0;
    t.ta_v = ta0 + (uint32_t)(P.tm_trips_c * DC);
    for (int tr = 0; tr < P.tm_trips_c; tr++) {
        const int c = tr * t.nthr + t.tid;
        uint32_t a[DC];
#pragma unroll
        for (int j = 0; j < DC; j++) a[j] = t.sa_x;
        if (c < P.m) gather_addrs<DC, true, false>(P, P.c2v, P.m_pad, c, t.sa_x, (uint32_t)sizeof(VecT<WORDS>), a);
        tm_st<DC>(t.ta_c + (uint32_t)(tr * DC), a);
    }
    for (int tr = 0; tr < P.tm_trips_v; tr++) {
        const int i = tr * t.nthr + t.tid;
        uint32_t a[DV];
#pragma unroll
        for (int j = 0; j < DV; j++) a[j] = t.sa_t;
        if (i < P.n) gather_addrs<DV, true, false>(P, P.v2c, P.n_pad, i, t.sa_t, (uint32_t)sizeof(VecT<WORDS>), a);
        tm_st<DV>(t.ta_v + (uint32_t)(tr * DV), a);
    }
    tm_wait_st();
}

__device__ __forceinline__ void red_add(unsigned long long* p, unsigned long long v) {
    if (v) atomicAdd(p, v);
}

// cold path of team_draw: finish a word in-thread when a continuation queue is full
__device__ __noinline__ uint32_t bec_finish(const BecParams& bp, uint32_t point, unsigned long long q, uint32_t var,
                                            int b, uint32_t eq) {
    uint32_t lt = 0u;
    for (;; b++) {
        bec_block(bp, point, q, var, b, lt, eq);
        if (bec_done_after(bp, b, eq)) break;
    }
    return lt;
}

// ---- channel draw for one frame group (channels.py:24-26), straight into shared memory ----
// Same generator as bec_word (bec_rng.cuh), bit for bit, scheduled differently: bec_word runs Philox blocks
// until the slowest of a word's 32 frames is decided, and inside a warp every thread waits for the slowest
// word of 32 variables (3.2 blocks per word on average).  Here every word takes exactly two blocks (eight
// levels, no exit tests); the ~12 % of words that still have an undecided frame are appended to a queue
// (in the `two` array, idle during the draw) and continued one block per round by whichever thread picks
// them up: 2.15 blocks per word.  Thread-local partial results: `red` (OR per word) and `cnt` (popcount).
// CHUNK = words of a variable whose Philox chains run side by side (2 * CHUNK independent chains): all of them where the
// registers allow, one at a time in the kernels that keep the socket lists in shared memory (the ensemble kernel at
// W = 128 lost 13 % to spills with four)
// RK: round keys read from the kernel parameters (E1Params::rk) instead of derived from the seed in the loop -- a win where
// the chains are interleaved (tensor-memory kernels), a loss in the register-bound ensemble kernel (1.34 -> 0.83e8 frames/s)
template <int WORDS, bool MASKED, int CHUNK = WORDS, bool RK = true>
__device__ __forceinline__ void team_draw(const E1Params& P, const TeamCtx& t, const BecParams& bp_, uint32_t point,
                                          unsigned long long q0, const uint32_t (&valid)[WORDS],
                                          const uint32_t (&mask)[WORDS], VecT<WORDS>* x, uint32_t* qbuf,
                                          uint32_t* qc, uint32_t (&red)[WORDS], uint32_t& cnt, uint32_t qentries = 0) {
    LDPC_FORW red[w] = 0u;
    cnt = 0u;
    BecParams bp = bp_;
    bp.seed_lo = P.seed_lo; bp.seed_hi = P.seed_hi;   // same values, but provably warp-uniform
    if (bp.mode != 1) {   // eps == 0 or eps == 1
        for (int i = t.tid; i < P.n; i += t.nthr) {
            VecT<WORDS> xi;
            LDPC_FORW {
                xi.a[w] = bp.mode == 2 ? valid[w] : 0u;
                red[w] |= xi.a[w];
                cnt += __popc(MASKED ? (xi.a[w] & mask[w]) : xi.a[w]);
            }
            x[i] = xi;
        }
        return;
    }
    uint32_t* xs = reinterpret_cast<uint32_t*>(x);
    uint2* q2 = reinterpret_cast<uint2*>(qbuf);
    // qentries != 0: the queue is a dedicated buffer of that many 8-byte entries (the overlapped stage-0 kernel,
    // e1_overlap.cuh, draws while `two` is in use); the variables are then drawn in slices small enough that the
    // expected 12 % of unfinished words (margin: 1/6) fit the first queue
    uint32_t entries = qentries ? qentries : (uint32_t)P.m * WORDS / 2u;
    if (P.draw_qcap && P.draw_qcap < entries) entries = P.draw_qcap;   // overflowing words finish in-thread (bec_finish)
    const uint32_t capB = entries / 8u, capA = entries - capB;
    uint2* const qA = q2;
    uint2* const qB = q2 + capA;
    int slice = P.n;
    if (qentries) slice = max(t.nthr, (int)(capA * 6u / WORDS) / t.nthr * t.nthr);
    const BecMasks8 M = bec_masks8(bp);
    const bool deeper = (bp.threshold << 8) != 0u;
    for (int s0 = 0; s0 < P.n; s0 += slice) {
    const int s1 = min(P.n, s0 + slice);
    if (t.tid < 8) qc[t.tid] = 0u;
    team_sync(t);
    // the original id of the NEXT variable is fetched (L2) one trip ahead, and the Philox chains of all the words of a
    // variable run side by side before any queue traffic: with few warps drawing (the producer warps of
    // e1_overlap.cuh) the loop is bound by latency, not by issue slots
    uint32_t var = 0u;
    if (s0 + t.tid < s1) var = P.var_orig ? (uint32_t)P.var_orig[s0 + t.tid] : (uint32_t)(s0 + t.tid);
    for (int i = s0 + t.tid; i < s1; i += t.nthr) {
        VecT<WORDS> xi;
        uint32_t var_next = 0u;
        if (i + t.nthr < s1) var_next = P.var_orig ? (uint32_t)P.var_orig[i + t.nthr] : (uint32_t)(i + t.nthr);
#pragma unroll
        for (int w0 = 0; w0 < WORDS; w0 += CHUNK) {
            uint32_t lt[CHUNK], eq[CHUNK];
#pragma unroll
            for (int c = 0; c < CHUNK; c++) {
                if (RK) bec_first8_rk(M, P.rk, point, q0 + w0 + c, var, lt[c], eq[c]);
                else bec_first8(bp, M, point, q0 + w0 + c, var, lt[c], eq[c]);
                eq[c] &= valid[w0 + c];
            }
#pragma unroll
            for (int c = 0; c < CHUNK; c++) {
                const int w = w0 + c;
                if (eq[c] != 0u && deeper) {
                    const uint32_t slot = atomicAdd(&qc[1], 1u);
                    if (slot < capA) qA[slot] = make_uint2((uint32_t)i * WORDS + w, eq[c]);
                    else lt[c] |= bec_finish(bp, point, q0 + w, var, 2, eq[c]);
                }
                xi.a[w] = lt[c] & valid[w];
                red[w] |= xi.a[w];
                cnt += __popc(MASKED ? (xi.a[w] & mask[w]) : xi.a[w]);
            }
        }
        x[i] = xi;
        var = var_next;
    }
    team_sync(t);
    for (int b = 2; b < 8; b++) {
        const bool fromA = (b & 1) == 0;
        const uint2* src = fromA ? qA : qB;
        uint2* dst = fromA ? qB : qA;
        const uint32_t capS = fromA ? capA : capB, capD = fromA ? capB : capA;
        const uint32_t cin = min(qc[b - 1], capS);
        if (cin == 0u) break;   // team-uniform
        for (uint32_t e = t.tid; e < cin; e += t.nthr) {
            const uint2 ent = src[e];
            const uint32_t i = ent.x / WORDS, w = ent.x % WORDS;
            const uint32_t var = P.var_orig ? (uint32_t)P.var_orig[i] : i;
            uint32_t lt = 0u, eq = ent.y;
            if (RK) bec_block_rk(bp, P.rk, point, q0 + w, var, b, lt, eq);
            else bec_block(bp, point, q0 + w, var, b, lt, eq);
            if (!bec_done_after(bp, b, eq)) {
                const uint32_t slot = atomicAdd(&qc[b], 1u);
                if (slot < capD) dst[slot] = make_uint2(ent.x, eq);
                else lt |= bec_finish(bp, point, q0 + w, var, b + 1, eq);
            }
            if (lt != 0u) {   // lt is a subset of the entry's eq, which was masked by `valid`
                xs[ent.x] |= lt;
#pragma unroll
                for (int k = 0; k < WORDS; k++) {
                    const uint32_t l = (w == (uint32_t)k) ? lt : 0u;
                    red[k] |= l;
                    cnt += __popc(MASKED ? (l & mask[k]) : l);
                }
            }
        }
        team_sync(t);
    }
    }   // slices
}

// stage the resident part of the variable-side lists (v2c_global): every packed array keeps its nodes [0, v2c_res)
__device__ __forceinline__ void stage_v2c_resident(const E1Params& P, unsigned char* smem) {
    if (!P.v2c_global || !P.v2c_res) return;
    uint16_t* dst = reinterpret_cast<uint16_t*>(smem + P.sm_vres);
    const int dv = P.dv, nq = dv >> 2, rem = dv & 3, R = P.v2c_res, n_pad = P.n_pad;
    for (int a = 0; a < nq; a++) {   // 8 bytes per node
        const uint2* g = reinterpret_cast<const uint2*>(P.v2c + (size_t)a * n_pad * 4);
        uint2* d = reinterpret_cast<uint2*>(dst + (size_t)a * R * 4);
        for (int i = threadIdx.x; i < R; i += blockDim.x) d[i] = __ldg(g + i);
    }
    if (rem >= 2) {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(P.v2c + (size_t)4 * nq * n_pad);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst + (size_t)4 * nq * R);
        for (int i = threadIdx.x; i < R; i += blockDim.x) d[i] = __ldg(g + i);
    }
    if (rem & 1) {
        const uint16_t* g = P.v2c + (size_t)(dv - 1) * n_pad;
        uint16_t* d = dst + (size_t)(dv - 1) * R;
        for (int i = threadIdx.x; i < R; i += blockDim.x) d[i] = __ldg(g + i);
    }
}

// team scratch, in 32-bit words: reduce area | work slot | per-frame counts (rf) | draw queue counters | frame ids
__host__ __device__ constexpr int e1_scr_words(int words) { return 32 * (2 * words + 1) + 2 + 32 * words + 8 + 32 * words; }

// Channel draw for a group of ARBITRARY frames (one chunk-relative frame id per lane in fid[]): a warp takes one
// variable at a time, lane l decides frame l's bit alone (bec_lane_bit) and a ballot assembles the bit-sliced word.
template <int WORDS, bool MASKED>
__device__ __forceinline__ void team_draw_lanes(const E1Params& P, const TeamCtx& t, const BecParams& bp_, uint32_t point,
                                                const uint32_t* fid, const uint32_t (&valid)[WORDS],
                                                const uint32_t (&mask)[WORDS], VecT<WORDS>* x, uint32_t (&red)[WORDS],
                                                uint32_t& cnt) {
    LDPC_FORW red[w] = 0u;
    cnt = 0u;
    BecParams bp = bp_;
    bp.seed_lo = P.seed_lo; bp.seed_hi = P.seed_hi;
    unsigned long long q[WORDS];
    uint32_t bit[WORDS];
    LDPC_FORW {
        const unsigned long long f = P.first_word * 32ull + (unsigned long long)fid[w * 32 + t.lane];
        q[w] = f >> 5; bit[w] = (uint32_t)f & 31u;
    }
    for (int i = t.warp; i < P.n; i += t.team_warps) {
        const uint32_t var = P.var_orig ? (uint32_t)P.var_orig[i] : (uint32_t)i;
        VecT<WORDS> xi;
        LDPC_FORW {
            const bool on = (valid[w] >> t.lane) & 1u;
            const uint32_t b = on ? bec_lane_bit(bp, point, q[w], var, bit[w]) : 0u;
            xi.a[w] = __ballot_sync(0xFFFFFFFFu, b != 0u);
        }
        if (t.lane == 0) {
            x[i] = xi;
            LDPC_FORW { red[w] |= xi.a[w]; cnt += __popc(MASKED ? (xi.a[w] & mask[w]) : xi.a[w]); }
        }
    }
}

// Decode one frame group for the Monte-Carlo path.  MASKED pass: erasure sums restricted to `mask`.
// Returns physical iterations executed; outputs the team-uniform masks / counts by reference.
template <int WORDS, int DV, int DC, bool MASKED, bool TM = false>
__device__ __forceinline__ int mc_group(const E1Params& P, const TeamCtx& t, const BecParams& bp, uint32_t point,
                                        unsigned long long q0, const uint32_t (&valid)[WORDS],
                                        const uint32_t (&mask)[WORDS], const uint16_t* s_c2v, const uint16_t* s_v2c,
                                        VecT<WORDS>* x, VecT<WORDS>* two, uint32_t* hbuf, uint32_t (&any)[WORDS],
                                        uint32_t& cnt, unsigned long long& exec_sum) {
    // draw the erasure pattern (channels.py:24-26) straight into shared memory
    uint32_t red[WORDS];
    if (P.recount)
        team_draw_lanes<WORDS, MASKED>(P, t, bp, point + P.point0, t.scr + 32 * (2 * WORDS + 1) + 2 + 32 * WORDS + 8, valid, mask, x, red, cnt);
    else
        team_draw<WORDS, MASKED, TM ? WORDS : 1, TM>(P, t, bp, point + P.point0, q0, valid, mask, x, reinterpret_cast<uint32_t*>(two),
                                 t.scr + 32 * (2 * WORDS + 1) + 2 + 32 * WORDS, red, cnt);
    team_reduce<WORDS>(t, red, cnt);
    LDPC_FORW any[w] = red[w];
    if (t.tid == 0) hbuf[0] = cnt;

    uint32_t running[WORDS];
    LDPC_FORW running[w] = MASKED ? mask[w] : valid[w];
    int it = 0;
    exec_sum = 0;
    for (; it < P.cap; it++) {
        uint32_t r = 0u, pc = 0u;
        LDPC_FORW { r |= running[w]; pc += __popc(running[w]); }
        if (r == 0u) break;
        exec_sum += pc;
        uint32_t oc[2 * WORDS], a2[WORDS], c2[WORDS];
        if constexpr (TM) {
            check_phase_tm<WORDS, DC>(P, t);
            team_sync(t);
            var_phase_tm<WORDS, DV, MASKED>(P, t, mask, a2, c2, cnt);
        } else {
            check_phase<WORDS, DV, DC, false>(P, t, s_c2v, x, two, nullptr, nullptr);
            team_sync(t);
            mc_var_phase<WORDS, DV, DC, MASKED>(P, t, s_v2c, reinterpret_cast<const uint16_t*>(t.sm + P.sm_vres), x, two, mask, a2, c2, cnt);
        }
        LDPC_FORW { oc[w] = a2[w]; oc[WORDS + w] = c2[w]; }
        team_reduce<2 * WORDS>(t, oc, cnt);
        if (t.tid == 0) hbuf[it + 1] = cnt;
        // message_passing.c:16-19: iterations 0 and 1 always run; from it>=2 a frame whose count did
        // not change in the previous iteration is skipped forever.  :76-78: stop when nothing is erased.
        LDPC_FORW {
            any[w] = oc[w];
            running[w] &= oc[w] & (it == 0 ? 0xFFFFFFFFu : oc[WORDS + w]);
        }
    }
    return it;
}

template <int WORDS, int DV, int DC, bool TM = false>
__global__ void __launch_bounds__(e1_threads(WORDS), 1) e1_mc_kernel(const __grid_constant__ E1Params P) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t tm_slot;
    uint16_t* s_c2v = reinterpret_cast<uint16_t*>(smem + P.sm_c2v);
    uint16_t* s_v2c = reinterpret_cast<uint16_t*>(smem + P.sm_v2c);

    // stage the graph once per CTA (16-byte vectors; arrays are padded to 16 bytes on the host)
    if (!P.ens && !TM) {
        const int dc = DC ? DC : P.dc, dv = DV ? DV : P.dv;
        const int nc = dc * P.m_pad / 8, nv = dv * P.n_pad / 8;
        const uint4* g0 = reinterpret_cast<const uint4*>(P.c2v);
        const uint4* g1 = reinterpret_cast<const uint4*>(P.v2c);
        uint4* d0 = reinterpret_cast<uint4*>(s_c2v);
        uint4* d1 = reinterpret_cast<uint4*>(s_v2c);
        for (int i = threadIdx.x; i < nc; i += blockDim.x) d0[i] = __ldg(g0 + i);
        if (!P.v2c_global) for (int i = threadIdx.x; i < nv; i += blockDim.x) d1[i] = __ldg(g1 + i);
        stage_v2c_resident(P, smem);
    }
    __syncthreads();

    TeamCtx t;
    t.team_warps = P.team_warps;
    t.nthr = P.team_warps * 32;
    const int team = threadIdx.x / t.nthr;
    t.tid = threadIdx.x - team * t.nthr;
    t.lane = threadIdx.x & 31;
    t.warp = t.tid >> 5;
    t.bar_id = 1 + team;
    unsigned char* base = smem + P.sm_team0 + (size_t)team * P.sm_team_stride;
    VecT<WORDS>* x = reinterpret_cast<VecT<WORDS>*>(base + P.sm_x);
    VecT<WORDS>* two = reinterpret_cast<VecT<WORDS>*>(base + P.sm_t);
    t.sa_x = (uint32_t)__cvta_generic_to_shared(x); t.sa_t = (uint32_t)__cvta_generic_to_shared(two); t.sm = smem;
    uint32_t* hbuf = reinterpret_cast<uint32_t*>(base + P.sm_hbuf);
    t.scr = reinterpret_cast<uint32_t*>(base + P.sm_scr);
    // scratch map: [0 .. 32*(2*WORDS+1))  reduce area ; then [work slot (2 words)] ; then rf[32*WORDS]
    uint32_t* work_slot = t.scr + 32 * (2 * WORDS + 1);
    uint32_t* rf = work_slot + 2;
    uint32_t tm_base = 0;
    if constexpr (TM) {   // socket lists -> tensor memory (e1_tmem.cuh)
        tm_base = tm_alloc_all(&tm_slot);
        tm_fill_lists<WORDS, DV, DC>(P, t, tm_base);
    }

    unsigned long long total = (unsigned long long)P.npoints * P.groups_per_point;
    const int W = 32 * WORDS;
    const int stride = raw_stride(P.cap);
    uint32_t* fid = rf + 32 * WORDS + 8;   // recount: chunk-relative ids of the group's frames
    if (P.recount) {
        total = 0;
        for (int p = 0; p < P.npoints; p++) total += (P.rc_cnt[p] + W - 1) / W;
    }
    if (P.ens) {
        s_c2v = reinterpret_cast<uint16_t*>(base + P.sm_c2v);
        s_v2c = reinterpret_cast<uint16_t*>(base + P.sm_v2c);
    }
    long long staged_graph = -1;

    // Ensemble mode: a team takes ALL the frame groups of a graph (ens_gpg consecutive items), so the graph is built once
    // per graph, not once per group (groups_per_point is a multiple of ens_gpg: a block never straddles two graphs).
    const unsigned long long fetch = (P.ens && !P.recount && P.ens_gpg > 1) ? P.ens_gpg : 1ull;
    unsigned long long next_item = 0, items_left = 0;
    for (;;) {
        // ---- fetch a work item (point, group) ----
        unsigned long long item;
        if (items_left > 0) {   // team-uniform
            item = next_item;
        } else if (t.team_warps == 1) {
            unsigned long long v = 0;
            if (t.lane == 0) v = atomicAdd(P.work_counter, fetch);
            item = __shfl_sync(0xFFFFFFFFu, v, 0);
            items_left = fetch;
        } else {
            if (t.tid == 0) {
                unsigned long long v = atomicAdd(P.work_counter, fetch);
                work_slot[0] = (uint32_t)v; work_slot[1] = (uint32_t)(v >> 32);
            }
            team_sync(t);
            item = (unsigned long long)work_slot[0] | ((unsigned long long)work_slot[1] << 32);
            items_left = fetch;
        }
        next_item = item + 1; items_left--;
        if (item >= total) break;
        uint32_t point = (uint32_t)(item / P.groups_per_point);
        unsigned long long gi = item - (unsigned long long)point * P.groups_per_point;
        uint32_t valid[WORDS];
        if (P.recount) {   // group gi of the listed frames of `point`
            unsigned long long rest = item;
            int p = 0;
            for (; p < P.npoints; p++) {
                const unsigned long long gp = (P.rc_cnt[p] + W - 1) / W;
                if (rest < gp) break;
                rest -= gp;
            }
            point = (uint32_t)p; gi = rest;
            const unsigned long long left = P.rc_cnt[p] - gi * W;
            LDPC_FORW {
                const unsigned long long f0 = (unsigned long long)w * 32ull;
                valid[w] = f0 >= left ? 0u : (left - f0 >= 32ull ? 0xFFFFFFFFu : ((1u << (uint32_t)(left - f0)) - 1u));
            }
            for (int f = t.tid; f < W; f += t.nthr)
                fid[f] = (unsigned long long)f < left ? P.rc_ids[(size_t)p * P.rc_stride + gi * W + f] : 0u;
            team_sync(t);
        }
        unsigned long long q0 = P.first_word + gi * WORDS;
        const BecParams bp = P.bec[point];

        if (P.recount) {
        } else if (!P.ens) {
            LDPC_FORW {
                const unsigned long long f0 = (gi * WORDS + w) * 32ull;  // first frame of this word, relative
                valid[w] = f0 >= P.num_frames ? 0u
                           : (P.num_frames - f0 >= 32ull ? 0xFFFFFFFFu : ((1u << (uint32_t)(P.num_frames - f0)) - 1u));
            }
        } else {
            const unsigned long long g = gi / P.ens_gpg, gg = gi - g * P.ens_gpg;
            if ((long long)g != staged_graph && P.ens == 2) {
                // in-kernel generator (e1_graphgen.cuh); its scratch aliases the state region, free between groups
                if (!team_generate_graph<DV, DC>(P, P.gg, t, base, P.gg.first_graph + g, s_c2v, s_v2c)) {
                    if (t.tid == 0) atomicAdd(P.gg.fail_flag, 1ull);
                    team_sync(t);
                    continue;
                }
                staged_graph = (long long)g;
            } else if ((long long)g != staged_graph) {   // the previous group of this team ended with a team_sync
                const int dc = DC ? DC : P.dc, dv = DV ? DV : P.dv;
                const int nc = dc * P.m_pad / 8, nv = dv * P.n_pad / 8;
                const uint4* g0 = reinterpret_cast<const uint4*>(P.c2v + g * P.ens_stride_c);
                const uint4* g1 = reinterpret_cast<const uint4*>(P.v2c + g * P.ens_stride_v);
                uint4* d0 = reinterpret_cast<uint4*>(s_c2v);
                uint4* d1 = reinterpret_cast<uint4*>(s_v2c);
                for (int i = t.tid; i < nc; i += t.nthr) d0[i] = __ldg(g0 + i);
                for (int i = t.tid; i < nv; i += t.nthr) d1[i] = __ldg(g1 + i);
                team_sync(t);
                staged_graph = (long long)g;
            }
            q0 = P.first_word + g * P.ens_wpg + gg * WORDS;
            LDPC_FORW {
                const unsigned long long f0 = (gg * WORDS + w) * 32ull;
                valid[w] = f0 >= P.ens_fpg ? 0u
                           : (P.ens_fpg - f0 >= 32ull ? 0xFFFFFFFFu : ((1u << (uint32_t)(P.ens_fpg - f0)) - 1u));
            }
        }

        uint32_t any[WORDS], cnt;
        unsigned long long exec_sum;
        int iters = mc_group<WORDS, DV, DC, false, TM>(P, t, bp, point, q0, valid, valid, s_c2v, s_v2c, x, two, hbuf,
                                                       any, cnt, exec_sum);
        unsigned long long* C = P.counters + (size_t)point * stride;
        unsigned long long phys = (unsigned long long)iters * W;
        uint32_t nvalid = 0;
        LDPC_FORW nvalid += __popc(valid[w]);

        // ---- expurgation predicate: a frame counts iff its final residual > expurgation
        //      (parallel_simulator_expurgated.py:238); denominators still count every frame (:258-260)
        uint32_t K[WORDS];
        LDPC_FORW K[w] = valid[w];
        bool second = false;
        if (P.expurgation >= 0) {
            // `any` already says which frames end with erasures left (residual > 0): with threshold 0 that IS the
            // predicate, and a group without such a frame counts nothing whatever the threshold -- the per-frame
            // residuals (a pass over the whole state, ~20 % of the ensemble kernel's instructions) are only counted
            // for a positive threshold in a group that has a failing frame
            uint32_t kall = 0, kany = 0, failing = 0;
            LDPC_FORW failing |= any[w] & valid[w];
            if (P.expurgation == 0 || failing == 0u) {
                LDPC_FORW K[w] = any[w] & valid[w];
            } else {
                for (int f = t.tid; f < W; f += t.nthr) rf[f] = 0u;
                team_sync(t);
                LDPC_FORW {
                    uint32_t c = 0;
                    for (int i = t.warp; i < P.n; i += t.team_warps) c += (x[i].a[w] >> t.lane) & 1u;
                    if (c) atomicAdd(&rf[w * 32 + t.lane], c);
                }
                team_sync(t);
                LDPC_FORW K[w] = __ballot_sync(0xFFFFFFFFu, (int)rf[w * 32 + t.lane] > P.expurgation) & valid[w];
                team_sync(t);  // rf reads done before anything reuses scratch
            }
            LDPC_FORW { kall |= K[w] ^ valid[w]; kany |= K[w]; }
            if (kany == 0u) {
                // nothing counted: only the denominators move
                if (t.tid == 0) {
                    if (!P.recount) { red_add(C + P.cap + 3, nvalid); red_add(C + P.cap + 5, exec_sum); }
                    red_add(C + 2 * P.cap + 8, phys);
                }
                continue;
            }
            if (kall != 0u) {
                // mixed group: re-decode with the erasure sums restricted to the counted frames
                unsigned long long e2;
                iters = mc_group<WORDS, DV, DC, true, TM>(P, t, bp, point, q0, valid, K, s_c2v, s_v2c, x, two, hbuf,
                                                          any, cnt, e2);
                phys += (unsigned long long)iters * W;
                second = true;
            }
        }
        // ---- flush the group's statistics (parallel_simulator_expurgated.py:239-243) ----
        team_sync(t);  // hbuf writes by thread 0 visible to the team
        for (int b = t.tid; b <= iters; b += t.nthr) red_add(C + b, hbuf[b]);
        if (t.tid == 0) {
            if (iters < P.cap) red_add(C + P.cap + 6 + iters + 1, cnt);  // bins iters+1 .. cap keep the residual
            uint32_t blk = 0, nk = 0;
            LDPC_FORW { blk += __popc(any[w] & K[w]); nk += __popc(K[w]); }
            red_add(C + P.cap + 1, blk);
            red_add(C + P.cap + 2, cnt);
            if (!P.recount) red_add(C + P.cap + 3, nvalid);   // the listing pass counted the denominators
            red_add(C + P.cap + 4, nk);
            if (!P.recount) red_add(C + P.cap + 5, exec_sum);
            red_add(C + 2 * P.cap + 8, phys);
            if (second) red_add(C + 2 * P.cap + 9, 1ull);
        }
        team_sync(t);  // hbuf is rewritten by the next group
    }
    if constexpr (TM) tm_free_all(tm_base);
}

// tail[b] is added to every bin >= b:  bins after a group's last physical iteration keep its residual
__global__ void e1_finalize_kernel(unsigned long long* counters, int npoints, int cap) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npoints) return;
    unsigned long long* C = counters + (size_t)p * raw_stride(cap);
    unsigned long long run = 0;
    for (int b = 0; b <= cap; b++) {
        run += C[cap + 6 + b];
        C[cap + 6 + b] = 0;
        C[b] += run;
    }
}

// ------------------------------------------------------------------------------------------------
// Batched decode (parity API and the single-frame drop-in): W = 32, value plane optional, per-frame
// errors[] / return value exactly as message_passing.c:7-82 produces them for zeroed `errors`.
// ------------------------------------------------------------------------------------------------
template <int DV, int DC, bool VALUES>
__global__ void __launch_bounds__(E1_MAX_THREADS, 1) e1_decode_kernel(const __grid_constant__ E1Params P) {
    constexpr int WORDS = 1;
    extern __shared__ __align__(16) unsigned char smem[];
    uint16_t* s_c2v = reinterpret_cast<uint16_t*>(smem + P.sm_c2v);
    uint16_t* s_v2c = reinterpret_cast<uint16_t*>(smem + P.sm_v2c);
    {
        const int dc = DC ? DC : P.dc, dv = DV ? DV : P.dv;
        const int nc = dc * P.m_pad / 8, nv = dv * P.n_pad / 8;
        const uint4* g0 = reinterpret_cast<const uint4*>(P.c2v);
        const uint4* g1 = reinterpret_cast<const uint4*>(P.v2c);
        uint4* d0 = reinterpret_cast<uint4*>(s_c2v);
        uint4* d1 = reinterpret_cast<uint4*>(s_v2c);
        for (int i = threadIdx.x; i < nc; i += blockDim.x) d0[i] = __ldg(g0 + i);
        for (int i = threadIdx.x; i < nv; i += blockDim.x) d1[i] = __ldg(g1 + i);
    }
    __syncthreads();

    TeamCtx t;
    t.team_warps = P.team_warps;
    t.nthr = P.team_warps * 32;
    const int team = threadIdx.x / t.nthr;
    t.tid = threadIdx.x - team * t.nthr;
    t.lane = threadIdx.x & 31;
    t.warp = t.tid >> 5;
    t.bar_id = 1 + team;
    unsigned char* base = smem + P.sm_team0 + (size_t)team * P.sm_team_stride;
    VecT<1>* x = reinterpret_cast<VecT<1>*>(base + P.sm_x);
    VecT<1>* two = reinterpret_cast<VecT<1>*>(base + P.sm_t);
    t.sa_x = (uint32_t)__cvta_generic_to_shared(x); t.sa_t = (uint32_t)__cvta_generic_to_shared(two); t.sm = smem;
    VecT<1>* val = reinterpret_cast<VecT<1>*>(base + P.sm_v);
    VecT<1>* par = reinterpret_cast<VecT<1>*>(base + P.sm_p);
    t.scr = reinterpret_cast<uint32_t*>(base + P.sm_scr);
    uint32_t* rf = t.scr + 32 * (2 * WORDS + 1) + 2;

    const unsigned long long groups = P.groups_per_point;
    for (unsigned long long gi = (unsigned long long)blockIdx.x * P.teams + team; gi < groups;
         gi += (unsigned long long)gridDim.x * P.teams) {
        const unsigned long long f0 = gi * 32ull;
        const uint32_t valid = P.num_frames - f0 >= 32ull ? 0xFFFFFFFFu : ((1u << (uint32_t)(P.num_frames - f0)) - 1u);
        if (P.mvc8) {
            const int nf = (int)min((unsigned long long)32, P.num_frames - f0);
            for (int i = t.tid; i < P.n; i += t.nthr) {
                uint32_t xi = 0u, vi = 0u;
                for (int b = 0; b < nf; b++) {
                    const int8_t sym = P.mvc8[(size_t)(f0 + b) * P.n + i];
                    xi |= (uint32_t)(sym == 2) << b;
                    vi |= (uint32_t)(sym == 1) << b;
                }
                x[i].a[0] = xi;
                if (VALUES) val[i].a[0] = vi;
            }
        } else {
            const uint32_t* gx = P.xin + (size_t)gi * P.n;
            const uint32_t* gv = P.vin + (size_t)gi * P.n;
            for (int i = t.tid; i < P.n; i += t.nthr) {
                const uint32_t xi = gx[i] & valid;
                x[i].a[0] = xi;
                if (VALUES) val[i].a[0] = gv[i] & ~xi & valid;
            }
        }
        team_sync(t);

        const unsigned long long frame = f0 + t.lane;
        const bool mine = (valid >> t.lane) & 1u;     // every warp of the team mirrors the 32 frames in its lanes
        int my_ret = P.cap;
        uint32_t running = valid;
        uint32_t my_count = 0;
        int it = 0;
        for (; it < P.cap; it++) {
            if (running == 0u) break;
            check_phase<1, DV, DC, VALUES>(P, t, s_c2v, x, two, val, par);
            if (t.tid < 32) rf[t.tid] = 0u;
            team_sync(t);
            uint32_t a2[1], c2[1], cnt, mask1[1] = {0xFFFFFFFFu};
            var_phase<1, DV, DC, VALUES, false>(P, t, s_v2c, x, two, val, par, mask1, a2, c2, cnt);
            uint32_t oc[2] = {a2[0], c2[0]};
            team_reduce<2>(t, oc, cnt);
            // per-frame erasure count after this iteration (message_passing.c:70-75)
            {
                uint32_t c = 0;
                for (int i = t.warp; i < P.n; i += t.team_warps) c += (x[i].a[0] >> t.lane) & 1u;
                if (c) atomicAdd(&rf[t.lane], c);
            }
            team_sync(t);
            my_count = rf[t.lane];
            if (mine && t.warp == 0) P.errors[frame * (unsigned long long)P.cap + it] = (int)my_count;
            if (((running >> t.lane) & 1u) && !((oc[0] >> t.lane) & 1u) && my_ret == P.cap) my_ret = it;  // :76-78
            running &= oc[0] & (it == 0 ? 0xFFFFFFFFu : oc[1]);  // :16-19
            team_sync(t);  // rf is zeroed again in the next iteration
        }
        // skipped iterations copy the previous count (:16-19); after success the caller's zeros stay
        if (mine && t.warp == 0) {
            for (int j = it; j < P.cap; j++) P.errors[frame * (unsigned long long)P.cap + j] = (int)my_count;
            P.ret[frame] = my_ret;
        }
        if (P.mvc8) {
            int8_t* out = const_cast<int8_t*>(P.mvc8);
            const int nf = (int)min((unsigned long long)32, P.num_frames - f0);
            for (int i = t.tid; i < P.n; i += t.nthr) {
                const uint32_t xw = x[i].a[0], vw = VALUES ? val[i].a[0] : 0u;
                for (int b = 0; b < nf; b++) out[(size_t)(f0 + b) * P.n + i] = ((xw >> b) & 1u) ? 2 : (int8_t)((vw >> b) & 1u);
            }
        } else {
            uint32_t* ox = P.xout + (size_t)gi * P.n;
            uint32_t* ov = P.vout + (size_t)gi * P.n;
            for (int i = t.tid; i < P.n; i += t.nthr) { ox[i] = x[i].a[0]; ov[i] = VALUES ? val[i].a[0] : 0u; }
        }
        team_sync(t);
    }
}

// Mvc int8 [frames][n] in {0,1,2}  ->  bit planes [groups][n] (bit b of word = frame 32g+b)
__global__ void pack_planes_kernel(const int8_t* __restrict__ mvc, uint32_t* __restrict__ xpl, uint32_t* __restrict__ vpl,
                                   unsigned long long frames, int n) {
    const unsigned long long g = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t xw = 0, vw = 0;
    for (int b = 0; b < 32; b++) {
        const unsigned long long f = g * 32ull + b;
        if (f < frames) {
            const int8_t s = mvc[f * (unsigned long long)n + i];
            xw |= (uint32_t)(s == 2) << b;
            vw |= (uint32_t)(s == 1) << b;
        }
    }
    xpl[g * (unsigned long long)n + i] = xw;
    vpl[g * (unsigned long long)n + i] = vw;
}

__global__ void unpack_planes_kernel(int8_t* __restrict__ mvc, const uint32_t* __restrict__ xpl,
                                     const uint32_t* __restrict__ vpl, unsigned long long frames, int n) {
    const unsigned long long g = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t xw = xpl[g * (unsigned long long)n + i], vw = vpl[g * (unsigned long long)n + i];
    for (int b = 0; b < 32; b++) {
        const unsigned long long f = g * 32ull + b;
        if (f < frames) mvc[f * (unsigned long long)n + i] = ((xw >> b) & 1u) ? 2 : (int8_t)((vw >> b) & 1u);
    }
}

// erasure patterns of the in-kernel generator, as bit planes [groups][n] (used by the parity tests
// to feed the reference C decoder exactly what the fused Monte-Carlo kernel decoded)
__global__ void draw_planes_kernel(uint32_t* __restrict__ xpl, BecParams bp, uint32_t point,
                                   unsigned long long first_word, unsigned long long words, int n) {
    const unsigned long long g = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || g >= words) return;
    xpl[g * (unsigned long long)n + i] = bec_word(bp, point, first_word + g, (uint32_t)i);
}

}  // namespace ldpc
