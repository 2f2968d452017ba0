// e2_kernels.cuh -- "E2": HBM/L2-streamed bit-sliced BEC flooding decoder for graphs that do not fit
// in shared memory (BASELINE config 3: n = 10^6), sm_100a.
//
// Same algorithm as E1 (message_passing.c:7-82 collapsed to OR/AND on bit-sliced words), different
// memory plan:
//   * state is FRAME-MINOR in global memory: X[n][R], T[m][R] 32-bit words, R words = 32*R frames per
//     node row (R = 32 -> 128-byte rows, 1024 frames in lock step).  A thread owns one 128-bit quad
//     of a row; the 8 lanes of a quarter warp read one full 128-byte row: every gather is a perfectly
//     coalesced, vectorised ld.global.v4 of four whole 32-byte sectors.
//   * index arrays (int32, the reference layouts [m][dc], [n][dv]) are streamed once per iteration
//     and amortised over the 32*R frames of the batch.
//     They reach the SM through the TMA: every WARP takes tiles of E2_TILE consecutive nodes from a global work
//     counter and has the int32 lookups of its next tile copied into shared memory by one 1-D bulk copy
//     (cp.async.bulk.shared.global + mbarrier complete_tx, double-buffered per warp) while it gathers for the
//     current one, so a gather address never waits for a dependent global load -- the phase is ONE memory latency
//     deep instead of two, and tiles balance the warps to within 32 nodes.
//   * one COOPERATIVE persistent kernel per launch: check phase | grid.sync | variable phase |
//     grid.sync per iteration, termination decided on the device from OR-reduced masks.
// Bound: HBM bandwidth (state 1.5*n*R*4 B = 192 MB at n = 10^6, R = 32, larger than L2).
#pragma once
#include <cooperative_groups.h>
#include <stdint.h>
#include "bec_rng.cuh"
#include "e1_kernels.cuh"   // raw_stride, red_add

namespace ldpc {
namespace cg = cooperative_groups;

constexpr int E2_RMAX = 32;
constexpr int E2_THREADS = 256;
constexpr int E2_COUNT_SLICE = 4080;  // variables per thread in the per-frame count phase (16 * 255)
constexpr int E2_TILE = 32;           // nodes per streamed tile of lookups (one warp's tile)
constexpr int E2_TILE_MAXDEG = 8;     // tiles of 32 x 8 x 4 bytes = 1 KB, two per warp; larger degrees take the plain loops
constexpr int E2_TILE_BYTES = E2_TILE * E2_TILE_MAXDEG * 4;
constexpr int E2_TMA_SMEM = (E2_THREADS / 32) * 2 * E2_TILE_BYTES + (E2_THREADS / 32) * 2 * 8;   // tiles + mbarriers per block

struct E2Params {
    const int* c2v;   // [m][dc]   (padded with zeros to a whole number of tiles when tma != 0)
    const int* v2c;   // [n][dv]
    int async_depth;                // > 0: row gathers go through cp.async into a per-thread shared-memory ring of this
                                    // many items (dynamic shared memory: depth * ring_width * 16 B per thread)
    int ring_width;                 // rows per item in the ring = max(dc, dv + 1)
    int tma;                        // 1: lookups streamed as tiles through the TMA (cp.async.bulk)
    unsigned long long* tile_ctr;   // global tile work counter, zero at launch, never reset (phases use consecutive ranges)
    int n, m, dv, dc;
    int R;            // words per row (multiple of 4, <= 32)
    uint4* X;         // [n][R/4]
    uint4* T;         // [m][R/4]
    uint4* V;         // value plane [n][R/4]   (VALUES)
    uint4* PAR;       // [m][R/4]               (VALUES)
    // control block in global memory (uint32): acc[2][2*RMAX + 1]  (any[R], chg[R], cnt)
    uint32_t* acc;
    uint32_t* framecnt;     // [32*RMAX] per-frame erased counts (decode mode / expurgation)
    // Monte-Carlo
    const BecParams* bec;
    int npoints, cap, expurgation;
    uint32_t point0;   // erasure point p of this launch is keyed as point0 + p by the channel generator
    unsigned long long first_word, num_frames, batches_per_point;
    unsigned long long* counters;   // raw layout of e1_kernels.cuh
    // decode
    unsigned long long* hbuf;       // [cap+2] per-batch erasure sums, flushed when the batch counts
    // decode
    int* errors;   // [frames][cap]
    int* ret;      // [frames]
    unsigned long long frame_base;  // first frame of this batch in errors/ret
};

__device__ __forceinline__ uint32_t& q_at(uint4& v, int k) { return reinterpret_cast<uint32_t*>(&v)[k]; }

// ---- TMA tile streaming of the lookups ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA; completion is signalled on the mbarrier (complete_tx::bytes)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// per-WARP streaming state: every warp streams its own tiles (its own two buffers and mbarriers), so warps drift
// freely as in a plain grid-stride loop -- a block-wide tile (tried first) put two __syncthreads around every 64
// nodes and measured 9 % slower than no streaming at all
struct E2Stream {
    int* buf[2];          // shared: this warp's two tiles of lookups
    uint64_t* bar;        // shared: this warp's bar[2]
    uint32_t parity[2];   // phase parity of each buffer's mbarrier
    unsigned long long base;   // value of the global tile counter at which the current phase starts
};

// Runs body(node, lookups_of_node_in_shared_memory) for every node of tiles taken from the global counter.
// Lookups of the NEXT tile are in flight (TMA) while this one is processed; the id of the tile after that is being
// fetched (global atomic) meanwhile.  Every warp draws exactly two ids past the end, so a phase advances the counter
// by ntiles + 2 * (warps of the grid) and all warps know the next phase's base without communicating.
template <class F>
__device__ __forceinline__ void e2_stream_tiles(const int* gidx, int nodes, int d, int Q, unsigned long long* ctr, E2Stream& st, F&& body) {
    const long long ntiles = (nodes + E2_TILE - 1) / E2_TILE;
    const uint32_t tile_bytes = (uint32_t)(E2_TILE * d * 4);
    const int lane = (int)threadIdx.x & 31;
    const int per_step = 32 / Q;   // nodes handled per step (a thread owns one 128-bit quad of a row)
    const int nl0 = lane / Q;
    auto fetch = [&]() -> long long {   // warp-uniform id of the next tile
        unsigned long long v = 0;
        if (lane == 0) v = atomicAdd(ctr, 1ull);
        return (long long)(__shfl_sync(0xFFFFFFFFu, v, 0) - st.base);
    };
    auto issue = [&](long long tile, int b) {
        if (lane == 0) {
            mbar_expect_tx(&st.bar[b], tile_bytes);
            bulk_g2s(st.buf[b], gidx + (size_t)tile * E2_TILE * d, tile_bytes, &st.bar[b]);
        }
    };
    long long cur = fetch(), nxt = fetch();
    if (cur < ntiles) issue(cur, 0);
    for (int b = 0; cur < ntiles; b ^= 1) {
        __syncwarp();                          // every lane is done with buffer b^1 (read in the previous step)
        if (nxt < ntiles) issue(nxt, b ^ 1);
        unsigned long long pend = 0;           // id of the tile after the next: the atomic is in flight while this tile is
        if (lane == 0) pend = atomicAdd(ctr, 1ull);   // processed; its result is only looked at afterwards
        mbar_wait(&st.bar[b], st.parity[b]);
        st.parity[b] ^= 1u;
        const int* tile = st.buf[b];
        const int first = (int)cur * E2_TILE;
        for (int nl = nl0; nl < E2_TILE; nl += per_step) {
            const int node = first + nl;
            if (node < nodes) body(node, tile + nl * d);
        }
        cur = nxt; nxt = (long long)(__shfl_sync(0xFFFFFFFFu, pend, 0) - st.base);
    }
    st.base += (unsigned long long)ntiles + 2ull * gridDim.x * (blockDim.x >> 5);
}

// OR-reduce per word column and add-reduce cnt over the block, then one global atomic per column.
__device__ __forceinline__ void e2_block_accumulate(uint32_t* s_or /*[2*RMAX]*/, uint32_t* s_cnt, uint32_t* g_acc,
                                                    int R, int q, const uint4& any, const uint4& chg, uint32_t cnt) {
    // s_or zeroed by caller before the phase
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t a = reinterpret_cast<const uint32_t*>(&any)[k], c = reinterpret_cast<const uint32_t*>(&chg)[k];
        if (a) atomicOr(&s_or[q * 4 + k], a);
        if (c) atomicOr(&s_or[E2_RMAX + q * 4 + k], c);
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(s_cnt, cnt);
    __syncthreads();
    for (int k = threadIdx.x; k < R; k += blockDim.x) {
        if (s_or[k]) atomicOr(&g_acc[k], s_or[k]);
        if (s_or[E2_RMAX + k]) atomicOr(&g_acc[E2_RMAX + k], s_or[E2_RMAX + k]);
    }
    if (threadIdx.x == 0 && *s_cnt) atomicAdd(&g_acc[2 * E2_RMAX], *s_cnt);
}

// one check, one quad column: two = [>= 2 erased sockets] over the dc gathered rows
template <bool VALUES, bool SMEM_IDX>
__device__ __forceinline__ void e2_check_node(const E2Params& P, int Q, unsigned long long t, int q, const int* idx) {
    {
        uint4 one = make_uint4(0, 0, 0, 0), two = one, par = one;
        // sockets in batches of 8: all row gathers of a batch are in flight before the first is combined
        // (memory-level parallelism; a missing socket contributes the neutral word 0)
        for (int j0 = 0; j0 < P.dc; j0 += 8) {
            uint4 w[8], vv[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                w[u] = make_uint4(0, 0, 0, 0);
                if (VALUES) vv[u] = make_uint4(0, 0, 0, 0);
                if (j0 + u < P.dc) {
                    const int v = SMEM_IDX ? idx[j0 + u] : __ldg(idx + j0 + u);
                    w[u] = P.X[(size_t)v * Q + q];
                    if (VALUES) vv[u] = P.V[(size_t)v * Q + q];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                two.x |= one.x & w[u].x; two.y |= one.y & w[u].y; two.z |= one.z & w[u].z; two.w |= one.w & w[u].w;
                one.x |= w[u].x; one.y |= w[u].y; one.z |= w[u].z; one.w |= w[u].w;
                if (VALUES) { par.x ^= vv[u].x; par.y ^= vv[u].y; par.z ^= vv[u].z; par.w ^= vv[u].w; }
            }
        }
        P.T[t] = two;
        if (VALUES) P.PAR[t] = par;
    }
}

// ---- asynchronous row gathers (cp.async, erasure plane only) -------------------------------------------------------
// The plain loops keep the gathered rows of ONE node in registers (6 x 16 B per thread at dc = 6) and their round trip
// (lookups -> rows -> combine -> store, ~2.5 us) is what bounds the phase: 1024 threads x 96 B per SM in flight.
// Here every thread owns a ring of `depth` items in shared memory; the rows of item k+depth-1 are requested with
// cp.async.cg (16 B each, global -> shared, no registers held, L2-coherent) before item k is combined, so `depth` times
// as many bytes are in flight per thread.  Ring element (slot, j) of thread tid sits at ((slot*width + j)*blockDim + tid)
// * 16: consecutive threads, consecutive 16-byte words (conflict-free LDS.128).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int n) {   // n = groups that may stay pending (depth - 1, depth <= 4)
    if (n <= 0) cp_async_wait<0>(); else if (n == 1) cp_async_wait<1>(); else if (n == 2) cp_async_wait<2>(); else cp_async_wait<3>();
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

__device__ __forceinline__ void e2_check_phase_async(const E2Params& P, unsigned long long gtid, unsigned long long gthreads, uint32_t ring) {
    const int Q = P.R >> 2, dc = P.dc, D = P.async_depth;
    const unsigned long long total = (unsigned long long)P.m * Q;
    const int q = (int)(gtid % (unsigned long long)Q);   // gthreads is a multiple of Q: a thread keeps its quad column
    const uint32_t estride = blockDim.x * 16u, sstride = (uint32_t)P.ring_width * estride;
    auto issue = [&](unsigned long long t, int slot) {
        const int c = (int)(t / Q);
        const int* idx = P.c2v + (size_t)c * dc;
        const uint32_t base = ring + (uint32_t)slot * sstride;
        for (int j = 0; j < dc; j++) cp_async16(base + (uint32_t)j * estride, &P.X[(size_t)__ldg(idx + j) * Q + q]);
    };
    unsigned long long t_issue = gtid;
    int s_issue = 0, s_use = 0;
    for (int d = 0; d < D - 1; d++) {
        if (t_issue < total) issue(t_issue, s_issue);
        cp_async_commit();
        t_issue += gthreads; s_issue = s_issue + 1 == D ? 0 : s_issue + 1;
    }
    for (unsigned long long t = gtid; t < total; t += gthreads) {
        if (t_issue < total) issue(t_issue, s_issue);
        cp_async_commit();
        t_issue += gthreads; s_issue = s_issue + 1 == D ? 0 : s_issue + 1;
        cp_async_wait_dyn(D - 1);
        uint4 one = make_uint4(0, 0, 0, 0), two = one;
        const uint32_t base = ring + (uint32_t)s_use * sstride;
        for (int j = 0; j < dc; j++) {
            const uint4 w = lds128(base + (uint32_t)j * estride);
            two.x |= one.x & w.x; two.y |= one.y & w.y; two.z |= one.z & w.z; two.w |= one.w & w.w;
            one.x |= w.x; one.y |= w.y; one.z |= w.z; one.w |= w.w;
        }
        P.T[t] = two;
        s_use = s_use + 1 == D ? 0 : s_use + 1;
    }
    cp_async_wait<0>();
}

__device__ __forceinline__ void e2_var_phase_async(const E2Params& P, unsigned long long gtid, unsigned long long gthreads, uint32_t ring,
                                                   const uint32_t* s_mask, uint32_t* s_or, uint32_t* s_cnt, uint32_t* g_acc) {
    const int Q = P.R >> 2, dv = P.dv, D = P.async_depth;
    const unsigned long long total = (unsigned long long)P.n * Q;
    const int q = (int)(gtid % (unsigned long long)Q);
    const uint32_t estride = blockDim.x * 16u, sstride = (uint32_t)P.ring_width * estride;
    uint4 any = make_uint4(0, 0, 0, 0), chg = any;
    uint32_t cnt = 0;
    auto issue = [&](unsigned long long t, int slot) {   // element 0: the variable's own row, 1..dv: its checks' rows
        const int i = (int)(t / Q);
        const int* idx = P.v2c + (size_t)i * dv;
        const uint32_t base = ring + (uint32_t)slot * sstride;
        cp_async16(base, &P.X[t]);
        for (int j = 0; j < dv; j++) cp_async16(base + (uint32_t)(j + 1) * estride, &P.T[(size_t)__ldg(idx + j) * Q + q]);
    };
    unsigned long long t_issue = gtid;
    int s_issue = 0, s_use = 0;
    for (int d = 0; d < D - 1; d++) {
        if (t_issue < total) issue(t_issue, s_issue);
        cp_async_commit();
        t_issue += gthreads; s_issue = s_issue + 1 == D ? 0 : s_issue + 1;
    }
    for (unsigned long long t = gtid; t < total; t += gthreads) {
        if (t_issue < total) issue(t_issue, s_issue);
        cp_async_commit();
        t_issue += gthreads; s_issue = s_issue + 1 == D ? 0 : s_issue + 1;
        cp_async_wait_dyn(D - 1);
        const uint32_t base = ring + (uint32_t)s_use * sstride;
        const uint4 xi = lds128(base);
        uint4 xn = xi;
        for (int j = 0; j < dv; j++) {
            const uint4 tw = lds128(base + (uint32_t)(j + 1) * estride);
            xn.x &= tw.x; xn.y &= tw.y; xn.z &= tw.z; xn.w &= tw.w;
        }
        if (xn.x != xi.x || xn.y != xi.y || xn.z != xi.z || xn.w != xi.w) P.X[t] = xn;
        any.x |= xn.x; any.y |= xn.y; any.z |= xn.z; any.w |= xn.w;
        chg.x |= xi.x ^ xn.x; chg.y |= xi.y ^ xn.y; chg.z |= xi.z ^ xn.z; chg.w |= xi.w ^ xn.w;
        cnt += __popc(xn.x & s_mask[q * 4]) + __popc(xn.y & s_mask[q * 4 + 1]) + __popc(xn.z & s_mask[q * 4 + 2]) +
               __popc(xn.w & s_mask[q * 4 + 3]);
        s_use = s_use + 1 == D ? 0 : s_use + 1;
    }
    cp_async_wait<0>();
    e2_block_accumulate(s_or, s_cnt, g_acc, P.R, q, any, chg, cnt);
}

template <bool VALUES>
__device__ __forceinline__ void e2_check_phase(const E2Params& P, unsigned long long gtid, unsigned long long gthreads, E2Stream& st) {
    const int Q = P.R >> 2;
    if (P.tma) {
        const int q = (int)threadIdx.x % Q;
        e2_stream_tiles(P.c2v, P.m, P.dc, Q, P.tile_ctr, st, [&](int c, const int* idx) {
            e2_check_node<VALUES, true>(P, Q, (unsigned long long)c * Q + q, q, idx);
        });
        return;
    }
    const unsigned long long total = (unsigned long long)P.m * Q;
    for (unsigned long long t = gtid; t < total; t += gthreads) {
        const int c = (int)(t / Q), q = (int)(t - (unsigned long long)c * Q);
        e2_check_node<VALUES, false>(P, Q, t, q, P.c2v + (size_t)c * P.dc);
    }
}

// one variable, one quad column: x &= AND of the dv gathered `two` rows; accumulates any / chg / cnt
template <bool VALUES, bool SMEM_IDX>
__device__ __forceinline__ void e2_var_node(const E2Params& P, int Q, unsigned long long t, int q, const int* idx,
                                            const uint32_t* s_mask, uint4& any, uint4& chg, uint32_t& cnt) {
    {
        const uint4 xi = P.X[t];
        uint4 xn = xi, nv = make_uint4(0, 0, 0, 0);
        for (int j0 = 0; j0 < P.dv; j0 += 4) {   // batches of 4 gathers in flight; neutral word = all ones
            uint4 tw[4], pr[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                tw[u] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                if (VALUES) pr[u] = make_uint4(0, 0, 0, 0);
                if (j0 + u < P.dv) {
                    const int c = SMEM_IDX ? idx[j0 + u] : __ldg(idx + j0 + u);
                    tw[u] = P.T[(size_t)c * Q + q];
                    if (VALUES) pr[u] = P.PAR[(size_t)c * Q + q];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                xn.x &= tw[u].x; xn.y &= tw[u].y; xn.z &= tw[u].z; xn.w &= tw[u].w;
                if (VALUES) {   // listed order, last writer wins (message_passing.c:56-61)
                    nv.x = (nv.x & tw[u].x) | (pr[u].x & ~tw[u].x); nv.y = (nv.y & tw[u].y) | (pr[u].y & ~tw[u].y);
                    nv.z = (nv.z & tw[u].z) | (pr[u].z & ~tw[u].z); nv.w = (nv.w & tw[u].w) | (pr[u].w & ~tw[u].w);
                }
            }
        }
        if (xn.x != xi.x || xn.y != xi.y || xn.z != xi.z || xn.w != xi.w) {
            P.X[t] = xn;
            if (VALUES) {
                uint4 vi = P.V[t];
                vi.x |= nv.x & xi.x & ~xn.x; vi.y |= nv.y & xi.y & ~xn.y; vi.z |= nv.z & xi.z & ~xn.z; vi.w |= nv.w & xi.w & ~xn.w;
                P.V[t] = vi;
            }
        }
        any.x |= xn.x; any.y |= xn.y; any.z |= xn.z; any.w |= xn.w;
        chg.x |= xi.x ^ xn.x; chg.y |= xi.y ^ xn.y; chg.z |= xi.z ^ xn.z; chg.w |= xi.w ^ xn.w;
        cnt += __popc(xn.x & s_mask[q * 4]) + __popc(xn.y & s_mask[q * 4 + 1]) + __popc(xn.z & s_mask[q * 4 + 2]) +
               __popc(xn.w & s_mask[q * 4 + 3]);
    }
}

template <bool VALUES>
__device__ __forceinline__ void e2_var_phase(const E2Params& P, unsigned long long gtid, unsigned long long gthreads,
                                             const uint32_t* s_mask, uint32_t* s_or, uint32_t* s_cnt, uint32_t* g_acc, E2Stream& st) {
    const int Q = P.R >> 2;
    uint4 any = make_uint4(0, 0, 0, 0), chg = any;
    uint32_t cnt = 0;
    int myq = 0;
    if (P.tma) {
        myq = (int)threadIdx.x % Q;   // blockDim.x is a multiple of Q: a thread keeps its quad column
        e2_stream_tiles(P.v2c, P.n, P.dv, Q, P.tile_ctr, st, [&](int i, const int* idx) {
            e2_var_node<VALUES, true>(P, Q, (unsigned long long)i * Q + myq, myq, idx, s_mask, any, chg, cnt);
        });
    } else {
        const unsigned long long total = (unsigned long long)P.n * Q;
        // gthreads is a multiple of Q (host guarantees), so a thread keeps the same quad column
        for (unsigned long long t = gtid; t < total; t += gthreads) {
            const int i = (int)(t / Q), q = (int)(t - (unsigned long long)i * Q);
            myq = q;
            e2_var_node<VALUES, false>(P, Q, t, q, P.v2c + (size_t)i * P.dv, s_mask, any, chg, cnt);
        }
    }
    e2_block_accumulate(s_or, s_cnt, g_acc, P.R, myq, any, chg, cnt);
}

// per-frame erased counts of the current state -> framecnt[32*R] (atomicAdd; zeroed beforehand)
__device__ __forceinline__ void e2_count_phase(const E2Params& P, unsigned long long gtid, unsigned long long gthreads) {
    const uint32_t* X = reinterpret_cast<const uint32_t*>(P.X);
    const int slices = (P.n + E2_COUNT_SLICE - 1) / E2_COUNT_SLICE;
    const unsigned long long total = (unsigned long long)slices * P.R;
    for (unsigned long long t = gtid; t < total; t += gthreads) {
        const int s = (int)(t / P.R), w = (int)(t - (unsigned long long)s * P.R);
        uint32_t cnt[32];
#pragma unroll
        for (int b = 0; b < 32; b++) cnt[b] = 0;
        const int i1 = min(P.n, (s + 1) * E2_COUNT_SLICE);
        for (int i0 = s * E2_COUNT_SLICE; i0 < i1; i0 += 255) {
            uint32_t by[8];
#pragma unroll
            for (int k = 0; k < 8; k++) by[k] = 0;
            const int ie = min(i1, i0 + 255);
            for (int i = i0; i < ie; i++) {
                const uint32_t xw = X[(size_t)i * P.R + w];
#pragma unroll
                for (int k = 0; k < 8; k++) by[k] += (xw >> k) & 0x01010101u;   // bit k+8j -> byte j
            }
#pragma unroll
            for (int k = 0; k < 8; k++)
#pragma unroll
                for (int j = 0; j < 4; j++) cnt[k + 8 * j] += (by[k] >> (8 * j)) & 0xFFu;
        }
#pragma unroll
        for (int b = 0; b < 32; b++)
            if (cnt[b]) atomicAdd(&P.framecnt[w * 32 + b], cnt[b]);
    }
}

template <bool VALUES, bool DECODE>
__global__ void __launch_bounds__(E2_THREADS, 4) e2_kernel(const __grid_constant__ E2Params P) {
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t s_or[2 * E2_RMAX];
    __shared__ uint32_t s_cnt;
    __shared__ uint32_t s_mask[E2_RMAX];     // frames whose erasures are summed (counted frames)
    __shared__ uint32_t s_running[E2_RMAX];
    __shared__ uint32_t s_valid[E2_RMAX];
    __shared__ uint32_t s_any[E2_RMAX];
    __shared__ uint32_t s_flag;
    // dynamic shared memory: the cp.async ring (async_depth > 0) OR the per-warp lookup tiles + mbarriers (tma)
    extern __shared__ __align__(128) unsigned char e2_dyn[];
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(e2_dyn) + threadIdx.x * 16u;
    E2Stream st;
    st.parity[0] = 0u; st.parity[1] = 0u; st.base = 0ull;
    st.buf[0] = st.buf[1] = nullptr; st.bar = nullptr;
    if (P.tma) {
        const int wb = (int)threadIdx.x >> 5;
        int* tiles = reinterpret_cast<int*>(e2_dyn);
        st.buf[0] = tiles + (size_t)(wb * 2) * (E2_TILE_BYTES / 4); st.buf[1] = st.buf[0] + E2_TILE_BYTES / 4;
        st.bar = reinterpret_cast<uint64_t*>(e2_dyn + (size_t)(E2_THREADS / 32) * 2 * E2_TILE_BYTES) + wb * 2;
        if ((threadIdx.x & 31) == 0) {
            mbar_init(&st.bar[0], 1u); mbar_init(&st.bar[1], 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    const int R = P.R, Q = R >> 2;
    const int ACC = 2 * E2_RMAX + 1;
    const unsigned long long gthreads = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool lead = (blockIdx.x == 0 && threadIdx.x == 0);
    const int stride = raw_stride(P.cap);
    const unsigned long long items = DECODE ? 1ull : (unsigned long long)P.npoints * P.batches_per_point;
    int par = 0;   // accumulator buffer the next phase adds into (host zeroes both before the launch)

    auto zero_block_acc = [&]() {
        __syncthreads();
        for (int k = threadIdx.x; k < 2 * E2_RMAX; k += blockDim.x) s_or[k] = 0;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
    };
    auto zero_global_acc = [&](int which) {   // block 0 only
        if (blockIdx.x == 0)
            for (int k = threadIdx.x; k < ACC; k += blockDim.x) P.acc[which * ACC + k] = 0;
    };

    for (unsigned long long item = 0; item < items; item++) {
        const uint32_t point = DECODE ? 0u : (uint32_t)(item / P.batches_per_point);
        const unsigned long long batch = DECODE ? 0ull : item - (unsigned long long)point * P.batches_per_point;
        const unsigned long long q0 = P.first_word + batch * R;
        unsigned long long* C = DECODE ? nullptr : P.counters + (size_t)point * stride;
        __syncthreads();
        for (int w = threadIdx.x; w < R; w += blockDim.x) {
            const unsigned long long f0 = (batch * R + w) * 32ull;
            const uint32_t v = f0 >= P.num_frames ? 0u
                               : (P.num_frames - f0 >= 32ull ? 0xFFFFFFFFu : ((1u << (uint32_t)(P.num_frames - f0)) - 1u));
            s_valid[w] = v; s_mask[w] = v;
        }
        __syncthreads();
        if (DECODE && blockIdx.x == 0)
            for (int f = threadIdx.x; f < 32 * R; f += blockDim.x)
                if ((s_valid[f >> 5] >> (f & 31)) & 1u) P.ret[P.frame_base + f] = P.cap;

        bool second_pass = false;
        unsigned long long exec_sum = 0, phys = 0;
        for (int pass = 0; pass < 2; pass++) {
            // ---- initial state (channels.py:24-26 drawn in place) and its erased count
            zero_block_acc();
            if (!DECODE) {
                const BecParams bp = P.bec[point];
                uint4 any = make_uint4(0, 0, 0, 0), chg = any;
                uint32_t cnt0 = 0;
                int myq = 0;
                const unsigned long long total = (unsigned long long)P.n * Q;
                for (unsigned long long t = gtid; t < total; t += gthreads) {
                    const int i = (int)(t / Q), q = (int)(t - (unsigned long long)i * Q);
                    myq = q;
                    uint4 xi;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t wv = bec_word(bp, point + P.point0, q0 + q * 4 + k, (uint32_t)i) & s_valid[q * 4 + k];
                        q_at(xi, k) = wv;
                        q_at(any, k) |= wv;
                        cnt0 += __popc(wv & s_mask[q * 4 + k]);
                    }
                    P.X[t] = xi;
                }
                e2_block_accumulate(s_or, &s_cnt, P.acc + par * ACC, R, myq, any, chg, cnt0);
            }
            zero_global_acc(par ^ 1);
            // framecnt is read by every block when the expurgation mask is formed at the end of pass 0,
            // so it is only cleared at the start of an item (the previous item ended with a grid.sync)
            if (blockIdx.x == 0 && pass == 0) for (int f = threadIdx.x; f < 32 * E2_RMAX; f += blockDim.x) P.framecnt[f] = 0;
            grid.sync();
            uint32_t cur_cnt = DECODE ? 0u : __ldcg(&P.acc[par * ACC + 2 * E2_RMAX]);   // erased count of the summed frames
            __syncthreads();
            for (int w = threadIdx.x; w < R; w += blockDim.x) {
                s_any[w] = DECODE ? s_valid[w] : __ldcg(&P.acc[par * ACC + w]);
                s_running[w] = pass == 0 ? s_valid[w] : s_mask[w];
            }
            if (!DECODE && lead) P.hbuf[0] = cur_cnt;
            __syncthreads();
            par ^= 1;

            int it = 0;
            for (; it < P.cap; it++) {
                if (threadIdx.x == 0) { uint32_t r = 0; for (int w = 0; w < R; w++) r |= s_running[w]; s_flag = r; }
                __syncthreads();
                if (s_flag == 0u) break;
                if (pass == 0) for (int w = 0; w < R; w++) exec_sum += __popc(s_running[w]);
                if (DECODE && blockIdx.x == 0)   // only block 0 reads framecnt, and it is past those reads
                    for (int f = threadIdx.x; f < 32 * E2_RMAX; f += blockDim.x) P.framecnt[f] = 0;
                if (!VALUES && P.async_depth > 0) e2_check_phase_async(P, gtid, gthreads, ring);
                else e2_check_phase<VALUES>(P, gtid, gthreads, st);
                zero_block_acc();
                grid.sync();
                if (!VALUES && P.async_depth > 0) e2_var_phase_async(P, gtid, gthreads, ring, s_mask, s_or, &s_cnt, P.acc + par * ACC);
                else e2_var_phase<VALUES>(P, gtid, gthreads, s_mask, s_or, &s_cnt, P.acc + par * ACC, st);
                zero_global_acc(par ^ 1);   // everyone finished reading it before the sync above
                grid.sync();
                if (DECODE) {
                    e2_count_phase(P, gtid, gthreads);
                    grid.sync();
                }
                const uint32_t* A = P.acc + par * ACC;
                cur_cnt = __ldcg(&A[2 * E2_RMAX]);
                if (!DECODE && lead) P.hbuf[it + 1] = cur_cnt;
                if (DECODE && blockIdx.x == 0) {
                    for (int f = threadIdx.x; f < 32 * R; f += blockDim.x) {
                        const int w = f >> 5, b = f & 31;
                        if ((s_valid[w] >> b) & 1u) {
                            const unsigned long long fr = P.frame_base + f;
                            P.errors[fr * (unsigned long long)P.cap + it] = (int)__ldcg(&P.framecnt[f]);      // :70-75
                            if (((s_running[w] >> b) & 1u) && !((__ldcg(&A[w]) >> b) & 1u) && P.ret[fr] == P.cap) P.ret[fr] = it;  // :76-78
                        }
                    }
                }
                __syncthreads();
                for (int w = threadIdx.x; w < R; w += blockDim.x) {
                    s_any[w] = __ldcg(&A[w]);
                    s_running[w] &= __ldcg(&A[w]) & (it == 0 ? 0xFFFFFFFFu : __ldcg(&A[E2_RMAX + w]));   // message_passing.c:16-19
                }
                __syncthreads();
                par ^= 1;
            }
            phys += (unsigned long long)it * 32ull * R;

            if (DECODE) {
                // skipped iterations copy the last count (:16-19); decoded frames keep the caller's zeros
                if (blockIdx.x == 0) {
                    for (int f = threadIdx.x; f < 32 * R; f += blockDim.x) {
                        const int w = f >> 5, b = f & 31;
                        if ((s_valid[w] >> b) & 1u) {
                            const unsigned long long fr = P.frame_base + f;
                            const int c = it > 0 ? (int)__ldcg(&P.framecnt[f]) : 0;
                            for (int j = it; j < P.cap; j++) P.errors[fr * (unsigned long long)P.cap + j] = c;
                        }
                    }
                }
                break;
            }

            // ---- expurgation predicate (parallel_simulator_expurgated.py:238), decided identically by
            //      every block from global values
            if (pass == 0 && P.expurgation >= 0) {
                e2_count_phase(P, gtid, gthreads);
                grid.sync();
                uint32_t kall = 0, kany = 0;
                __syncthreads();
                for (int w = 0; w < R; w++) {
                    uint32_t k = 0;
                    for (int b = 0; b < 32; b++) k |= (uint32_t)((int)__ldcg(&P.framecnt[w * 32 + b]) > P.expurgation) << b;
                    k &= s_valid[w];
                    kall |= k ^ s_valid[w]; kany |= k;
                    if (threadIdx.x == 0) s_mask[w] = k;
                }
                __syncthreads();
                if (kany == 0u) {
                    if (lead) {
                        uint32_t nv = 0; for (int w = 0; w < R; w++) nv += __popc(s_valid[w]);
                        red_add(C + P.cap + 3, nv); red_add(C + P.cap + 5, exec_sum); red_add(C + 2 * P.cap + 8, phys);
                    }
                    break;
                }
                if (kall != 0u) { second_pass = true; continue; }   // mixed batch: re-decode with masked sums
            }
            if (lead) {
                for (int b = 0; b <= it; b++) red_add(C + b, __ldcg(&P.hbuf[b]));
                if (it < P.cap) red_add(C + P.cap + 6 + it + 1, cur_cnt);   // later bins keep the residual
                uint32_t blk = 0, nk = 0, nv = 0;
                for (int w = 0; w < R; w++) { blk += __popc(s_any[w] & s_mask[w]); nk += __popc(s_mask[w]); nv += __popc(s_valid[w]); }
                red_add(C + P.cap + 1, blk);
                red_add(C + P.cap + 2, cur_cnt);
                red_add(C + P.cap + 3, nv); red_add(C + P.cap + 4, nk); red_add(C + P.cap + 5, exec_sum);
                red_add(C + 2 * P.cap + 8, phys);
                if (second_pass) red_add(C + 2 * P.cap + 9, 1ull);
            }
            break;
        }
        grid.sync();   // state buffers, framecnt and hbuf are reused by the next item
    }
}

}  // namespace ldpc
