// e1_graphgen.cuh -- in-kernel generator of random simple (dv,dc)-regular Tanner graphs (ensemble mode).
//
// The reference draws a fresh graph for every frame (parallel_simulator_expurgated.py:200-223) with
// random_code_generator.c:21-67: shuffle the n*dv variable sockets, socket p of the permutation belongs to
// check p/dc and carries variable sequence[p]/dv (:34-36), reject graphs with a repeated variable in a
// check (:39-47).  Here every team of the E1 kernel builds the graph of its work item itself, in shared
// memory, from a counter-based key, so ensemble mode needs neither host cores nor PCIe:
//   1. keys:    key[s] = Philox4x32-10(counter (s>>2, G_lo, G_hi, 'GRAH'), key = graph seed)[s & 3]
//   2. permute: sockets in ascending (key, s) order (bucket sort in shared memory); position p carries socket s
//   3. repair:  every position (c, a) whose variable already occurs at an earlier socket of check c is
//               listed (ascending); one thread walks the list and swaps the socket with a pseudo-random
//               socket of another check when the swap leaves both checks repeat-free (same rule as
//               graphgen.cpp mode 1; rejection as in the reference would need ~36 000 shuffles for (4,8))
//   4. v2c:     checks of every variable via shared-memory atomics (any order: AND is commutative)
// graphgen.cpp mode 2 is the host twin of steps 1-3 (same keys, same order, same swaps): parity tests
// rebuild the graphs on the host and feed them to the reference decoder.
// (included by e1_kernels.cuh after E1Params / TeamCtx / team_sync are defined)
#pragma once
#include <stdint.h>
#include "bec_rng.cuh"

namespace ldpc {

constexpr uint32_t GG_TAG_KEYS = 0x47524148u;    // 'GRAH'
constexpr uint32_t GG_TAG_REPAIR = 0x52455052u;  // 'REPR'
constexpr int GG_MAX_CONFLICTS = 512;

// returns false if the conflict list overflowed (practically impossible: the expected number of repeats is
// (dv-1)(dc-1)/2); the caller then counts the group as failed-to-generate and skips it
template <int DV, int DC>
__device__ __forceinline__ bool team_generate_graph(const E1Params& P, const GraphGenParams& Gp, const TeamCtx& t,
                                                    unsigned char* team_base, unsigned long long G,
                                                    uint16_t* s_c2v, uint16_t* s_v2c) {
    const int dv = DV ? DV : P.dv, dc = DC ? DC : P.dc;
    const int E = P.n * dv;
    uint32_t* srt = reinterpret_cast<uint32_t*>(team_base + Gp.sm_sort);    // [E] (key high bits | socket id), sorted
    uint32_t* bkt = reinterpret_cast<uint32_t*>(team_base + Gp.sm_bkt);     // [nb] bucket counts -> ends
    uint32_t* part = bkt + (1u << Gp.nb_log2);                              // [nthr] scan partials
    uint32_t* fill = reinterpret_cast<uint32_t*>(team_base + Gp.sm_fill);
    uint32_t* conf = reinterpret_cast<uint32_t*>(team_base + Gp.sm_conf);   // [0] = count, [1..] = positions
    const uint32_t g_lo = (uint32_t)G, g_hi = (uint32_t)(G >> 32);
    const int nb = 1 << Gp.nb_log2, bsh = 32 - Gp.nb_log2;
    const uint32_t idmask = (1u << Gp.id_bits) - 1u;
    auto full_key = [&](uint32_t sock) {
        uint32_t r0 = sock >> 2, r1 = g_lo, r2 = g_hi, r3 = GG_TAG_KEYS;
        philox4x32_10(r0, r1, r2, r3, Gp.seed_lo, Gp.seed_hi);
        const uint32_t k = sock & 3u;
        return k == 0 ? r0 : k == 1 ? r1 : k == 2 ? r2 : r3;
    };

    // 2. the permutation = sockets in ascending (key, socket) order.  Bucket sort instead of a full sorting
    //    network: count the keys per bucket (top bits), prefix-sum, scatter (key high bits | socket) words,
    //    then every bucket (a handful of entries) is insertion-sorted by one thread.  Words that tie on the
    //    stored key bits are ordered by the full keys, so the order is exactly the (key, socket) order.
    for (int b = t.tid; b < nb; b += t.nthr) bkt[b] = 0u;
    if (t.tid == 0) conf[0] = 0u;
    team_sync(t);
    for (int b = t.tid; b < (E + 3) / 4; b += t.nthr) {
        uint32_t r0 = (uint32_t)b, r1 = g_lo, r2 = g_hi, r3 = GG_TAG_KEYS;
        philox4x32_10(r0, r1, r2, r3, Gp.seed_lo, Gp.seed_hi);
        const uint32_t r[4] = {r0, r1, r2, r3};
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (4 * b + k < E) atomicAdd(&bkt[r[k] >> bsh], 1u);
    }
    team_sync(t);
    {   // exclusive scan of the bucket counts, in place: chunk sums -> one warp scans them -> chunk prefixes
        const int L = (nb + t.nthr - 1) / t.nthr, lo = t.tid * L, hi = min(nb, lo + L);
        uint32_t sum = 0;
        for (int b = lo; b < hi; b++) sum += bkt[b];
        part[t.tid] = sum;
        team_sync(t);
        if (t.warp == 0) {
            const int per = (t.nthr + 31) / 32, p0 = t.lane * per, p1 = min(t.nthr, p0 + per);
            uint32_t mine = 0;
            for (int q = p0; q < p1; q++) mine += part[q];
            uint32_t inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (t.lane >= d) inc += o;
            }
            uint32_t run = inc - mine;
            for (int q = p0; q < p1; q++) { const uint32_t v = part[q]; part[q] = run; run += v; }
        }
        team_sync(t);
        uint32_t run = part[t.tid];
        for (int b = lo; b < hi; b++) { const uint32_t v = bkt[b]; bkt[b] = run; run += v; }
    }
    team_sync(t);
    for (int b = t.tid; b < (E + 3) / 4; b += t.nthr) {
        uint32_t r0 = (uint32_t)b, r1 = g_lo, r2 = g_hi, r3 = GG_TAG_KEYS;
        philox4x32_10(r0, r1, r2, r3, Gp.seed_lo, Gp.seed_hi);
        const uint32_t r[4] = {r0, r1, r2, r3};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int sck = 4 * b + k;
            if (sck < E) srt[atomicAdd(&bkt[r[k] >> bsh], 1u)] = (r[k] & ~idmask) | (uint32_t)sck;
        }
    }
    team_sync(t);   // bkt[b] is now the END of bucket b
    for (int b = t.tid; b < nb; b += t.nthr) {
        const int lo = b ? (int)bkt[b - 1] : 0, hi = (int)bkt[b];
        for (int i = lo + 1; i < hi; i++) {
            const uint32_t w = srt[i];
            int k = i - 1;
            while (k >= lo) {
                const uint32_t u = srt[k];
                bool greater;   // u > w in (key, socket) order?
                if (((u ^ w) & ~idmask) != 0u) greater = u > w;
                else {
                    const uint32_t ku = full_key(u & idmask), kw = full_key(w & idmask);
                    greater = ku > kw || (ku == kw && u > w);
                }
                if (!greater) break;
                srt[k + 1] = u;
                k--;
            }
            srt[k + 1] = w;
        }
    }
    team_sync(t);
    // check_lookup in the packed socket-list layout (socklist.hpp): position p -> check p/dc, socket p%dc,
    // variable (socket at position p)/dv
    for (int p = t.tid; p < E; p += t.nthr) s_c2v[sock_off(dc, P.m_pad, p % dc, p / dc)] = (uint16_t)((srt[p] & idmask) / dv);
    team_sync(t);
    // 3a. list the repeated positions
    for (int c = t.tid; c < P.m; c += t.nthr)
        for (int a = 1; a < dc; a++) {
            const uint16_t v = s_c2v[sock_off(dc, P.m_pad, a, c)];
            bool rep = false;
            for (int b = 0; b < a; b++) rep |= (s_c2v[sock_off(dc, P.m_pad, b, c)] == v);
            if (rep) {
                const uint32_t k = atomicAdd(&conf[0], 1u);
                if (k < GG_MAX_CONFLICTS) conf[1 + k] = (uint32_t)(c * dc + a);
            }
        }
    team_sync(t);
    const uint32_t nconf = conf[0];
    if (nconf > GG_MAX_CONFLICTS) return false;
    // The swap partner of a repeat's FIRST try depends only on its rank i in the list, not on the swaps made before it:
    // the team computes those in parallel (into the bucket array, free by now), so that the serial walk below only
    // runs Philox for the rare second and later tries (it used to be ~10 dependent Philox calls on one thread while
    // the rest of the team waited at the barrier).
    uint32_t* first_try = bkt;
    const uint32_t nfirst = min(nconf, (uint32_t)nb);
    for (uint32_t i = t.tid; i < nfirst; i += t.nthr) {
        uint32_t r0 = i << 16, r1 = g_lo, r2 = g_hi, r3 = GG_TAG_REPAIR;
        philox4x32_10(r0, r1, r2, r3, Gp.seed_lo, Gp.seed_hi);
        first_try[i] = __umulhi(r0, (uint32_t)E);
    }
    if (nconf) team_sync(t);   // team-uniform
    // 3b. one thread: ascending order, then the swaps
    if (t.tid == 0 && nconf) {
        for (uint32_t i = 1; i < nconf; i++) {   // insertion sort
            const uint32_t v = conf[1 + i];
            int k = (int)i - 1;
            while (k >= 0 && conf[1 + k] > v) { conf[2 + k] = conf[1 + k]; k--; }
            conf[2 + k] = v;
        }
        auto at = [&](int c, int j) -> uint16_t& { return s_c2v[sock_off(dc, P.m_pad, j, c)]; };
        auto holds = [&](int c, uint16_t v, int skip) {
            for (int j = 0; j < dc; j++) if (j != skip && at(c, j) == v) return true;
            return false;
        };
        for (uint32_t i = 0; i < nconf; i++) {
            const int c = (int)(conf[1 + i] / dc), a = (int)(conf[1 + i] % dc);
            const uint16_t v = at(c, a);
            bool rep = false;
            for (int b = 0; b < a; b++) rep |= (at(c, b) == v);
            if (!rep) continue;   // an earlier swap already removed it
            for (uint32_t tries = 0; tries < 65536u; tries++) {
                int s2;
                if (tries == 0u && i < nfirst) s2 = (int)first_try[i];
                else {
                    uint32_t r0 = (i << 16) | tries, r1 = g_lo, r2 = g_hi, r3 = GG_TAG_REPAIR;
                    philox4x32_10(r0, r1, r2, r3, Gp.seed_lo, Gp.seed_hi);
                    s2 = (int)__umulhi(r0, (uint32_t)E);
                }
                const int c2 = s2 / dc, j2 = s2 % dc;
                if (c2 == c) continue;
                const uint16_t v2 = at(c2, j2);
                if (holds(c, v2, a) || holds(c2, v, j2)) continue;
                at(c, a) = v2; at(c2, j2) = v;
                break;
            }
        }
    }
    // 4. variable side
    for (int i = t.tid; i < P.n; i += t.nthr) fill[i] = 0u;
    team_sync(t);
    for (int p = t.tid; p < E; p += t.nthr) {
        const int c = p / dc, j = p % dc;
        const int v = s_c2v[sock_off(dc, P.m_pad, j, c)];
        const uint32_t slot = atomicAdd(&fill[v], 1u);
        if (slot < (uint32_t)dv) s_v2c[sock_off(dv, P.n_pad, (int)slot, v)] = (uint16_t)c;
    }
    team_sync(t);
    return true;
}

}  // namespace ldpc
