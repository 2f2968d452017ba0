// e1_graphgen.cuh -- in-kernel generator of random simple (dv,dc)-regular Tanner graphs (ensemble mode).
//
// The reference draws a fresh graph for every frame (parallel_simulator_expurgated.py:200-223) with
// random_code_generator.c:21-67: shuffle the n*dv variable sockets, socket p of the permutation belongs to
// check p/dc and carries variable sequence[p]/dv (:34-36), reject graphs with a repeated variable in a
// check (:39-47).  Here every team of the E1 kernel builds the graph of its work item itself, in shared
// memory, from a counter-based key, so ensemble mode needs neither host cores nor PCIe:
//   1. keys:    key[s] = Philox4x32-10(counter (s>>2, G_lo, G_hi, 'GRAH'), key = graph seed)[s & 3]
//   2. permute: bitonic sort of (key, s) in shared memory; position p of the sorted order carries socket s
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
    const int E = P.n * dv, epad = Gp.epad;
    uint32_t* keys = reinterpret_cast<uint32_t*>(team_base + Gp.sm_keys);
    uint16_t* vals = reinterpret_cast<uint16_t*>(team_base + Gp.sm_vals);
    uint32_t* fill = reinterpret_cast<uint32_t*>(team_base + Gp.sm_fill);
    uint32_t* conf = reinterpret_cast<uint32_t*>(team_base + Gp.sm_conf);   // [0] = count, [1..] = positions
    const uint32_t g_lo = (uint32_t)G, g_hi = (uint32_t)(G >> 32);

    // 1. keys
    for (int b = t.tid; b < epad / 4; b += t.nthr) {
        uint32_t r0 = (uint32_t)b, r1 = g_lo, r2 = g_hi, r3 = GG_TAG_KEYS;
        philox4x32_10(r0, r1, r2, r3, Gp.seed_lo, Gp.seed_hi);
        const uint32_t r[4] = {r0, r1, r2, r3};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int s = 4 * b + k;
            keys[s] = s < E ? r[k] : 0xFFFFFFFFu;
            vals[s] = s < E ? (uint16_t)s : (uint16_t)0xFFFFu;
        }
    }
    if (t.tid == 0) conf[0] = 0u;
    team_sync(t);
    // 2. bitonic sort by (key, val)
    for (int k = 2; k <= epad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t.tid; i < epad / 2; i += t.nthr) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));   // element without bit j
                const int hi = lo | j;
                const bool up = (lo & k) == 0;
                const uint32_t ka = keys[lo], kb = keys[hi];
                const uint16_t va = vals[lo], vb = vals[hi];
                const bool gt = ka > kb || (ka == kb && va > vb);
                if (gt == up) { keys[lo] = kb; keys[hi] = ka; vals[lo] = vb; vals[hi] = va; }
            }
            team_sync(t);
        }
    // check_lookup in the packed socket-list layout (socklist.hpp): position p -> check p/dc, socket p%dc, variable vals[p]/dv
    for (int p = t.tid; p < E; p += t.nthr) s_c2v[sock_off(dc, P.m_pad, p % dc, p / dc)] = (uint16_t)(vals[p] / dv);
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
                uint32_t r0 = (i << 16) | tries, r1 = g_lo, r2 = g_hi, r3 = GG_TAG_REPAIR;
                philox4x32_10(r0, r1, r2, r3, Gp.seed_lo, Gp.seed_hi);
                const int s2 = (int)__umulhi(r0, (uint32_t)E);
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
