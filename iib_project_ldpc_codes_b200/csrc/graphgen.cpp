// graphgen.cpp -- host-side random (dv,dc)-regular Tanner-graph generators.
//
// The graph is built on the CPU in the reference too (random_code_generator.c:21-67, called from
// parallel_simulator_expurgated.py:203-227); what the GPU path consumes is its output contract:
//   check_lookup[c*dc + j]    = variable at socket j of check c, in permutation order (unsorted)  (:34-36)
//   variable_lookup[i*dv + t] = checks containing variable i in ascending check order            (:53-64)
// No check may contain a variable twice (:39-47).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <vector>
#include <algorithm>
#include <utility>

#include "../../include/ldpc_b200.h"

namespace {

// true iff some check holds the same variable in two sockets
bool has_repeat(const int* check_lookup, int m, int dc) {
    for (int c = 0; c < m; c++) {
        const int* row = check_lookup + (size_t)c * dc;
        for (int a = 1; a < dc; a++)
            for (int b = 0; b < a; b++)
                if (row[a] == row[b]) return true;
    }
    return false;
}

// checks of every variable in ascending check order, without the dense matrix
void build_variable_lookup(const int* check_lookup, int n, int m, int dv, int dc, int* variable_lookup) {
    std::vector<int> fill((size_t)n, 0);
    for (int c = 0; c < m; c++)
        for (int j = 0; j < dc; j++) {
            const int v = check_lookup[(size_t)c * dc + j];
            if (fill[v] < dv) variable_lookup[(size_t)v * dv + fill[v]] = c;
            fill[v]++;
        }
}

struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    // unbiased integer in [0, bound)
    uint32_t below(uint32_t bound) {
        uint64_t x = next() >> 32, mres = x * bound;
        uint32_t lo = (uint32_t)mres;
        if (lo < bound) {
            const uint32_t thresh = (0u - bound) % bound;
            while (lo < thresh) { x = next() >> 32; mres = x * bound; lo = (uint32_t)mres; }
        }
        return (uint32_t)(mres >> 32);
    }
};

// Philox4x32-10 (host twin of bec_rng.cuh)
void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int round = 0; round < 10; round++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// Host twin of the in-kernel generator (e1_graphgen.cuh): keys -> sort -> listed repeats repaired in
// ascending order with Philox-drawn swap partners.  `seed` = graph seed, `G` = graph index.
int generate_device_compatible(int n, int dv, int dc, uint64_t seed, uint64_t G, int* check_lookup) {
    const int E = n * dv, m = E / dc;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), g_lo = (uint32_t)G, g_hi = (uint32_t)(G >> 32);
    std::vector<std::pair<uint32_t, int>> ks((size_t)E);
    for (int b = 0; 4 * b < E; b++) {
        uint32_t r[4] = {(uint32_t)b, g_lo, g_hi, 0x47524148u};
        philox4x32_10(r, k0, k1);
        for (int k = 0; k < 4 && 4 * b + k < E; k++) ks[(size_t)4 * b + k] = {r[k], 4 * b + k};
    }
    std::sort(ks.begin(), ks.end());
    for (int p = 0; p < E; p++) check_lookup[p] = ks[(size_t)p].second / dv;
    std::vector<int> conf;
    for (int c = 0; c < m; c++)
        for (int a = 1; a < dc; a++) {
            bool rep = false;
            for (int b = 0; b < a; b++) rep |= (check_lookup[(size_t)c * dc + b] == check_lookup[(size_t)c * dc + a]);
            if (rep) conf.push_back(c * dc + a);
        }
    if (conf.size() > 512) return 0;
    auto holds = [&](int c, int v, int skip) {
        for (int j = 0; j < dc; j++) if (j != skip && check_lookup[(size_t)c * dc + j] == v) return true;
        return false;
    };
    for (size_t i = 0; i < conf.size(); i++) {
        const int c = conf[i] / dc, a = conf[i] % dc;
        const int v = check_lookup[(size_t)c * dc + a];
        bool rep = false;
        for (int b = 0; b < a; b++) rep |= (check_lookup[(size_t)c * dc + b] == v);
        if (!rep) continue;
        for (uint32_t tries = 0; tries < 65536u; tries++) {
            uint32_t r[4] = {((uint32_t)i << 16) | tries, g_lo, g_hi, 0x52455052u};
            philox4x32_10(r, k0, k1);
            const int s2 = (int)(((uint64_t)r[0] * (uint32_t)E) >> 32);
            const int c2 = s2 / dc, j2 = s2 % dc;
            if (c2 == c) continue;
            const int v2 = check_lookup[s2];
            if (holds(c, v2, a) || holds(c2, v, j2)) continue;
            check_lookup[(size_t)c * dc + a] = v2;
            check_lookup[s2] = v;
            break;
        }
    }
    return 1;
}

}  // namespace

// Drop-in for random_code_generator.c:21.  Same libc rand() consumption (Fisher-Yates from the top,
// `rand() % (i+1)`, :13-19), same retry budget (an attempt with retry index > 10000 gives up, :26-28),
// `sequence` shuffled in place and kept across retries, `seed` ignored exactly like the reference.
extern "C" int generate_random_code(int n, int dv, int dc, int* variable_lookup, int* check_lookup, int* sequence,
                                    bool* parity_check, int iterations, bool first_run, int seed) {
    (void)seed;
    if (first_run) srand((unsigned)time(NULL));
    if (n <= 0 || dv <= 0 || dc <= 0 || !variable_lookup || !check_lookup || !sequence) return 0;
    const long long E = (long long)n * dv;
    const int k = (int)((long long)n * (dc - dv) / dc);
    const int m = n - k;
    for (int attempt = iterations; attempt <= 10000; attempt++) {
        for (long long i = E - 1; i > 0; i--) {
            const long long j = rand() % (i + 1);
            const int tmp = sequence[i]; sequence[i] = sequence[j]; sequence[j] = tmp;
        }
        for (long long e = 0; e < E; e++) check_lookup[e] = sequence[e] / dv;
        if (has_repeat(check_lookup, m, dc)) continue;
        build_variable_lookup(check_lookup, n, m, dv, dc, variable_lookup);
        if (parity_check)
            for (int c = 0; c < m; c++)
                for (int j = 0; j < dc; j++) parity_check[(size_t)c * n + check_lookup[(size_t)c * dc + j]] = true;
        return 1;
    }
    return 0;
}

// Seeded generator, same output contract.  mode 0: permutation + rejection (the reference's ensemble);
// mode 1: permutation, then repeated variables are removed by swapping the offending socket with a
// uniformly chosen socket of another check when the swap leaves both checks repeat-free.
extern "C" int ldpc_b200_generate_graph(int n, int dv, int dc, uint64_t seed, int mode, int* variable_lookup,
                                        int* check_lookup) {
    if (n <= 0 || dv <= 0 || dc <= 0 || !variable_lookup || !check_lookup) return 0;
    const long long E = (long long)n * dv;
    if (E % dc != 0 || E > 0x7FFFFFFFll) return 0;
    const int m = (int)(E / dc);
    if (dc > n) return 0;  // a repeat-free check needs dc distinct variables
    SplitMix64 rng(seed * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
    std::vector<int> seq((size_t)E);
    const long long max_attempts = mode == 0 ? 2000000 : 1;
    for (long long attempt = 0; attempt < max_attempts; attempt++) {
        for (long long e = 0; e < E; e++) seq[(size_t)e] = (int)e;
        for (long long i = E - 1; i > 0; i--) {
            const uint32_t j = rng.below((uint32_t)(i + 1));
            const int tmp = seq[(size_t)i]; seq[(size_t)i] = seq[j]; seq[j] = tmp;
        }
        for (long long e = 0; e < E; e++) check_lookup[e] = seq[(size_t)e] / dv;
        if (mode == 0) {
            if (has_repeat(check_lookup, m, dc)) continue;
        } else {
            auto holds = [&](int c, int v, int skip_socket) {
                for (int j = 0; j < dc; j++)
                    if (j != skip_socket && check_lookup[(size_t)c * dc + j] == v) return true;
                return false;
            };
            long long guard = 0;
            bool dirty = true;
            while (dirty) {
                dirty = false;
                for (int c = 0; c < m; c++)
                    for (int a = 1; a < dc; a++) {
                        const int v = check_lookup[(size_t)c * dc + a];
                        bool rep = false;
                        for (int b = 0; b < a; b++) rep |= (check_lookup[(size_t)c * dc + b] == v);
                        if (!rep) continue;
                        for (;;) {
                            if (++guard > 64 * E + 1000000) return 0;
                            const uint32_t s2 = rng.below((uint32_t)E);
                            const int c2 = (int)(s2 / dc), j2 = (int)(s2 % dc);
                            if (c2 == c) continue;
                            const int v2 = check_lookup[s2];
                            if (holds(c, v2, a) || holds(c2, v, j2)) continue;
                            check_lookup[(size_t)c * dc + a] = v2;
                            check_lookup[s2] = v;
                            break;
                        }
                        dirty = true;
                    }
            }
        }
        build_variable_lookup(check_lookup, n, m, dv, dc, variable_lookup);
        return 1;
    }
    return 0;
}

// Graph `index` of the in-kernel ensemble generator for `graph_seed` (host twin of e1_graphgen.cuh), with the
// (variable_lookup, check_lookup) contract of random_code_generator.c.  n*dv must be < 65536.
extern "C" int ldpc_b200_generate_graph_indexed(int n, int dv, int dc, uint64_t graph_seed, uint64_t index,
                                                int* variable_lookup, int* check_lookup) {
    if (n <= 0 || dv <= 0 || dc <= 0 || !variable_lookup || !check_lookup) return 0;
    const long long E = (long long)n * dv;
    if (E % dc != 0 || E > 65535 || dc > n) return 0;
    if (!generate_device_compatible(n, dv, dc, graph_seed, index, check_lookup)) return 0;
    build_variable_lookup(check_lookup, n, (int)(E / dc), dv, dc, variable_lookup);
    return 1;
}
