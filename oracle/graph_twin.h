/*
 * oracle/graph_twin.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Independent C twin of the seeded simple-regular-graph generator the benchmarks and tests use
 * (ldpc_b200_generate_graph, mode 1): SplitMix64-driven Fisher-Yates socket permutation
 * (the construction of /root/reference/random_code_generator.c:13-36 with a seeded generator in
 * place of libc rand()), repeated variables inside a check repaired by socket swaps, then
 * check_lookup[c*dc+j] = socket/dv (:34-36) and variable_lookup in ascending check order (:53-64).
 * It exists so that the CPU reference arm of bench.py can build the benchmark's graph without
 * loading the product library; tests/test_oracle.py checks it against the product generator.
 */
#ifndef ORACLE_GRAPH_TWIN_H
#define ORACLE_GRAPH_TWIN_H
#include <stdint.h>
#include <stdlib.h>

static uint64_t orc_sm64_state;
static uint64_t orc_sm64_next(void) {
    uint64_t z = (orc_sm64_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint32_t orc_sm64_below(uint32_t bound) { /* unbiased integer in [0, bound): multiply-shift with rejection */
    uint64_t x = orc_sm64_next() >> 32, mres = x * bound;
    uint32_t lo = (uint32_t)mres;
    if (lo < bound) {
        uint32_t thresh = (0u - bound) % bound;
        while (lo < thresh) { x = orc_sm64_next() >> 32; mres = x * bound; lo = (uint32_t)mres; }
    }
    return (uint32_t)(mres >> 32);
}

static int orc_check_holds(const int *cl, int dc, int c, int v, int skip) {
    for (int j = 0; j < dc; j++) if (j != skip && cl[(size_t)c * dc + j] == v) return 1;
    return 0;
}

/* returns 1 on success; not re-entrant (test infrastructure) */
static int orc_generate_graph_twin(int n, int dv, int dc, uint64_t seed, int *variable_lookup, int *check_lookup) {
    long long E = (long long)n * dv;
    if (n <= 0 || dv <= 0 || dc <= 0 || E % dc || dc > n) return 0;
    int m = (int)(E / dc);
    orc_sm64_state = seed * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull;
    int *seq = (int *)malloc(sizeof(int) * (size_t)E);
    for (long long e = 0; e < E; e++) seq[e] = (int)e;
    for (long long i = E - 1; i > 0; i--) {
        uint32_t j = orc_sm64_below((uint32_t)(i + 1));
        int tmp = seq[i]; seq[i] = seq[j]; seq[j] = tmp;
    }
    for (long long e = 0; e < E; e++) check_lookup[e] = seq[e] / dv;
    free(seq);
    long long guard = 0;
    for (int dirty = 1; dirty;) {
        dirty = 0;
        for (int c = 0; c < m; c++)
            for (int a = 1; a < dc; a++) {
                int v = check_lookup[(size_t)c * dc + a], rep = 0;
                for (int b = 0; b < a; b++) rep |= (check_lookup[(size_t)c * dc + b] == v);
                if (!rep) continue;
                for (;;) {
                    if (++guard > 64 * E + 1000000) return 0;
                    uint32_t s2 = orc_sm64_below((uint32_t)E);
                    int c2 = (int)(s2 / dc), j2 = (int)(s2 % dc);
                    if (c2 == c) continue;
                    int v2 = check_lookup[s2];
                    if (orc_check_holds(check_lookup, dc, c, v2, a) || orc_check_holds(check_lookup, dc, c2, v, j2)) continue;
                    check_lookup[(size_t)c * dc + a] = v2;
                    check_lookup[s2] = v;
                    break;
                }
                dirty = 1;
            }
    }
    int *fill = (int *)calloc((size_t)n, sizeof(int));
    for (int c = 0; c < m; c++)
        for (int j = 0; j < dc; j++) {
            int v = check_lookup[(size_t)c * dc + j];
            if (fill[v] < dv) variable_lookup[(size_t)v * dv + fill[v]] = c;
            fill[v]++;
        }
    free(fill);
    return 1;
}
#endif
