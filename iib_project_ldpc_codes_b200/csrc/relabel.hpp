// relabel.hpp -- bank-conflict-aware placement of a Tanner graph in shared memory (host side).
//
// In the E1 kernels lane l of a warp handles node slot (base + l) and, for socket position j, gathers
// the state word of the neighbour stored at slot s: shared-memory bank group = s mod B, where
// B = 32 / WORDS (a 32*WORDS-bit word per node: LDS.32 -> 32 lanes x 32 banks, LDS.64 -> half-warps of
// 16 lanes x 16 bank pairs, LDS.128 -> quarter-warps of 8 lanes x 8 bank quads).  With the reference's
// random socket permutation (random_code_generator.c:13-36) the B gathers of a phase fall on random
// bank groups: E[max load] = 3.5 (B=32) / 2.4 (B=8) wavefronts instead of 1  (profiles/r1a_*).
//
// Node numbering and socket order are free as far as the decoder's result is concerned (OR/AND are
// commutative; the erased count is a sum over variables), so we choose
//   * a slot for every variable and every check (a permutation), and
//   * the order of the sockets inside every check and every variable,
// such that in every aligned row of B consecutive slots, for every socket position, the B neighbours
// lie in B distinct bank groups.  A row admits a conflict-free socket order iff every bank group
// occurs at most d times among its B*d sockets (Koenig: a bipartite multigraph of maximum degree d
// splits into d matchings), so the search minimises the total excess  sum_rows sum_g (count_g - d)^+
// over both sides by greedy row construction + simulated annealing on slot swaps, then orders the
// sockets row by row with bipartite matchings.
//
// The erasure pattern stays keyed by the ORIGINAL variable id (var_orig), so results are bit-identical
// to the natural layout.
#pragma once
#include <stdint.h>
#include <vector>

namespace ldpc {

struct Placement {
    int n = 0, m = 0, dv = 0, dc = 0, B = 0;
    std::vector<int> var_slot;    // [n]  variable -> slot
    std::vector<int> chk_slot;    // [m]  check -> slot
    std::vector<int> c2v;         // [m][dc] in slot space: c2v[cs*dc + j] = variable slot at socket j of check slot cs
    std::vector<int> v2c;         // [n][dv] in slot space
    std::vector<int> var_orig;    // [n]  slot -> original variable id
    double wavefronts_natural = 0, wavefronts_placed = 0, wavefronts_ideal = 0;  // gather wavefronts per iteration
};

// gather wavefronts of one flooding iteration for a layout given in slot space
double gather_wavefronts(int n, int m, int dv, int dc, const int* v2c, const int* c2v, int B);

// effort: 0 = natural layout (identity), >= 1 = optimise (annealing moves ~ effort * 200 * (n + m))
// budget_scale multiplies the move budget (the quick first pass of the background search uses 0.25)
void place_graph(int n, int m, int dv, int dc, const int* v2c, const int* c2v, int B, int effort, uint64_t seed,
                 Placement& out, double budget_scale = 1.0);

}  // namespace ldpc
