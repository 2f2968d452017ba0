// relabel.cpp -- see relabel.hpp.  Host-side preprocessing, runs once per (graph, words-per-group).
#include "relabel.hpp"

#include <math.h>
#include <string.h>
#include <algorithm>
#include <numeric>
#include <stdlib.h>
#include <thread>

namespace ldpc {

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t b) { return (uint32_t)(((next() >> 32) * (uint64_t)b) >> 32); }
    double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// One side of the bipartite graph: N nodes of degree d processed in rows of B slots; cnt[row][g] =
// sockets of the row whose neighbour sits in bank group g.
struct SideState {
    int N = 0, d = 0, B = 0, rows = 0;
    const int* nbr = nullptr;       // [N][d] neighbour ids (other side), original numbering
    std::vector<int> slot;          // node -> slot
    std::vector<int> inv;           // slot -> node
    std::vector<int> cnt;           // [rows][B]
    int row_of(int node) const { return slot[node] / B; }
    int grp_of(int node) const { return slot[node] % B; }
};

struct Search {
    SideState V, C;   // V: variables (neighbours = checks), C: checks (neighbours = variables)
    long long cost = 0;

    inline void bump(SideState& S, int row, int g, int delta) {
        int& c = S.cnt[(size_t)row * S.B + g];
        if (delta > 0) { if (c >= S.d) cost++; c++; }
        else           { if (c > S.d) cost--; c--; }
    }
    void init_counts() {
        std::fill(V.cnt.begin(), V.cnt.end(), 0);
        std::fill(C.cnt.begin(), C.cnt.end(), 0);
        cost = 0;
        // every edge once from the variable side: both tables
        for (int v = 0; v < V.N; v++)
            for (int j = 0; j < V.d; j++) {
                const int c = V.nbr[(size_t)v * V.d + j];
                bump(V, V.row_of(v), C.grp_of(c), +1);
                bump(C, C.row_of(c), V.grp_of(v), +1);
            }
    }
};

// Socket order of one row: `cnt_nodes` nodes, each with d sockets of bank group grp[node*d + s].
// order[node*d + j] = socket placed at position j.  Column by column: a matching that covers every
// bank group whose remaining degree is >= the remaining columns (else it would collide later), then
// extended to as many nodes as possible; leftovers take the least loaded group.
void arrange_row(int cnt_nodes, int d, int B, const int* grp, int* order) {
    std::vector<char> used((size_t)cnt_nodes * d, 0);
    std::vector<int> match_g(B), match_n(cnt_nodes), sock_of(cnt_nodes), deg(B), load(B);
    std::vector<char> vis;
    for (int j = 0; j < d; j++) {
        const int rem = d - j;
        std::fill(deg.begin(), deg.end(), 0);
        for (int u = 0; u < cnt_nodes; u++)
            for (int s = 0; s < d; s++)
                if (!used[(size_t)u * d + s]) deg[grp[(size_t)u * d + s]]++;
        std::fill(match_g.begin(), match_g.end(), -1);
        std::fill(match_n.begin(), match_n.end(), -1);
        auto has = [&](int u, int g) {
            for (int s = 0; s < d; s++)
                if (!used[(size_t)u * d + s] && grp[(size_t)u * d + s] == g) return s;
            return -1;
        };
        // augment from a group: g takes node u, the group that held u looks for another node
        struct Aug {
            static bool from_group(int g, int cnt_nodes, std::vector<char>& vis, std::vector<int>& match_g,
                                   std::vector<int>& match_n, const decltype(has)& has) {
                for (int u = 0; u < cnt_nodes; u++) {
                    if (vis[u] || has(u, g) < 0) continue;
                    vis[u] = 1;
                    if (match_n[u] < 0 || from_group(match_n[u], cnt_nodes, vis, match_g, match_n, has)) {
                        match_n[u] = g; match_g[g] = u;
                        return true;
                    }
                }
                return false;
            }
            static bool from_node(int u, int B, std::vector<char>& visg, std::vector<int>& match_g,
                                  std::vector<int>& match_n, const decltype(has)& has) {
                for (int g = 0; g < B; g++) {
                    if (visg[g] || has(u, g) < 0) continue;
                    visg[g] = 1;
                    if (match_g[g] < 0 || from_node(match_g[g], B, visg, match_g, match_n, has)) {
                        match_g[g] = u; match_n[u] = g;
                        return true;
                    }
                }
                return false;
            }
        };
        std::vector<int> gs(B);
        std::iota(gs.begin(), gs.end(), 0);
        std::sort(gs.begin(), gs.end(), [&](int a, int b) { return deg[a] > deg[b]; });
        for (int g : gs) {
            if (deg[g] < rem) break;
            vis.assign(cnt_nodes, 0);
            Aug::from_group(g, cnt_nodes, vis, match_g, match_n, has);
        }
        for (int u = 0; u < cnt_nodes; u++) {
            if (match_n[u] >= 0) continue;
            vis.assign(B, 0);
            Aug::from_node(u, B, vis, match_g, match_n, has);
        }
        std::fill(load.begin(), load.end(), 0);
        for (int u = 0; u < cnt_nodes; u++)
            if (match_n[u] >= 0) load[match_n[u]]++;
        for (int u = 0; u < cnt_nodes; u++) {
            int s;
            if (match_n[u] >= 0) s = has(u, match_n[u]);
            else {
                s = -1;
                for (int t = 0; t < d; t++)
                    if (!used[(size_t)u * d + t] && (s < 0 || load[grp[(size_t)u * d + t]] < load[grp[(size_t)u * d + s]])) s = t;
                load[grp[(size_t)u * d + s]]++;
            }
            used[(size_t)u * d + s] = 1;
            order[(size_t)u * d + j] = s;
        }
    }
}

// Parallel version of the annealing loop in place_graph.  The run is cut into phases; a phase changes the slots of
// ONE side only, and every thread owns a pseudo-random subset of that side's rows (a different partition each
// phase, so nodes migrate between threads over time) and swaps only nodes inside rows it owns:
//   * a row swap (same bank group, two owned rows) reads and writes cells of the own side's table in owned rows only;
//   * a bank-group swap inside an owned row changes cells of the OTHER side's table, which other threads change
//     too: relaxed atomic adds keep the counts exact, only the evaluation of a move can see a count that is being
//     changed -- a little noise the annealing does not mind.
// The other side's slots are constant during a phase.  The cost is recounted exactly afterwards by the caller.
// Which thread wins a race is not reproducible, so neither is the placement; results never depend on it.
void anneal_parallel(Search& S, int B, long long moves_per_thread, uint64_t seed, int T) {
    const int phases = 96;
    const double t0 = 0.25, t1 = 0.03;
    const long long per_phase = std::max<long long>(1, moves_per_thread / phases);
    for (int ph = 0; ph < phases; ph++) {
        SideState& A = (ph & 1) ? S.C : S.V;
        SideState& O = (ph & 1) ? S.V : S.C;
        if (A.d > 64) continue;
        const int full_rows = A.N / B;   // the partial last row stays where it is
        if (full_rows < 2) continue;
        std::vector<std::vector<int>> own(T);
        {
            Rng pr(seed * 7919u + (uint64_t)ph);
            for (int r = 0; r < full_rows; r++) own[pr.below((uint32_t)T)].push_back(r);
        }
        const double Tlo = t0 * pow(t1 / t0, (double)ph / phases), Thi = t0 * pow(t1 / t0, (double)(ph + 1) / phases);
        auto work = [&](int t) {
            const std::vector<int>& rows = own[t];
            if (rows.size() < 2) return;
            Rng rng(seed + 0x9E37u * (uint64_t)(ph * 64 + t + 1));
            for (long long it = 0; it < per_phase; it++) {
                const double temp = Tlo + (Thi - Tlo) * (double)it / (double)per_phase;
                const int r1 = rows[rng.below((uint32_t)rows.size())];
                const int u1 = A.inv[r1 * B + (int)rng.below((uint32_t)B)];
                const bool group_swap = rng.below(2u) != 0;
                int s2;
                if (group_swap) s2 = r1 * B + (int)rng.below((uint32_t)B);
                else s2 = rows[rng.below((uint32_t)rows.size())] * B + A.grp_of(u1);
                const int u2 = A.inv[s2];
                if (u1 == u2) continue;
                SideState& Tb = group_swap ? O : A;
                int* cells[4 * 64];
                int dl[4 * 64];
                int nc = 0;
                auto add = [&](int row, int g, int k) {
                    int* c = &Tb.cnt[(size_t)row * B + g];
                    for (int q = 0; q < nc; q++) if (cells[q] == c) { dl[q] += k; return; }
                    cells[nc] = c; dl[nc] = k; nc++;
                };
                if (group_swap) {
                    const int g1 = A.grp_of(u1), g2 = A.grp_of(u2);
                    for (int j = 0; j < A.d; j++) {
                        const int ra = O.row_of(A.nbr[(size_t)u1 * A.d + j]), rb = O.row_of(A.nbr[(size_t)u2 * A.d + j]);
                        add(ra, g1, -1); add(ra, g2, +1);
                        add(rb, g2, -1); add(rb, g1, +1);
                    }
                } else {
                    const int ra = A.row_of(u1), rb = A.row_of(u2);
                    for (int j = 0; j < A.d; j++) {
                        const int ga = O.grp_of(A.nbr[(size_t)u1 * A.d + j]), gb = O.grp_of(A.nbr[(size_t)u2 * A.d + j]);
                        add(ra, ga, -1); add(rb, ga, +1);
                        add(rb, gb, -1); add(ra, gb, +1);
                    }
                }
                long long delta = 0;
                for (int q = 0; q < nc; q++) {
                    const int c = __atomic_load_n(cells[q], __ATOMIC_RELAXED);
                    delta += std::max(0, c + dl[q] - Tb.d) - std::max(0, c - Tb.d);
                }
                if (delta > 0 && rng.unit() >= exp(-(double)delta / temp)) continue;
                for (int q = 0; q < nc; q++) if (dl[q]) __atomic_fetch_add(cells[q], dl[q], __ATOMIC_RELAXED);
                std::swap(A.slot[u1], A.slot[u2]);
                A.inv[A.slot[u1]] = u1; A.inv[A.slot[u2]] = u2;
            }
        };
        std::vector<std::thread> th;
        std::vector<int> inline_work{0};
        for (int t = 1; t < T; t++) {
            try { th.emplace_back(work, t); }
            catch (...) { inline_work.push_back(t); }   // no more threads to be had: this share runs on the caller's thread
        }
        for (int t : inline_work) work(t);
        for (auto& x : th) x.join();
    }
}

double side_wavefronts(int N, int d, int B, const int* nbr_slots) {
    double total = 0;
    std::vector<int> load(B);
    for (int base = 0; base < N; base += B) {
        const int cntn = std::min(B, N - base);
        for (int j = 0; j < d; j++) {
            std::fill(load.begin(), load.end(), 0);
            int mx = 0;
            for (int u = 0; u < cntn; u++) mx = std::max(mx, ++load[nbr_slots[(size_t)(base + u) * d + j] % B]);
            total += mx;
        }
    }
    return total;
}

}  // namespace

double gather_wavefronts(int n, int m, int dv, int dc, const int* v2c, const int* c2v, int B) {
    return side_wavefronts(m, dc, B, c2v) + side_wavefronts(n, dv, B, v2c);
}

void place_graph(int n, int m, int dv, int dc, const int* v2c, const int* c2v, int B, int effort, uint64_t seed,
                 Placement& out, double budget_scale) {
    out.n = n; out.m = m; out.dv = dv; out.dc = dc; out.B = B;
    out.wavefronts_natural = gather_wavefronts(n, m, dv, dc, v2c, c2v, B);
    out.wavefronts_ideal = 0;
    for (int base = 0; base < m; base += B) out.wavefronts_ideal += dc;
    for (int base = 0; base < n; base += B) out.wavefronts_ideal += dv;

    Search S;
    S.V.N = n; S.V.d = dv; S.V.B = B; S.V.rows = (n + B - 1) / B; S.V.nbr = v2c;
    S.C.N = m; S.C.d = dc; S.C.B = B; S.C.rows = (m + B - 1) / B; S.C.nbr = c2v;
    S.V.slot.resize(n); S.V.inv.resize(n); S.C.slot.resize(m); S.C.inv.resize(m);
    std::iota(S.V.slot.begin(), S.V.slot.end(), 0); std::iota(S.V.inv.begin(), S.V.inv.end(), 0);
    std::iota(S.C.slot.begin(), S.C.slot.end(), 0); std::iota(S.C.inv.begin(), S.C.inv.end(), 0);
    S.V.cnt.assign((size_t)S.V.rows * B, 0);
    S.C.cnt.assign((size_t)S.C.rows * B, 0);
    S.init_counts();

    // threads: as many as give every thread >= 12 full rows of the smaller side, at most 16 (LDPC_B200_PLACE_THREADS overrides)
    int threads = 1;
    {
        const char* e = getenv("LDPC_B200_PLACE_THREADS");
        const int hw = (int)std::thread::hardware_concurrency();
        threads = e ? atoi(e) : std::min(std::min(hw > 0 ? hw : 1, 16), (m / B) / 12);
        if (threads < 1) threads = 1;
    }
    if (effort > 0 && n > 1 && m > 1 && threads > 1) {
        // `effort` is a time budget: every thread makes as many moves as the single-threaded search would
        anneal_parallel(S, B, (long long)((double)effort * 200.0 * ((double)n + m) * budget_scale), seed, threads);
        S.init_counts();   // exact tables and cost again
    } else if (effort > 0 && n > 1 && m > 1) {
        Rng rng(seed);
        const long long moves = (long long)((double)effort * 200.0 * ((double)n + m) * budget_scale);
        const double t0 = 0.25, t1 = 0.03;
        double T = t0;
        for (long long it = 0; it < moves && S.cost > 0; it++) {
            if ((it & 1023) == 0) T = t0 * pow(t1 / t0, (double)it / (double)moves);
            const bool var_side = rng.below((uint32_t)(n + m)) < (uint32_t)n;
            SideState& A = var_side ? S.V : S.C;
            SideState& O = var_side ? S.C : S.V;
            // Structured moves: a swap inside a row only permutes bank groups (it changes the histograms of the OTHER
            // side's rows and nothing else); a swap of two nodes of the same bank group between two rows only changes
            // this side's two rows.  Each touches one of the two cost terms, so far more of them are accepted at low
            // temperature than swaps of arbitrary pairs: n = 10^4, B = 16, same run time: 1.35x -> 1.21x of the
            // conflict-free wavefront count (n = 1000, B = 8: 1.14x -> 1.06x).
            const int u1 = (int)rng.below((uint32_t)A.N);
            const bool group_swap = rng.below(2u) != 0;
            int s2;
            if (group_swap) s2 = A.row_of(u1) * B + (int)rng.below((uint32_t)B);
            else s2 = (int)rng.below((uint32_t)A.rows) * B + A.grp_of(u1);
            if (s2 >= A.N) continue;
            const int u2 = A.inv[s2];
            if (u1 == u2) continue;
            // Net change per touched cell first (a cell can be hit several times), the cost change from that, and the
            // tables are only written when the move is accepted -- most moves are rejected at low temperature.
            SideState& Tb = group_swap ? O : A;   // the table the move changes
            int* cells[4 * 64];
            int dl[4 * 64];
            int nc = 0;
            auto add = [&](int row, int g, int k) {
                int* c = &Tb.cnt[(size_t)row * B + g];
                for (int q = 0; q < nc; q++) if (cells[q] == c) { dl[q] += k; return; }
                cells[nc] = c; dl[nc] = k; nc++;
            };
            if (A.d > 64) continue;
            if (group_swap) {
                const int g1 = A.grp_of(u1), g2 = A.grp_of(u2);
                for (int j = 0; j < A.d; j++) {
                    const int r1 = O.row_of(A.nbr[(size_t)u1 * A.d + j]), r2 = O.row_of(A.nbr[(size_t)u2 * A.d + j]);
                    add(r1, g1, -1); add(r1, g2, +1);
                    add(r2, g2, -1); add(r2, g1, +1);
                }
            } else {
                const int r1 = A.row_of(u1), r2 = A.row_of(u2);
                for (int j = 0; j < A.d; j++) {
                    const int ga = O.grp_of(A.nbr[(size_t)u1 * A.d + j]), gb = O.grp_of(A.nbr[(size_t)u2 * A.d + j]);
                    add(r1, ga, -1); add(r2, ga, +1);
                    add(r2, gb, -1); add(r1, gb, +1);
                }
            }
            long long delta = 0;
            for (int q = 0; q < nc; q++) {
                const int c = *cells[q];
                delta += std::max(0, c + dl[q] - Tb.d) - std::max(0, c - Tb.d);
            }
            if (delta > 0 && rng.unit() >= exp(-(double)delta / T)) continue;   // reject: nothing was touched
            for (int q = 0; q < nc; q++) *cells[q] += dl[q];
            S.cost += delta;
            std::swap(A.slot[u1], A.slot[u2]);
            A.inv[A.slot[u1]] = u1; A.inv[A.slot[u2]] = u2;
        }
    }

    // ---- socket order, row by row ----
    out.var_slot = S.V.slot; out.chk_slot = S.C.slot;
    out.var_orig = S.V.inv;
    out.c2v.assign((size_t)m * dc, 0);
    out.v2c.assign((size_t)n * dv, 0);
    auto emit = [&](SideState& A, SideState& O, std::vector<int>& dst) {
        const int d = A.d;
        std::vector<int> grp((size_t)B * d), order((size_t)B * d);
        for (int base = 0; base < A.N; base += B) {
            const int cntn = std::min(B, A.N - base);
            for (int u = 0; u < cntn; u++) {
                const int node = A.inv[base + u];
                for (int s = 0; s < d; s++) grp[(size_t)u * d + s] = O.slot[A.nbr[(size_t)node * d + s]] % B;
            }
            if (effort > 0) arrange_row(cntn, d, B, grp.data(), order.data());
            else for (int u = 0; u < cntn; u++) for (int s = 0; s < d; s++) order[(size_t)u * d + s] = s;
            for (int u = 0; u < cntn; u++) {
                const int node = A.inv[base + u];
                for (int j = 0; j < d; j++)
                    dst[(size_t)(base + u) * d + j] = O.slot[A.nbr[(size_t)node * d + order[(size_t)u * d + j]]];
            }
        }
    };
    emit(S.C, S.V, out.c2v);
    emit(S.V, S.C, out.v2c);
    out.wavefronts_placed = gather_wavefronts(n, m, dv, dc, out.v2c.data(), out.c2v.data(), B);
}

}  // namespace ldpc

#ifdef RELABEL_MAIN
// g++ -O2 -DRELABEL_MAIN relabel.cpp graphgen.cpp -o /tmp/relabel && /tmp/relabel n dv dc B effort
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
extern "C" int ldpc_b200_generate_graph(int n, int dv, int dc, uint64_t seed, int mode, int* vl, int* cl);
int main(int argc, char** argv) {
    int n = atoi(argv[1]), dv = atoi(argv[2]), dc = atoi(argv[3]), B = atoi(argv[4]), effort = atoi(argv[5]);
    int m = n * dv / dc;
    std::vector<int> vl((size_t)n * dv), cl((size_t)n * dv);
    ldpc_b200_generate_graph(n, dv, dc, 1, 1, vl.data(), cl.data());
    ldpc::Placement P;
    auto t0 = std::chrono::steady_clock::now();
    ldpc::place_graph(n, m, dv, dc, vl.data(), cl.data(), B, effort, 1, P);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("n=%d B=%d effort=%d: natural %.0f placed %.0f ideal %.0f  (%.3fx -> %.3fx of ideal) %.2fs\n", n, B, effort,
           P.wavefronts_natural, P.wavefronts_placed, P.wavefronts_ideal, P.wavefronts_natural / P.wavefronts_ideal,
           P.wavefronts_placed / P.wavefronts_ideal, sec);
    return 0;
}
#endif
