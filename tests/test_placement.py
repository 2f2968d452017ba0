"""CPU test of the shared-memory placement search (csrc/relabel.cpp): whatever the search does -- serial or on several
threads -- its output must be the SAME graph under a relabelling (slots are permutations, the placed lists hold exactly
the relabelled edges, var_orig inverts var_slot), and it must not be worse than the natural layout."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "iib_project_ldpc_codes_b200", "csrc")

SRC = r'''
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "relabel.hpp"
extern "C" int ldpc_b200_generate_graph(int n, int dv, int dc, uint64_t seed, int mode, int* vl, int* cl);
int main(int argc, char** argv) {
    const int n = atoi(argv[1]), dv = atoi(argv[2]), dc = atoi(argv[3]), B = atoi(argv[4]), effort = atoi(argv[5]);
    const int m = n * dv / dc;
    std::vector<int> vl((size_t)n * dv), cl((size_t)m * dc);
    if (ldpc_b200_generate_graph(n, dv, dc, 5, 1, vl.data(), cl.data()) != 1) { printf("FAIL generate\n"); return 1; }
    ldpc::Placement P;
    ldpc::place_graph(n, m, dv, dc, vl.data(), cl.data(), B, effort, 42, P);
    auto is_perm = [](std::vector<int> v) { std::sort(v.begin(), v.end()); for (size_t i = 0; i < v.size(); i++) if (v[i] != (int)i) return false; return true; };
    if (!is_perm(P.var_slot) || !is_perm(P.chk_slot) || !is_perm(P.var_orig)) { printf("FAIL permutation\n"); return 1; }
    for (int i = 0; i < n; i++) if (P.var_orig[P.var_slot[i]] != i) { printf("FAIL var_orig\n"); return 1; }
    // edges of the placed lists == relabelled edges of the input (as sorted multisets per node)
    for (int c = 0; c < m; c++) {
        std::vector<int> a, b;
        for (int j = 0; j < dc; j++) { a.push_back(P.var_slot[cl[(size_t)c * dc + j]]); b.push_back(P.c2v[(size_t)P.chk_slot[c] * dc + j]); }
        std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
        if (a != b) { printf("FAIL c2v check %d\n", c); return 1; }
    }
    for (int i = 0; i < n; i++) {
        std::vector<int> a, b;
        for (int j = 0; j < dv; j++) { a.push_back(P.chk_slot[vl[(size_t)i * dv + j]]); b.push_back(P.v2c[(size_t)P.var_slot[i] * dv + j]); }
        std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
        if (a != b) { printf("FAIL v2c variable %d\n", i); return 1; }
    }
    if (P.wavefronts_placed > P.wavefronts_natural || P.wavefronts_placed < P.wavefronts_ideal) { printf("FAIL wavefronts\n"); return 1; }
    printf("OK %.0f %.0f %.0f\n", P.wavefronts_natural, P.wavefronts_placed, P.wavefronts_ideal);
    return 0;
}
'''


@pytest.fixture(scope="module")
def exe():
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "t.cpp")
        out = os.path.join(tmp, "t")
        with open(src, "w") as f:
            f.write(SRC)
        subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", CSRC, "-o", out, src, os.path.join(CSRC, "relabel.cpp"),
                        os.path.join(CSRC, "graphgen.cpp")], check=True)
        yield out


@pytest.mark.parametrize("n,dv,dc,B,effort,threads", [(1000, 3, 6, 8, 2, 1), (1000, 3, 6, 8, 2, 4), (2000, 4, 8, 16, 1, 3),
                                                    (504, 3, 6, 32, 2, 2), (96, 2, 4, 8, 3, 1), (1003 * 2, 3, 6, 16, 1, 5),
                                                    (300, 5, 10, 8, 0, 1)])
def test_placement_is_a_relabelling(exe, n, dv, dc, B, effort, threads):
    env = dict(os.environ, LDPC_B200_PLACE_THREADS=str(threads))
    out = subprocess.run([exe, str(n), str(dv), str(dc), str(B), str(effort)], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
    nat, placed, ideal = (float(v) for v in out.stdout.split()[1:4])
    if effort > 0 and n >= 500:
        assert placed < 0.75 * nat          # the search removes most of the conflict replays
