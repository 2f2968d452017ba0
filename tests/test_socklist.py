"""CPU test of the packed socket-list layout (csrc/socklist.hpp): for every degree the offsets of (socket, node) are a
bijection onto [0, d*pad), quads are 8-byte aligned words, pairs 4-byte aligned, and the layout is the one the kernels'
vector loads assume (gather_addrs in e1_kernels.cuh)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include <vector>
#include "socklist.hpp"
int main() {
    for (int d = 1; d <= 13; d++)
        for (int pad = 8; pad <= 40; pad += 8) {
            std::vector<int> seen((size_t)d * pad, 0);
            for (int j = 0; j < d; j++)
                for (int node = 0; node < pad; node++) {
                    const size_t o = ldpc::sock_off(d, pad, j, node);
                    if (o >= seen.size() || seen[o]++) { printf("FAIL bijection d=%d pad=%d j=%d node=%d\n", d, pad, j, node); return 1; }
                    const int nq = d / 4;
                    if (j < 4 * nq) {            // quad g = j/4 of `node` is one 8-byte word
                        if (o != ((size_t)(j / 4) * pad + node) * 4 + (j % 4)) { printf("FAIL quad\n"); return 1; }
                    } else if ((d & 3) >= 2 && j < 4 * nq + 2) {   // pair: one 4-byte word per node after the quads
                        if (o != (size_t)4 * nq * pad + (size_t)node * 2 + (j - 4 * nq)) { printf("FAIL pair\n"); return 1; }
                    } else {                     // single: last socket of an odd degree
                        if (j != d - 1 || o != (size_t)(d - 1) * pad + node) { printf("FAIL single\n"); return 1; }
                    }
                }
        }
    printf("OK\n");
    return 0;
}
'''


def test_sock_off_is_a_bijection_with_aligned_vectors():
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "t.cpp")
        exe = os.path.join(tmp, "t")
        with open(src, "w") as f:
            f.write(SRC)
        subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "iib_project_ldpc_codes_b200", "csrc"),
                        "-o", exe, src], check=True)
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.strip() == "OK", out.stdout + out.stderr
