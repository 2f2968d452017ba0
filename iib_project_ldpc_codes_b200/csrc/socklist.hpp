// socklist.hpp -- memory layout of the per-node socket lists (check -> variables, variable -> checks) of E1.
//
// A node with d sockets stores d uint16 neighbour slots.  The reference keeps them row-major per node
// (check_to_variable_list[c*dc + j], message_passing.c:24-36); a warp whose lanes are consecutive nodes
// wants every load conflict-free and as few load instructions as possible (on B200 one LDS costs >= 1.3
// cycles of the shared-memory pipe however narrow it is: tools/ubench/smem_rate.cu).  Layout, for `pad`
// nodes (pad a multiple of 8):
//
//     sockets 4g .. 4g+3  ("quad" g < d/4):  one 8-byte word per node   at u16 offset  (g*pad + node)*4
//     then, if d%4 >= 2, a "pair":           one 4-byte word per node   at            4*(d/4)*pad + node*2
//     then, if d is odd, a "single":         one uint16 per node        at            (d-1)*pad + node
//
// d = 6: LDS.64 + LDS.32 per check instead of six LDS.U16; d = 3: LDS.32 + LDS.U16; d = 4 / 8: LDS.64 (x2).
// The array has exactly d*pad uint16, like the socket-major layout it replaces.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LDPC_HD __host__ __device__
#else
#define LDPC_HD
#endif

namespace ldpc {

LDPC_HD inline size_t sock_off(int d, int pad, int j, int node) {
    const int nq = d >> 2;
    if (j < 4 * nq) return ((size_t)(j >> 2) * pad + node) * 4 + (j & 3);
    if ((d & 3) >= 2 && j < 4 * nq + 2) return (size_t)4 * nq * pad + (size_t)node * 2 + (j - 4 * nq);
    return (size_t)(d - 1) * pad + node;
}

}  // namespace ldpc
