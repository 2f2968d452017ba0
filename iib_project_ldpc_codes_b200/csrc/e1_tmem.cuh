// e1_tmem.cuh -- the Tanner-graph socket lists of the E1 Monte-Carlo kernels in TENSOR MEMORY (sm_100a).
//
// message_passing.c:24-36 / :52-60 walk check_to_variable_list / variable_to_check_list once per iteration.  In the
// E1 kernels thread `tid` of a team always handles the same nodes (tid, tid + T, tid + 2T, ...), so its slice of
// the lists is private to it and never changes during a launch.  Blackwell's tensor memory (256 KB per SM, 128
// lanes x 512 columns x 32 bit, reached with tcgen05.ld / tcgen05.st; a warp owns the 32 lanes 32*(warp%4)..+31)
// is exactly such a per-thread private store.  Every thread keeps, per node it handles, the READY-MADE 32-bit
// shared-memory addresses of the state vectors that node gathers (base + slot * vector size):
//
//   * no index loads on the shared-memory pipe (they were 3 of 17 wavefronts per check and 2 of 12 per variable
//     at W = 64) -- tcgen05.ld runs on its own datapath, next to a saturated LDS pipe
//     (tools/ubench/tmem_probe.cu on the box: 6 LDS.64 gathers per trip through TMEM-supplied addresses run at
//     127.9 B/clk/SM = the shared-memory peak, against 73 B/clk/SM when the addresses are computed in the loop);
//   * no unpack arithmetic (12 ALU instructions per check, 6 per variable);
//   * the 60-120 KB of shared memory the lists occupied are free for state (n = 8000 fits W = 128, n = 10^4 keeps
//     a second pattern buffer so that the channel draw overlaps the decoding).
//
// Column map of a thread (warp w of the CTA, lane l -> TMEM lane 32*(w%4)+l): the warp's columns start at
// (w/4) * cols_per_warp; check trip t uses DC columns from t*DC, variable trip t uses DV columns from
// trips_c*DC + t*DV.  Trip t of a team thread handles node t*T + tid; nodes past the end hold a harmless address.
#pragma once
#include <stdint.h>

namespace ldpc {

// ---- tcgen05.ld / tcgen05.st, shape 32x32b: thread l of the warp <-> TMEM lane (quarter base + l), D consecutive columns
__device__ __forceinline__ void tm_ld_x1(uint32_t ta, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(ta));
}
__device__ __forceinline__ void tm_ld_x2(uint32_t ta, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(ta));
}
__device__ __forceinline__ void tm_ld_x4(uint32_t ta, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ta));
}
__device__ __forceinline__ void tm_ld_x8(uint32_t ta, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta));
}
__device__ __forceinline__ void tm_st_x1(uint32_t ta, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(ta), "r"(r[0]) : "memory");
}
__device__ __forceinline__ void tm_st_x2(uint32_t ta, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(ta), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void tm_st_x4(uint32_t ta, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tm_st_x8(uint32_t ta, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// D columns as the fewest power-of-two pieces (warp-collective: every lane of the warp must call it)
template <int D> __device__ __forceinline__ void tm_ld(uint32_t ta, uint32_t (&r)[D]) {
    int o = 0;
    if constexpr (D >= 8) { tm_ld_x8(ta, r); o = 8; }
    if constexpr ((D & 7) >= 4) { tm_ld_x4(ta + o, r + o); o += 4; }
    if constexpr ((D & 3) >= 2) { tm_ld_x2(ta + o, r + o); o += 2; }
    if constexpr (D & 1) { tm_ld_x1(ta + o, r + o); }
    static_assert(D >= 1 && D <= 15, "tm_ld: 1..15 columns");
}
template <int D> __device__ __forceinline__ void tm_st(uint32_t ta, const uint32_t (&r)[D]) {
    int o = 0;
    if constexpr (D >= 8) { tm_st_x8(ta, r); o = 8; }
    if constexpr ((D & 7) >= 4) { tm_st_x4(ta + o, r + o); o += 4; }
    if constexpr ((D & 3) >= 2) { tm_st_x2(ta + o, r + o); o += 2; }
    if constexpr (D & 1) { tm_st_x1(ta + o, r + o); }
}
// tcgen05.wait::ld with the loaded registers as in/out operands: every use of them is ordered after the wait
template <int D> __device__ __forceinline__ void tm_wait_ld(uint32_t (&r)[D]) {
    if constexpr (D == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]));
    else if constexpr (D == 4) asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]));
    else if constexpr (D == 6) asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]));
    else if constexpr (D == 8)
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
    else {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int k = 0; k < D; k++) asm volatile("" : "+r"(r[k]));
    }
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// whole-CTA: warp 0 allocates all 512 columns (one CTA per SM), everybody learns the base address
__device__ __forceinline__ uint32_t tm_alloc_all(uint32_t* slot) {
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}
__device__ __forceinline__ void tm_free_all(uint32_t base) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

}  // namespace ldpc
