// Micro-benchmark: tensor memory (TMEM) as a per-thread private store for the Tanner-graph socket lists (sm_100a).
//
// Question it answers for the E1 decoder design: can every thread keep the (pre-scaled) shared-memory addresses of its
// nodes' neighbours in TMEM and fetch them with tcgen05.ld instead of LDS + unpack, (a) correctly for 24-28 warps per CTA
// (warp w reaches TMEM lanes 32*(w%4) .. +31 only), (b) at what latency, (c) at what throughput, alone and next to a
// saturated shared-memory pipe?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st4(uint32_t ta, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(ta), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tm_st2(uint32_t ta, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(ta), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tm_st1(uint32_t ta, uint32_t a) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(ta), "r"(a) : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t ta, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(ta) : "memory");
}
__device__ __forceinline__ void tm_ld2(uint32_t ta, uint32_t& a, uint32_t& b) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(ta) : "memory");
}
__device__ __forceinline__ void tm_ld1(uint32_t ta, uint32_t& a) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(a) : "r"(ta) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t tmem_setup(uint32_t* slot) {
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}
__device__ __forceinline__ void tmem_free(uint32_t base) {
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

// (a) correctness: every warp fills its column range in its lane quarter with f(tid, col) and reads it back
__global__ void __launch_bounds__(1024, 1) k_check(int cols_per_warp, unsigned long long* bad, uint32_t* base_out) {
    __shared__ uint32_t slot;
    const uint32_t base = tmem_setup(&slot);
    const int w = threadIdx.x >> 5;
    const uint32_t ta0 = base + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)((w >> 2) * cols_per_warp);
    auto f = [&](int c) { return (uint32_t)threadIdx.x * 2654435761u + (uint32_t)c * 40503u + 7u; };
    int c = 0;
    for (; c + 6 <= cols_per_warp; c += 6) {
        tm_st4(ta0 + c, f(c), f(c + 1), f(c + 2), f(c + 3));
        tm_st2(ta0 + c + 4, f(c + 4), f(c + 5));
    }
    for (; c < cols_per_warp; c++) tm_st1(ta0 + c, f(c));
    tm_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned long long nb = 0;
    c = 0;
    for (; c + 6 <= cols_per_warp; c += 6) {
        uint32_t r[6];
        tm_ld4(ta0 + c, r[0], r[1], r[2], r[3]);
        tm_ld2(ta0 + c + 4, r[4], r[5]);
        tm_wait_ld();
        for (int k = 0; k < 6; k++) nb += (r[k] != f(c + k));
    }
    for (; c < cols_per_warp; c++) { uint32_t r; tm_ld1(ta0 + c, r); tm_wait_ld(); nb += (r != f(c)); }
    if (nb) atomicAdd(bad, nb);
    if (threadIdx.x == 0 && blockIdx.x == 0) *base_out = base;
    tmem_free(base);
}

// (b)+(c): per trip one x4 + one x2 TMEM load (six addresses) and, optionally, six LDS.64 gathers through those addresses
// MODE 0: TMEM loads only; 1: LDS.64 only (addresses computed); 2: both (the decoder's check loop shape);
// MODE 3: both, next trip's addresses prefetched before this trip's gathers are consumed
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_rate(int trips, int iters, uint32_t* out, long long* clk) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ uint32_t slot;
    const uint32_t base = tmem_setup(&slot);
    uint32_t* s32 = reinterpret_cast<uint32_t*>(sm);
    for (int i = threadIdx.x; i < 20480; i += blockDim.x) s32[i] = i * 2654435761u;   // 80 KB of "state"
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ta0 = base + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)((w >> 2) * trips * 6);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    // conflict-free gather addresses: lane l of a half warp -> bank pair (l + rotation)
    for (int t = 0; t < trips; t++) {
        uint32_t a[6];
        for (int k = 0; k < 6; k++) {
            const uint32_t row = (uint32_t)((t * 6 + k) * 37 + w * 11) % 640u;          // 640 rows of 16 vectors (8 B each)
            a[k] = sbase + (row * 16u + (uint32_t)((lane + k + t) & 15)) * 8u;
        }
        tm_st4(ta0 + t * 6, a[0], a[1], a[2], a[3]);
        tm_st2(ta0 + t * 6 + 4, a[4], a[5]);
    }
    tm_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t acc0 = 0, acc1 = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 3) {
            uint32_t a[6], nx[6];
            tm_ld4(ta0, a[0], a[1], a[2], a[3]); tm_ld2(ta0 + 4, a[4], a[5]);
            for (int t = 0; t < trips; t++) {
                tm_wait_ld();
                uint32_t v[6][2];
#pragma unroll
                for (int k = 0; k < 6; k++) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v[k][0]), "=r"(v[k][1]) : "r"(a[k] + (uint32_t)(it & 1) * 0u));
                const int tn = t + 1 < trips ? t + 1 : 0;
                tm_ld4(ta0 + tn * 6, nx[0], nx[1], nx[2], nx[3]); tm_ld2(ta0 + tn * 6 + 4, nx[4], nx[5]);
#pragma unroll
                for (int k = 0; k < 6; k++) { acc0 ^= v[k][0]; acc1 |= v[k][1]; }
                tm_wait_ld();
#pragma unroll
                for (int k = 0; k < 6; k++) a[k] = nx[k];
            }
        } else {
            for (int t = 0; t < trips; t++) {
                uint32_t a[6];
                if (MODE != 1) {
                    tm_ld4(ta0 + t * 6, a[0], a[1], a[2], a[3]);
                    tm_ld2(ta0 + t * 6 + 4, a[4], a[5]);
                    tm_wait_ld();
                } else {
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        const uint32_t row = (uint32_t)((t * 6 + k) * 37 + w * 11) % 640u;
                        a[k] = sbase + (row * 16u + (uint32_t)((lane + k + t) & 15)) * 8u;
                    }
                }
                if (MODE == 0) {
#pragma unroll
                    for (int k = 0; k < 6; k++) acc0 ^= a[k];
                } else {
                    uint32_t v[6][2];
#pragma unroll
                    for (int k = 0; k < 6; k++) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v[k][0]), "=r"(v[k][1]) : "r"(a[k]));
#pragma unroll
                    for (int k = 0; k < 6; k++) { acc0 ^= v[k][0]; acc1 |= v[k][1]; }
                }
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 ^ acc1;
    tmem_free(base);
}

// latency of one dependent tcgen05.ld (+wait) chain: address of the next load comes from the previous one
__global__ void k_lat(int iters, uint32_t* out, long long* clk) {
    __shared__ uint32_t slot;
    const uint32_t base = tmem_setup(&slot);
    tm_st1(base, 0u);
    tm_wait_st();
    uint32_t off = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) { uint32_t r; tm_ld1(base + off, r); tm_wait_ld(); off = r; }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { clk[0] = t1 - t0; out[0] = off; }
    tmem_free(base);
}

template <int MODE> void rate(const char* name, int threads, int trips) {
    uint32_t* out; long long* clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    const int iters = 2000;
    cudaFuncSetAttribute(k_rate<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 81920);
    k_rate<MODE><<<148, threads, 81920>>>(trips, 10, out, clk);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_rate<MODE><<<148, threads, 81920>>>(trips, iters, out, clk);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    const double warp_trips = (double)iters * trips * (threads / 32);
    printf("%-34s threads %4d trips %2d: %.2f clk per warp-trip/SM, TMEM %.1f B/clk/SM, smem %.1f B/clk/SM (%.3f ms, %s)\n", name, threads, trips,
           c / warp_trips, MODE == 1 ? 0.0 : warp_trips * 32 * 24 / c, MODE == 0 ? 0.0 : warp_trips * 32 * 48 / c, ms,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(clk);
}

int main() {
    unsigned long long* bad; uint32_t* b0;
    cudaMalloc(&bad, 8); cudaMalloc(&b0, 4);
    for (int threads : {128, 768, 864, 1024}) {
        const int wpq = (threads / 32 + 3) / 4, cpw = 512 / wpq;
        cudaMemset(bad, 0, 8);
        k_check<<<148, threads>>>(cpw, bad, b0);
        unsigned long long h = 0; uint32_t hb = 0;
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&hb, b0, 4, cudaMemcpyDeviceToHost);
        printf("check: threads %4d, %3d columns per warp: %llu mismatches, base 0x%08x (%s)\n", threads, cpw, h, hb, cudaGetErrorString(e));
    }
    {
        uint32_t* out; long long* clk; cudaMalloc(&out, 4); cudaMalloc(&clk, 8);
        k_lat<<<1, 32>>>(10000, out, clk);
        long long c; cudaDeviceSynchronize(); cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
        printf("latency: tcgen05.ld.x1 + wait::ld dependent chain: %.1f clk (%s)\n", c / 10000.0, cudaGetErrorString(cudaGetLastError()));
    }
    for (int threads : {768, 1024}) {
        const int trips = 512 / ((threads / 32 + 3) / 4) / 6;
        rate<0>("TMEM x4+x2 only", threads, trips);
        rate<1>("LDS.64 x6 only (computed addrs)", threads, trips);
        rate<2>("TMEM addrs -> LDS.64 x6", threads, trips);
        rate<3>("same, next trip prefetched", threads, trips);
    }
    return 0;
}
