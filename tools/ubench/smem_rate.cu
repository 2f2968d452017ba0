// Micro-benchmark: shared-memory load issue rate by access width (sm_100a).  Prints per-SM bytes/clk and LDS/clk.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_rate smem_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int BYTES> struct Ld;
template <> struct Ld<2> { static __device__ __forceinline__ uint32_t ld(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory"); return v; } };
template <> struct Ld<4> { static __device__ __forceinline__ uint32_t ld(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; } };
template <> struct Ld<8> { static __device__ __forceinline__ uint32_t ld(uint32_t a) { uint32_t v, w; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v), "=r"(w) : "r"(a) : "memory"); return v ^ w; } };
template <> struct Ld<16> { static __device__ __forceinline__ uint32_t ld(uint32_t a) { uint32_t v, w, x, y; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v), "=r"(w), "=r"(x), "=r"(y) : "r"(a) : "memory"); return v ^ w ^ x ^ y; } };

template <int BYTES>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, int iters, long long* clk, uint32_t stride, uint32_t mask) {
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i * 2654435761u;
    __syncthreads();
    uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * BYTES + (threadIdx.x >> 5) * 32 * BYTES;
    uint32_t acc = 0, off = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) { acc ^= Ld<BYTES>::ld(base + off); off = (off + stride) & mask; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int BYTES> void run(int threads) {
    uint32_t* out; long long* clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    const int iters = 4000;
    cudaFuncSetAttribute(k<BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<BYTES><<<148, threads, 65536>>>(out, 10, clk, 1024 * BYTES, 0x7FFF & ~(1024 * BYTES - 1));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<BYTES><<<148, threads, 65536>>>(out, iters, clk, 1024 * BYTES, 0x7FFF & ~(1024 * BYTES - 1));
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    const double lds = (double)iters * 16 * (threads / 32);
    printf("width %2d B threads %4d: %.3f warp-LDS/clk/SM, %.1f B/clk/SM, %.2f TB/s total (%.3f ms, %s)\n", BYTES, threads,
           lds / c, lds * 32 * BYTES / c, 148.0 * lds * 32 * BYTES / (ms * 1e-3) / 1e12, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(clk);
}

int main() {
    for (int th : {256, 512, 768, 1024}) { run<2>(th); run<4>(th); run<8>(th); run<16>(th); }
    return 0;
}
