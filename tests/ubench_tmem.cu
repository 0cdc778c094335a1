// Micro-benchmark (not part of the library): tcgen05.ld / tcgen05.st throughput of one SM as a function of the number of
// warps that issue them.  Build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_tmem tests/ubench_tmem.cu
// Output feeds DESIGN.md section 4.1 (what bounds an accumulator pass of the epilogues).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// mode 0: ld + wait per instruction; 1: 4 lds in flight before one wait; 2: st + wait per instruction; 3: 4 sts then wait
__global__ void __launch_bounds__(512, 1) k(int mode, int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 64);
    uint32_t v[4][16];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int e = 0; e < 16; ++e) v[b][e] = threadIdx.x + e + b;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                ld16(base + 16 * b, v[b]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += v[b][3];
            }
        } else if (mode == 1) {
#pragma unroll
            for (int b = 0; b < 4; ++b) ld16(base + 16 * b, v[b]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int b = 0; b < 4; ++b) acc += v[b][3];
        } else if (mode == 2) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                st16(base + 16 * b, v[b]);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) st16(base + 16 * b, v[b]);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

int main() {
    long long* out;
    uint32_t* sink;
    cudaMalloc(&out, 8);
    cudaMalloc(&sink, 4096);
    const char* names[4] = {"ld x16 + wait each", "4 x ld x16, one wait", "st x16 + wait each", "4 x st x16, one wait"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            const int iters = 2000;
            k<<<1, warps * 32, 0>>>(mode, iters, out, sink);
            long long c = 0;
            cudaError_t e = cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            const double bytes = (double)iters * 4 * 2048 * warps;  // 4 instructions x 32 lanes x 64 B per warp and iteration
            printf("%-22s warps %2d: %8.1f cycles per 128 KB-equivalent pass, %6.1f B/clk per SM\n", names[mode], warps,
                   (double)c / bytes * 131072.0, bytes / (double)c);
        }
    return 0;
}
