// Micro-benchmark (developer tool, not part of the library): TMEM -> register read rate of tcgen05.ld.32x32b.x32 on
// sm_100a with 4 / 8 / 16 / 24 warps of one CTA reading their lane quarter back to back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_read_bw tmem_read_bw.cu && ./tmem_read_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void k(int iters, unsigned long long *out, uint32_t *sink) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tptr)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tptr;
    const int q = warp & 3;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        uint32_t v[32];
        tmem_ld32(base + ((uint32_t)(q * 32) << 16) + (uint32_t)(((i + warp) & 15) * 32), v);
#pragma unroll
        for (int j = 0; j < 32; j++) acc ^= v[j];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}

int main() {
    unsigned long long *out;
    uint32_t *sink;
    cudaMalloc(&out, 148 * 8);
    cudaMalloc(&sink, 4);
    const int iters = 4096;
    for (int warps : {4, 8, 16, 24}) {
        k<<<148, warps * 32>>>(iters, out, sink);
        cudaDeviceSynchronize();
        k<<<148, warps * 32>>>(iters, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[148];
        cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
        const double cyc = (double)h[0];
        const double bytes = (double)warps * iters * 4096.0;
        printf("%2d warps: %s  %.0f cycles for %d loads per warp: %.1f B/clk/SM, %.1f clk per 4 KB load per quarter\n", warps,
               cudaGetErrorString(e), cyc, iters, bytes / cyc, cyc / iters / (warps / 4.0));
    }
    return 0;
}
