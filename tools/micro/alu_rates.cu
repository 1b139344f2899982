// Micro-benchmark (developer tool): issue rate per scheduler (SMSP) of the ALU / FMA-pipe instructions the tensor-core
// epilogue is made of, on sm_100a.  16 warps per CTA (4 per SMSP), 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rates alu_rates.cu && ./alu_rates
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    if (OP == 0) asm volatile("shf.l.wrap.b32 %0, %1, %2, 1;" : "=r"(r) : "r"(b), "r"(a));          // funnel shift, imm count
    else if (OP == 1) asm volatile("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(c));                   // shift by register
    else if (OP == 2) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 3) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 4) {
        float f;
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(f) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));
        r = __float_as_uint(f);
    } else if (OP == 5) {
        float f;
        asm volatile("min.abs.f32 %0, %1, %2, %3;" : "=f"(f) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));
        r = __float_as_uint(f);
    } else if (OP == 6) {
        float f;
        asm volatile("min.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
        r = __float_as_uint(f);
    } else if (OP == 7) asm volatile("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(c));   // funnel, reg count
    else if (OP == 8) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else if (OP == 9) asm volatile("{ .reg .pred p; setp.lt.s32 p, %1, %2; selp.u32 %0, %3, %1, p; }" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 10) {
        float f;
        asm volatile("mul.rn.sat.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
        r = __float_as_uint(f);
    } else if (OP == 11) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else if (OP == 12) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

template <int OP>
__global__ void k(unsigned long long *out, uint32_t *sink, uint32_t seed) {
    uint32_t x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) x[i] = seed * (i + 1) + threadIdx.x;
    const uint32_t b = seed | 1u, c = (seed >> 3) & 7u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) x[i] = op<OP>(x[i], b, c);
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc ^= x[i];
    if (acc == 0x1234567u) sink[0] = acc;
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int OP>
void run(const char *name, unsigned long long *out, uint32_t *sink) {
    k<OP><<<148, 512>>>(out, sink, 12345u);
    cudaDeviceSynchronize();
    k<OP><<<148, 512>>>(out, sink, 12345u);
    cudaDeviceSynchronize();
    unsigned long long h;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    const double per_smsp = 4.0 * CHAINS * ITERS;   // warp instructions per scheduler
    printf("%-34s %.2f clk per warp instruction per scheduler\n", name, (double)h / per_smsp);
}

int main() {
    unsigned long long *out;
    uint32_t *sink;
    cudaMalloc(&out, 148 * 8);
    cudaMalloc(&sink, 4);
    run<0>("SHF.L.W (funnel, imm)", out, sink);
    run<7>("SHF.L.W (funnel, reg count)", out, sink);
    run<1>("SHF.L (shift by register)", out, sink);
    run<2>("LOP3", out, sink);
    run<8>("IADD", out, sink);
    run<3>("IMAD", out, sink);
    run<11>("IMAD.HI", out, sink);
    run<4>("FFMA", out, sink);
    run<10>("FMUL.SAT", out, sink);
    run<6>("FMNMX", out, sink);
    run<5>("FMNMX3 (min.abs 3-input)", out, sink);
    run<9>("ISETP + SEL", out, sink);
    run<12>("PRMT", out, sink);
    return 0;
}
