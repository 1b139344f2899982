// Small element-wise kernels around the scoring kernels: score epilogue after a cross-GPU reduction,
// prediction column, row-major -> column-major staging transpose, exact quantile (radix select).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ifb_internal.h"

namespace ifb {

namespace {
// SM count of the current device (grids are sized in multiples of it), cached per device
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return device_sm_count(dev);
    if (!cached[dev]) cached[dev] = device_sm_count(dev);
    return cached[dev];
}
}  // namespace

namespace {

// IF/IsolationForestModel.scala:137-138 applied to an already reduced f32 path-length sum.
__global__ void finalize_kernel(const float *__restrict__ path_sum, int64_t n, float total_trees, float avg_path,
                                double *__restrict__ scores) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float e = __fdiv_rn(path_sum[i], total_trees);
        const float z = __fdiv_rn(-e, avg_path);
        scores[i] = exp2((double)z);
    }
}

// Second half of the fused reduce-scatter: partial sums of `world` tree shards, added in RANK ORDER (a fixed,
// reproducible order -- unlike a ring all-reduce), then the same epilogue.
__global__ void finalize_gathered_kernel(const float *__restrict__ partials, int world, int64_t rows, float total_trees,
                                         float avg_path, double *__restrict__ scores) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < world; r++) s = s + partials[(int64_t)r * rows + i];
        const float e = __fdiv_rn(s, total_trees);
        const float z = __fdiv_rn(-e, avg_path);
        scores[i] = exp2((double)z);
    }
}

// ---- device-side barrier over peer memory (fused tree-sharded layout) ----
struct PeerFlags {
    uint32_t *p[kMaxScatterRanks];
};
__global__ void peer_signal_kernel(PeerFlags f, int world, int rank, uint32_t epoch) {
    // launched after the scatter kernel on the same stream: its peer stores are complete; publish the epoch
    const int o = threadIdx.x;
    if (o < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[o] + rank), "r"(epoch) : "memory");
    }
}
__global__ void peer_wait_kernel(const uint32_t *flags, int world, uint32_t epoch) {
    const int r = threadIdx.x;
    if (r < world) {
        uint32_t v = 0;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (true) {   // bounded to 30 s: a lost peer must surface as an error, not hang the GPU
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
            if ((int32_t)(v - epoch) >= 0) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 30000000000ull) __trap();
        }
    }
}

// IF/IsolationForestModel.scala:143-148
__global__ void predict_kernel(const double *__restrict__ scores, int64_t n, double thr, double *__restrict__ labels) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        labels[i] = (thr > 0.0 && scores[i] >= thr) ? 1.0 : 0.0;
}

// [n][ld_in] row-major  ->  [d][ld_out] column-major, 32x32 tiles through padded shared memory.
__global__ void transpose_rm_to_cm_kernel(const float *__restrict__ in, int64_t n, int32_t d, int64_t ld_in,
                                          float *__restrict__ out, int64_t ld_out) {
    __shared__ float tile[32][33];
    const int64_t row_tiles = (n + 31) / 32;
    const int col_tiles = (d + 31) / 32;
    const int64_t total = row_tiles * col_tiles;
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        const int64_t r0 = (t / col_tiles) * 32;
        const int c0 = (int)(t % col_tiles) * 32;
        for (int j = threadIdx.y; j < 32; j += blockDim.y) {
            const int64_t r = r0 + j;
            const int c = c0 + threadIdx.x;
            tile[j][threadIdx.x] = (r < n && c < d) ? in[r * ld_in + c] : 0.f;
        }
        __syncthreads();
        for (int j = threadIdx.y; j < 32; j += blockDim.y) {
            const int c = c0 + j;
            const int64_t r = r0 + threadIdx.x;
            if (r < n && c < d) out[(int64_t)c * ld_out + r] = tile[threadIdx.x][j];
        }
        __syncthreads();
    }
}

// ---- exact order statistic of non-negative doubles by 8-bit MSD radix select -----------------------
// Scores are in (0, 1], so their IEEE-754 bit patterns order like unsigned integers.  For generality
// the key transform below handles any finite double.
__device__ __forceinline__ unsigned long long key_of(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

struct SelectState {
    unsigned long long prefix;     // key bits decided so far
    unsigned long long rank;       // 0-based rank still to find inside the prefix bucket
    unsigned int hist[256];
    unsigned long long count_ge;
};

__global__ void select_hist_kernel(const double *__restrict__ v, int64_t n, int shift, SelectState *st) {
    __shared__ unsigned int h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const unsigned long long prefix = st->prefix;
    const unsigned long long mask = shift >= 56 ? 0ull : (~0ull << (shift + 8));
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = key_of(v[i]);
        if ((k & mask) == (prefix & mask)) atomicAdd(&h[(k >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (h[i]) atomicAdd(&st->hist[i], h[i]);
}

__global__ void select_pick_kernel(int shift, SelectState *st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long r = st->rank;
    int b = 0;
    for (; b < 256; b++) {
        const unsigned long long c = st->hist[b];
        if (r < c) break;
        r -= c;
    }
    if (b > 255) b = 255;
    st->rank = r;
    st->prefix |= ((unsigned long long)b) << shift;
    for (int i = 0; i < 256; i++) st->hist[i] = 0;
}

__global__ void count_ge_kernel(const double *__restrict__ v, int64_t n, const SelectState *st,
                                unsigned long long *count) {
    const unsigned long long kthr = st->prefix;
    unsigned long long local = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        local += key_of(v[i]) >= kthr ? 1ull : 0ull;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

}  // namespace

int launch_finalize(const float *path_sum, int64_t n_rows, int32_t total_trees, float avg_path, double *scores,
                    cudaStream_t stream) {
    if (n_rows == 0) return IFB_OK;
    const int grid = (int)std::min<int64_t>((n_rows + 255) / 256, sm_count() * 16);
    finalize_kernel<<<grid, 256, 0, stream>>>(path_sum, n_rows, (float)total_trees, avg_path, scores);
    IFB_CUDA(cudaGetLastError());
    count_launch();
    return IFB_OK;
}

int launch_finalize_gathered(const float *partials, int32_t world, int64_t rows_local, int32_t total_trees, float avg_path,
                             double *scores, cudaStream_t stream) {
    if (rows_local == 0) return IFB_OK;
    const int grid = (int)std::min<int64_t>((rows_local + 255) / 256, sm_count() * 16);
    finalize_gathered_kernel<<<grid, 256, 0, stream>>>(partials, world, rows_local, (float)total_trees, avg_path, scores);
    IFB_CUDA(cudaGetLastError());
    count_launch();
    return IFB_OK;
}

int launch_peer_signal(int world, int rank, uint32_t *const *peer_flags, uint32_t epoch, cudaStream_t stream) {
    PeerFlags f;
    for (int i = 0; i < kMaxScatterRanks; i++) f.p[i] = i < world ? peer_flags[i] : nullptr;
    peer_signal_kernel<<<1, 32, 0, stream>>>(f, world, rank, epoch);
    IFB_CUDA(cudaGetLastError());
    count_launch();
    return IFB_OK;
}
int launch_peer_wait(int world, const uint32_t *local_flags, uint32_t epoch, cudaStream_t stream) {
    peer_wait_kernel<<<1, 32, 0, stream>>>(local_flags, world, epoch);
    IFB_CUDA(cudaGetLastError());
    count_launch();
    return IFB_OK;
}

int launch_predict(const double *scores, int64_t n_rows, double threshold, double *labels, cudaStream_t stream) {
    if (n_rows == 0) return IFB_OK;
    const int grid = (int)std::min<int64_t>((n_rows + 255) / 256, sm_count() * 16);
    predict_kernel<<<grid, 256, 0, stream>>>(scores, n_rows, threshold, labels);
    IFB_CUDA(cudaGetLastError());
    count_launch();
    return IFB_OK;
}

int launch_transpose(const float *in, int64_t n, int32_t d, int64_t ld_in, float *out, int64_t ld_out,
                     cudaStream_t stream) {
    if (n == 0) return IFB_OK;
    const int64_t tiles = ((n + 31) / 32) * ((d + 31) / 32);
    const int grid = (int)std::min<int64_t>(tiles, sm_count() * 32);
    transpose_rm_to_cm_kernel<<<grid, dim3(32, 8), 0, stream>>>(in, n, d, ld_in, out, ld_out);
    IFB_CUDA(cudaGetLastError());
    count_launch();
    return IFB_OK;
}

// value = element of 0-based rank `rank0` of the sorted scores; also #(score >= value).
int launch_select(const double *scores, int64_t n, int64_t rank0, double *value, unsigned long long *count_ge,
                  cudaStream_t stream) {
    SelectState *st = nullptr;
    IFB_CUDA(cudaMallocAsync((void **)&st, sizeof(SelectState), stream));
    IFB_CUDA(cudaMemsetAsync(st, 0, sizeof(SelectState), stream));
    unsigned long long r = (unsigned long long)rank0;
    IFB_CUDA(cudaMemcpyAsync(&st->rank, &r, sizeof r, cudaMemcpyHostToDevice, stream));
    const int grid = (int)std::min<int64_t>((n + 255) / 256, sm_count() * 8);
    for (int shift = 56; shift >= 0; shift -= 8) {
        select_hist_kernel<<<grid, 256, 0, stream>>>(scores, n, shift, st);
        select_pick_kernel<<<1, 32, 0, stream>>>(shift, st);
        count_launch(2);
    }
    count_ge_kernel<<<grid, 256, 0, stream>>>(scores, n, st, &st->count_ge);
    count_launch();
    IFB_CUDA(cudaGetLastError());
    SelectState h;
    IFB_CUDA(cudaMemcpyAsync(&h, st, sizeof h, cudaMemcpyDeviceToHost, stream));
    IFB_CUDA(cudaStreamSynchronize(stream));
    IFB_CUDA(cudaFreeAsync(st, stream));
    // invert key_of
    unsigned long long k = h.prefix;
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    double v;
    memcpy(&v, &b, 8);
    *value = v;
    *count_ge = h.count_ge;
    return IFB_OK;
}

}  // namespace ifb
