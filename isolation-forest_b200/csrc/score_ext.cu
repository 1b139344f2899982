// Extended isolation-forest scoring for sm_100a.
//
// Replaces ExtendedIsolationForestModel.transform's UDF body
// (IF/extended/ExtendedIsolationForestModel.scala:114-120), ExtendedIsolationTree.pathLength
// (IF/extended/ExtendedIsolationTree.scala:283-355) and SplitHyperplane.dot
// (IF/extended/ExtendedUtils.scala:36-55).
//
// Arithmetic contract of a node visit (bit-for-bit the reference's):
//     sum: f64 = 0;  for i ascending:  sum += (double) __fmul_rn(w[i], x[idx[i]])   // f32 product, ONE
//     go left iff sum < offset (f64, strict)                                        // rounding, no FMA
//
// Kernels
//   score_ext_dense_kernel<D>   every hyperplane is (0..D-1) (fully extended forests, the BASELINE configs
//                               with extensionLevel = d-1, D <= 64): a thread keeps its row in registers,
//                               one tree at a time is staged into shared memory with a 1-D bulk async
//                               copy (cp.async.bulk + mbarrier, double buffered), node rows padded to
//                               D+4 floats so that per-lane LDS.128 gathers spread over the banks.
//   score_ext_generic_kernel    any width / index pattern / d: row tile in shared memory, node tables read
//                               through L1/L2.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "ifb_internal.h"

namespace ifb {

namespace {

struct ScoreExtParams {
    const float *X;
    int64_t n_rows, ld;
    int32_t d;
    int32_t layout;
    const float *w;
    const int32_t *idx;       // may be null (dense identity)
    const double *off;
    const float *leaf;
    const int32_t *child, *hp, *len;
    const int64_t *tree_node;  // [T+1]
    int32_t k;
    int32_t num_trees, total_trees;
    float avg_path;
    int32_t accumulate_only;
    double *scores;
    float *path_sum;
    int32_t *depth_sum;
};

// Generic kernel: one thread per row; the thread's features are read from a shared-memory tile laid out
// [feature][row] (bank = row % 32, conflict free for any per-lane feature index).
template <int R>
__global__ void __launch_bounds__(R) score_ext_generic_kernel(const ScoreExtParams p, int use_smem) {
    extern __shared__ float xs[];
    const int tid = threadIdx.x;
    const int64_t n_tiles = (p.n_rows + R - 1) / R;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * R + tid;
        const bool live = row < p.n_rows;
        const int64_t rs = p.layout == IFB_COL_MAJOR ? 1 : p.ld;
        const int64_t cs = p.layout == IFB_COL_MAJOR ? p.ld : 1;
        if (use_smem) {
            __syncthreads();
            if (p.layout == IFB_COL_MAJOR) {
                for (int f = 0; f < p.d; f++) xs[f * R + tid] = live ? __ldg(p.X + row + (int64_t)f * p.ld) : 0.f;
            } else {
                // row-major source: consecutive threads read consecutive floats of the tile's rows
                const int64_t base = tile * R * p.ld;
                const int64_t lim = p.n_rows * p.ld;
                for (int64_t e = tid; e < (int64_t)R * p.d; e += R) {
                    const int r = (int)(e / p.d), f = (int)(e % p.d);
                    const int64_t g = base + (int64_t)r * p.ld + f;
                    xs[f * R + r] = g < lim ? __ldg(p.X + g) : 0.f;
                }
            }
            __syncthreads();
        }
        if (!live) continue;
        const float *xg = p.X + row * rs;
        float s = p.accumulate_only ? p.path_sum[row] : 0.f;
        int32_t dsum = (p.accumulate_only && p.depth_sum) ? p.depth_sum[row] : 0;
        for (int t = 0; t < p.num_trees; t++) {
            const int64_t base = p.tree_node[t];
            int32_t node = 0;
            int32_t c = p.child[base];
            while (c >= 0) {
                const int64_t g = base + node;
                const int64_t slot = (int64_t)p.hp[g] * p.k;
                const int32_t len = p.len[g];
                double sum = 0.0;
                for (int i = 0; i < len; i++) {
                    const int32_t j = p.idx ? p.idx[slot + i] : i;
                    const float xv = use_smem ? xs[j * R + tid] : __ldg(xg + (int64_t)j * cs);
                    sum = __dadd_rn(sum, (double)__fmul_rn(p.w[slot + i], xv));
                }
                node = c + ((sum < p.off[g]) ? 0 : 1);
                c = p.child[base + node];
                dsum++;
            }
            s = s + p.leaf[base + node];
        }
        if (!p.accumulate_only) {
            const float e = __fdiv_rn(s, (float)p.total_trees);
            const float z = __fdiv_rn(-e, p.avg_path);
            p.scores[row] = exp2((double)z);
        }
        if (p.path_sum) p.path_sum[row] = s;
        if (p.depth_sum) p.depth_sum[row] = dsum;
    }
}


// ---- dense kernel -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32e(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_e(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32e(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_e(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32e(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_e(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32e(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < 0x7fffffffu; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32e(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32e(bar))
                 : "memory");
}

struct ScoreExtDenseParams {
    const float *X;
    int64_t n_rows, ld;
    int32_t d, layout;
    const unsigned char *blob;
    const int64_t *blob_off;
    int32_t num_trees, total_trees;
    int64_t blob_max;
    float avg_path;
    int32_t accumulate_only;
    double *scores;
    float *path_sum;
    int32_t *depth_sum;
};

// One thread owns one row, held in registers (D floats, zero padded); the trees stream through a 2-slot
// shared-memory ring, one self-contained blob per tree (forest.cu::build_extended_tables).
template <int D, int R>
__global__ void __launch_bounds__(R) score_ext_dense_kernel(const ScoreExtDenseParams p) {
    extern __shared__ __align__(128) unsigned char smem_e[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_e);
    const uint32_t slot_bytes = (uint32_t)((p.blob_max + 127) & ~127LL);
    unsigned char *ring = smem_e + 128;
    const int tid = threadIdx.x;
    constexpr int WS = D + 4;
    if (tid == 0) {
        mbar_init_e(&bars[0], 1);
        mbar_init_e(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t n_tiles = (p.n_rows + R - 1) / R;
    const int T = p.num_trees;
    // the (tile, tree) stream of this CTA is linear: item j uses ring slot j&1, parity (j>>1)&1
    int64_t my_tiles = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) my_tiles++;
    const int64_t n_items = my_tiles * T;
    auto issue = [&](int64_t j) {  // thread 0 only
        const int t = (int)(j % T);
        const int64_t b0 = p.blob_off[t], b1 = p.blob_off[t + 1];
        const int sl = (int)(j & 1);
        mbar_expect_e(&bars[sl], (uint32_t)(b1 - b0));
        bulk_g2s(ring + (size_t)sl * slot_bytes, p.blob + b0, (uint32_t)(b1 - b0), &bars[sl]);
    };
    if (tid == 0 && n_items > 0) issue(0);
    int64_t j = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * R + tid;
        const bool live = row < p.n_rows;
        float xr[D];
#pragma unroll
        for (int c = 0; c < D; c++) {
            float v = 0.f;
            if (live && c < p.d)
                v = p.layout == IFB_COL_MAJOR ? __ldg(p.X + (int64_t)c * p.ld + row) : __ldg(p.X + row * p.ld + c);
            xr[c] = v;
        }
        float s = (p.accumulate_only && live) ? p.path_sum[row] : 0.f;
        int32_t dsum = (p.accumulate_only && live && p.depth_sum) ? p.depth_sum[row] : 0;
        for (int t = 0; t < T; t++, j++) {
            const int sl = (int)(j & 1);
            if (tid == 0 && j + 1 < n_items) issue(j + 1);   // slot (j+1)&1 was released by the sync ending item j-1
            mbar_wait_e(&bars[sl], (uint32_t)((j >> 1) & 1));
            const unsigned char *B = ring + (size_t)sl * slot_bytes;
            const int32_t *hdr = reinterpret_cast<const int32_t *>(B);
            const int npad = hdr[2], ipad = hdr[3];
            const int32_t *child = reinterpret_cast<const int32_t *>(B + 16);
            const int32_t *slot = child + npad;
            const float *leaf = reinterpret_cast<const float *>(slot + npad);
            const double *off = reinterpret_cast<const double *>(leaf + npad);
            const float *w = reinterpret_cast<const float *>(off + ipad);
            int node = 0;
            int c = child[0];
            while (c >= 0) {
                const int hs = slot[node];
                const float4 *wp = reinterpret_cast<const float4 *>(w + (size_t)hs * WS);
                double sum = 0.0;
#pragma unroll
                for (int q = 0; q < D / 4; q++) {
                    const float4 w4 = wp[q];
                    // Float * Float -> Float (one rounding, no FMA), then += in Double, ascending index
                    sum = __dadd_rn(sum, (double)__fmul_rn(w4.x, xr[4 * q + 0]));
                    sum = __dadd_rn(sum, (double)__fmul_rn(w4.y, xr[4 * q + 1]));
                    sum = __dadd_rn(sum, (double)__fmul_rn(w4.z, xr[4 * q + 2]));
                    sum = __dadd_rn(sum, (double)__fmul_rn(w4.w, xr[4 * q + 3]));
                }
                node = c + ((sum < off[hs]) ? 0 : 1);
                c = child[node];
                dsum++;
            }
            s = s + leaf[node];
            __syncthreads();  // everyone is done with this slot before it is refilled two items later
        }
        if (live) {
            if (!p.accumulate_only) {
                const float e = __fdiv_rn(s, (float)p.total_trees);
                const float z = __fdiv_rn(-e, p.avg_path);
                p.scores[row] = exp2((double)z);
            }
            if (p.path_sum) p.path_sum[row] = s;
            if (p.depth_sum) p.depth_sum[row] = dsum;
        }
    }
}

template <int D>
int launch_dense(const ifb_forest *f, const ScoreExtDenseParams &p, cudaStream_t stream) {
    constexpr int R = 256;
    const size_t slot_bytes = (size_t)((f->ext_blob_max + 127) & ~127LL);
    const size_t smem = 128 + 2 * slot_bytes;
    const int smem_max = device_smem_optin(f->device);
    if (smem > (size_t)smem_max) return -1;  // caller falls back to the generic kernel
    const int sms = device_sm_count(f->device);
    int per_sm = (int)std::min<size_t>(2, (size_t)(smem_max + 1024) / (smem + 1024));
    per_sm = std::max(per_sm, 1);
    const int64_t n_tiles = (p.n_rows + R - 1) / R;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)sms * per_sm);
    IFB_CUDA(cudaFuncSetAttribute(score_ext_dense_kernel<D, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    score_ext_dense_kernel<D, R><<<grid, R, smem, stream>>>(p);
    IFB_CUDA(cudaGetLastError());
    count_launch();
    return IFB_OK;
}

}  // namespace

int launch_score_extended(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                          double *scores, int32_t *depth_sum, float *path_sum, bool accumulate_only,
                          cudaStream_t stream) {
    if (n_rows == 0) return IFB_OK;
    IFB_REQUIRE(f->max_feature_index < d, "forest reads feature index %d but the matrix has only %d columns",
                f->max_feature_index, d);
    if (f->ext_blob_D > 0 && d <= 64 && getenv("IFB_EXT_GENERIC") == nullptr) {
        ScoreExtDenseParams q;
        q.X = X; q.n_rows = n_rows; q.ld = ld; q.d = d; q.layout = layout;
        q.blob = f->d_ext_blob; q.blob_off = f->d_ext_blob_off;
        q.num_trees = f->num_trees; q.total_trees = f->num_trees; q.blob_max = f->ext_blob_max;
        q.avg_path = f->avg_path_norm; q.accumulate_only = accumulate_only ? 1 : 0;
        q.scores = scores; q.path_sum = path_sum; q.depth_sum = depth_sum;
        int rc;
        switch (f->ext_blob_D) {
            case 8: rc = launch_dense<8>(f, q, stream); break;
            case 16: rc = launch_dense<16>(f, q, stream); break;
            case 32: rc = launch_dense<32>(f, q, stream); break;
            default: rc = launch_dense<64>(f, q, stream); break;
        }
        if (rc >= 0) return rc;
    }
    ScoreExtParams p;
    p.X = X;
    p.n_rows = n_rows;
    p.ld = ld;
    p.d = d;
    p.layout = layout;
    p.w = f->d_ext_w;
    p.idx = f->ext_dense_identity ? nullptr : f->d_ext_idx;
    p.off = f->d_ext_off;
    p.leaf = f->d_ext_leaf;
    p.child = f->d_ext_child;
    p.hp = f->d_ext_hp;
    p.len = f->d_ext_len;
    p.tree_node = f->d_ext_tree_node;
    p.k = f->max_nnz;
    p.num_trees = f->num_trees;
    p.total_trees = f->num_trees;
    p.avg_path = f->avg_path_norm;
    p.accumulate_only = accumulate_only ? 1 : 0;
    p.scores = scores;
    p.path_sum = path_sum;
    p.depth_sum = depth_sum;
    constexpr int R = 128;
    const int sms = device_sm_count(f->device);
    const size_t tile_bytes = (size_t)R * d * 4;
    const int smem_max = device_smem_optin(f->device);
    const int use_smem = tile_bytes <= (size_t)smem_max / 2 ? 1 : 0;
    const size_t smem = use_smem ? tile_bytes : 0;
    int per_sm = use_smem ? std::max<int>(1, std::min<int>(8, (int)((size_t)smem_max / std::max<size_t>(tile_bytes, 1)))) : 8;
    const int64_t n_tiles = (n_rows + R - 1) / R;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)sms * per_sm);
    IFB_CUDA(cudaFuncSetAttribute(score_ext_generic_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)std::max<size_t>(smem, 1024)));
    score_ext_generic_kernel<R><<<grid, R, smem, stream>>>(p, use_smem);
    IFB_CUDA(cudaGetLastError());
    count_launch();
    return IFB_OK;
}

}  // namespace ifb
