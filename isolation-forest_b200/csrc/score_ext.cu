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

}  // namespace

int launch_score_extended(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                          double *scores, int32_t *depth_sum, float *path_sum, bool accumulate_only,
                          cudaStream_t stream) {
    if (n_rows == 0) return IFB_OK;
    IFB_REQUIRE(f->max_feature_index < d, "forest reads feature index %d but the matrix has only %d columns",
                f->max_feature_index, d);
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
