// Internal declarations shared by the translation units of libifb200.so (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "ifb200.h"

namespace ifb {

void set_error(const char *fmt, ...);

#define IFB_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            ::ifb::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,      \
                             __LINE__, cudaGetErrorString(_e));                                 \
            return _e == cudaErrorMemoryAllocation ? IFB_ENOMEM : IFB_ECUDA;                    \
        }                                                                                       \
    } while (0)

#define IFB_REQUIRE(cond, ...)             \
    do {                                   \
        if (!(cond)) {                     \
            ::ifb::set_error(__VA_ARGS__); \
            return IFB_EINVAL;             \
        }                                  \
    } while (0)

void count_launch(int n = 1);
void tune_mempool(int device);   // api.cu: the device's stream-ordered pool keeps freed blocks

// RAII device guard
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ------------------------------------------------------------------------------------------------
// Device-side node tables ("kernel layout").
//
// Standard forest.  Every tree is renumbered in level order (BFS) with the two children of a node
// adjacent (right = left + 1).  One node = two 32-bit words kept in two parallel arrays so that a warp
// whose lanes sit on different nodes of one level touches distinct banks for the top levels:
//     val [g] : internal: smallest f32 >= splitValue  (x < splitValue  <=>  x < val, x being f32)
//               leaf    : (float)depth + avgPathLength(numInstances)   (the value pathLength returns)
//     meta[g] : (child << 16) | feature   where child = WORD index of the LEFT child in the chunk-local
//               val[] array; leaves store feature = d (the pseudo column of NaNs every row tile carries)
//               and child = self - 1, so "next = child + !(x < val)" maps a leaf onto itself: walks run
//               a fixed number of steps with no per-level branch.  Levels 0 and 1 of every tree are also
//               flattened into a kernel-parameter table (constant bank, warp-uniform loads).
// Trees are grouped into chunks whose tables fit in shared memory next to the row tiles.
// ------------------------------------------------------------------------------------------------
struct StdChunk {
    int32_t tree_begin = 0, tree_end = 0;
    int32_t node_begin = 0, node_count = 0;  // slice of the device arrays (node_count includes 1 pad slot)
};

// Node record of the wide extended kernel: everything a visit needs in one 32-byte load, including the weight
// slots of BOTH children, so the next node's weight row can be requested as soon as the decision is known.
struct WideNode {
    double off;        // split offset (internal)
    float wnorm;       // ||w||_2 inflated, rounded up to f32 (bound of the f32 tier)
    float leaf;        // (float)depth + c(numInstances) at leaves
    int32_t cbase;     // tree-local BFS index of the left child, -1 at leaves
    int32_t slot_l;    // weight slot of the left child (-1: it is a leaf)
    int32_t slot_r;    // weight slot of the right child
    int32_t slot;      // own weight slot (-1 at leaves)
};

}  // namespace ifb

namespace ifb {
struct RankPlan;   // score_std_rank.cu
}

struct ifb_forest {
    int32_t device = 0;
    bool extended = false;
    int32_t num_trees = 0;
    int32_t num_samples = 0;
    int32_t total_num_features = -1;
    int32_t max_feature_index = -1;
    int32_t max_depth = 0;
    int32_t max_nnz = 1;
    float avg_path_norm = 0.f;  // c(numSamples)

    // persisted layout (host copy; what ifb_forest_export returns)
    std::vector<int32_t> node_off, left, right, feature;
    std::vector<double> threshold, offset;
    std::vector<int64_t> num_instances, hp_off;
    std::vector<int32_t> hp_idx;
    std::vector<float> hp_w;
    // Forests fitted on the device with wide dense hyperplanes (k == d > 64) never bring the weights to the host:
    // hp_idx / hp_w stay empty, ifb_forest_export gathers them from d_ext_w on demand (forest.cu).
    bool hp_lazy = false;
    std::vector<int32_t> lazy_slot_of_node;   // [nodes] pre-order row -> weight slot (row of d_ext_w), -1 at leaves

    // ---- standard kernel layout ----
    // host mirrors are kept so that chunking can be re-planned for a different feature count
    std::vector<float> h_val;
    std::vector<uint32_t> h_meta_feat;   // feature per node (d placeholder = 0xFFFFFFFF for leaves)
    std::vector<int32_t> h_child;        // tree-local BFS index of left child, or -1 for leaf
    std::vector<uint8_t> h_depth;        // depth of the node
    std::vector<int32_t> bfs_off;        // [T+1] first BFS node of each tree (== node_off)

    // plans are keyed by the feature count d of the scored matrix (pseudo column index + smem budget)
    struct StdPlan {
        int32_t d = -1;
        int32_t rows_per_tile = 0;
        int32_t stages = 2;              // row-tile ring depth (1 when rows are wide: bigger tiles beat double buffering)
        std::vector<ifb::StdChunk> chunks;
        float *d_val = nullptr;
        uint32_t *d_meta = nullptr;
        uint32_t *d_tree_root = nullptr;  // [T] chunk-local word index of each tree's root
        int64_t total_words = 0;
        std::vector<unsigned char> h_top;  // per chunk: the TopTable kernel parameter (tree levels 0 and 1)
    };
    std::mutex plan_mu;
    std::vector<StdPlan *> std_plans;
    std::vector<ifb::RankPlan *> rank_plans;   // rank-word tables for matrices of <= 32 features (score_std_rank.cu)
    // global-memory tables for the generic fallback kernel (rows too wide for a shared-memory plan); lazily built
    float *d_gval = nullptr;        // [nodes] BFS order: threshold (f32 ceil) or leaf value
    int32_t *d_gfeat = nullptr;     // [nodes] feature, -1 at leaves
    int32_t *d_gchild = nullptr;    // [nodes] tree-local BFS index of the left child, -1 at leaves
    int32_t *d_groot = nullptr;     // [T+1] first BFS node of each tree

    // ---- extended kernel layout ----
    // BFS order per tree; hyperplanes stored densely per internal node with a fixed width k = max_nnz.
    float *d_ext_w = nullptr;          // [internal_slots * k]
    int32_t *d_ext_idx = nullptr;      // [internal_slots * k] (nullptr when every hyperplane is the identity 0..k-1)
    double *d_ext_off = nullptr;       // [nodes] split offset (internal) / unused (leaf)
    float *d_ext_leaf = nullptr;       // [nodes] (float)depth + c(n) at leaves
    int32_t *d_ext_child = nullptr;    // [nodes] left child (tree-local BFS) or -1 at leaves
    int32_t *d_ext_hp = nullptr;       // [nodes] hyperplane slot (index into w/idx rows) or -1 at leaves
    int32_t *d_ext_len = nullptr;      // [nodes] number of hyperplane terms (0 at leaves)
    double *d_ext_wabs = nullptr;      // [internal_slots] sum |w_i| of the slot (rounding bound of the wide kernel)
    double *d_ext_wnorm = nullptr;     // [internal_slots] ||w||_2 of the slot, inflated (f32 fast-path bound)
    void *d_ext_wide_nodes = nullptr;  // [nodes] ifb::WideNode records (wide kernel: one load per visit)
    int32_t *d_ext_tree_slot = nullptr;  // [T] weight slot of each tree's root (-1 if the root is a leaf)
    int64_t *d_ext_tree_node = nullptr;  // [T+1]
    bool ext_dense_identity = false;
    int64_t ext_internal_slots = 0;
    // fully-extended forests with k <= 64: one self-contained blob per tree (header, child[], slot[], leaf[],
    // off[], w[internal][D+4]) that score_ext_dense_kernel streams into shared memory with a bulk async copy
    unsigned char *d_ext_blob = nullptr;
    int64_t *d_ext_blob_off = nullptr;   // [T+1] byte offsets (16-byte aligned)
    unsigned char *d_ext_arena = nullptr; // the one allocation every d_ext_* table above is a slice of
    bool ext_blob_tried = false;         // ensure_ext_blob ran (the blobs are built on first use of the dense kernel)
    int32_t ext_blob_D = 0;              // padded hyperplane width (8/16/32/64), 0 = no blob layout (yet)
    int64_t ext_blob_max = 0;            // largest blob in bytes
    bool ext_w_safe = false;             // every hyperplane weight is normal with 2^-60 <= |w| <= 2^40

    // ---- tensor-core layout of fully-extended forests (score_ext_tc.cu): every hyperplane of the forest is one
    // column of a [rows x nodes] GEMM evaluated by tcgen05.mma on fp16 hi/lo splits; 256-column blocks of whole trees
    bool tc_ok = false;                  // tables below are valid
    int32_t tc_k = 0, tc_kp = 0;         // hyperplane width and its padding to the K chunk (32)
    int32_t tc_blocks = 0;               // 256-column blocks
    void *d_tc_wh = nullptr;             // fp16 [tc_blocks*256][tc_kp]: hi part of the scaled weights
    void *d_tc_wl = nullptr;             // fp16, lo part
    unsigned char *d_tc_meta = nullptr;  // [tc_blocks] TcBlockMeta (node records, leaf values, tree roots)
    int32_t *d_tc_col_slot = nullptr;    // [tc_blocks*256] weight slot of the column's node (-1: padding column)
    double *d_tc_col_off = nullptr;      // [tc_blocks*256] f64 split offset of the column's node
    int32_t *d_tc_flag = nullptr;        // set by the column-preparation kernel when a weight row is not fp16-safe
    bool tc_sparse = false;              // hyperplanes narrower than the matrix: columns are zero-padded, the exact path
    int32_t *d_tc_slot_len = nullptr;    // reads the stored terms (d_ext_w / d_ext_idx) with these per-slot counts
    unsigned char *d_tc_arena = nullptr;

    int64_t device_bytes = 0;

    ~ifb_forest();
};

namespace ifb {

// forest.cu
int build_standard_tables(ifb_forest *f);
// Hyperplane weights that already live on the device in the builder's layout (fit.cu): row src_row[g] of `w`
// (k floats each) belongs to pre-order node g; indices are the identity 0..k-1.
struct DeviceHyperplanes {
    const float *w = nullptr;
    const int64_t *src_row = nullptr;   // host, [nodes], -1 at leaves
};
int build_extended_tables(ifb_forest *f, const DeviceHyperplanes *dev = nullptr);
int create_extended_from_device(int32_t device, int32_t num_trees, const int32_t *node_off, const int32_t *left,
                                const int32_t *right, const int64_t *num_instances, const double *offset, int32_t k,
                                const DeviceHyperplanes &dev, int32_t num_samples, int32_t total_num_features,
                                ifb_forest **out);
int ensure_ext_blob(ifb_forest *f);   // dense CUDA-core kernel's per-tree blobs, built on first use
int get_std_plan(ifb_forest *f, int32_t d, ifb_forest::StdPlan **out);   // *out = nullptr: no smem plan, use generic
int ensure_std_generic_tables(ifb_forest *f);
int launch_score_standard_generic(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld,
                                  int32_t layout, double *scores, int32_t *depth_sum, float *path_sum,
                                  bool accumulate_only, cudaStream_t stream);

// tree-sharded multi-GPU scoring: the scoring kernel's epilogue writes this rank's per-row partial sums straight
// into the buffer of the rank that owns the row (peer memory over NVLink) -- a reduce-scatter fused into the kernel
constexpr int kMaxScatterRanks = 8;
struct ScatterTarget {
    int32_t world = 0, rank = 0;
    int64_t cut[kMaxScatterRanks + 1];   // rows [cut[o], cut[o+1]) are owned by rank o
    float *peer[kMaxScatterRanks];       // peer[o]: rank o's buffer [world][rows_o] (device pointer mapped here)
};

// score_std.cu
size_t std_top_table_bytes();
int std_top_table_max_trees();
int std_rows_per_box(int rows_per_tile, int stages);   // rows per TMA box / shared-memory sub-tile of a plan
void std_fill_top_table(void *dst, const float *val, const uint32_t *meta, const uint32_t *root_byte, int n_trees,
                        int rows_per_box);
int launch_score_standard(const ifb_forest *f, ifb_forest::StdPlan *plan, const float *X, int64_t n_rows,
                          int32_t d, int64_t ld, int32_t layout, double *scores, int32_t *depth_sum,
                          float *path_sum, bool accumulate_only, cudaStream_t stream,
                          const ScatterTarget *scatter = nullptr);
// score_std_rank.cu: the standard walk on per-feature ranks (opt-in with IFB_STD_RANK=1; d <= 32, finite thresholds,
// no depth sums); a plan with rank_plan_chunks() == 0 means "does not qualify, use score_std.cu"
bool std_rank_enabled();   // IFB_STD_RANK=1 (score_std.cu)
int get_rank_plan(ifb_forest *f, int32_t d, RankPlan **out);
int rank_plan_chunks(const RankPlan *rp);
void free_rank_plans(ifb_forest *f);
int launch_score_standard_rank(const ifb_forest *f, RankPlan *rp, const float *X, int64_t n_rows, int32_t d, int64_t ld,
                               double *scores, float *path_sum, bool accumulate_only, cudaStream_t stream,
                               const ScatterTarget *scatter);
// score_ext.cu
int launch_score_extended(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld,
                          int32_t layout, double *scores, int32_t *depth_sum, float *path_sum,
                          bool accumulate_only, cudaStream_t stream);
// score_ext_tc.cu: tensor-core path of fully-extended forests; returns -1 when the forest / call does not qualify
int build_ext_tc_tables(ifb_forest *f, const std::vector<int32_t> &child, const std::vector<int32_t> &hp,
                        const std::vector<float> &leaf, const std::vector<double> &off, const std::vector<uint8_t> &depth,
                        const std::vector<int32_t> &len);
int launch_score_extended_tc(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                             double *scores, int32_t *depth_sum, float *path_sum, bool accumulate_only,
                             cudaStream_t stream, float *probe_out = nullptr);
// epilogue.cu
int launch_finalize(const float *path_sum, int64_t n_rows, int32_t total_trees, float avg_path, double *scores,
                    cudaStream_t stream);
int launch_predict(const double *scores, int64_t n_rows, double threshold, double *labels, cudaStream_t stream);
int launch_peer_signal(int world, int rank, uint32_t *const *peer_flags, uint32_t epoch, cudaStream_t stream);
int launch_peer_wait(int world, const uint32_t *local_flags, uint32_t epoch, cudaStream_t stream);
int launch_finalize_gathered(const float *partials, int32_t world, int64_t rows_local, int32_t total_trees, float avg_path,
                             double *scores, cudaStream_t stream);
int launch_transpose(const float *in, int64_t n, int32_t d, int64_t ld_in, float *out, int64_t ld_out,
                     cudaStream_t stream);
int launch_select(const double *scores, int64_t n, int64_t rank0, double *value, unsigned long long *count_ge,
                  cudaStream_t stream);

float avg_path_length_host(int64_t n);
int device_smem_optin(int device);
int device_sm_count(int device);

}  // namespace ifb
