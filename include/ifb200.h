/*
 * ifb200.h -- C ABI of the B200-native isolation-forest engine (libifb200.so).
 *
 * This is the drop-in boundary for the reference's scoring / fit hot path.  The reference
 * (linkedin/isolation-forest @ 10b5f0a) has no FFI seam of its own; each entry point below names the
 * reference code it replaces (paths relative to
 * isolation-forest/src/main/scala/com/linkedin/relevance/isolationforest/, abbreviated IF/).
 * INTEGRATION.md shows the JNI binding a maintainer would add on the Scala side.
 *
 * Conventions
 *  - plain C, no C++/torch types; every function returns an ifb_status (0 = OK) and never throws;
 *    the message of the last failure on the calling thread is ifb_last_error().
 *  - "host" pointers are ordinary (ideally pinned, see ifb_host_alloc) CPU memory; "device" pointers are
 *    CUDA device memory on the forest's device.  Feature matrices are f32, score vectors f64.
 *  - feature matrix layout: IFB_COL_MAJOR  x[r][c] = X[c*ld + r]  (ld >= n_rows; the layout
 *    BASELINE.json prescribes) or IFB_ROW_MAJOR  x[r][c] = X[r*ld + c]  (ld >= d; what a Spark
 *    DenseVector column is before transposition).  The cast Double->Float happens on the caller's side
 *    (`.toFloat`, IF/IsolationForestModel.scala:136, IF/IsolationForest.scala:54).
 *  - a forest handle is immutable after creation and may be scored from many host threads at once
 *    (the reference shares the broadcast forest across executor task threads,
 *    IF/IsolationForestModel.scala:129-142); buffers are never retained past a call.
 *  - node tables use the reference's persisted layout (IF/IsolationForestModelReadWrite.scala:60-67,
 *    IF/extended/ExtendedIsolationForestModelReadWrite.scala:59-67): per tree, nodes in pre-order, ids
 *    from 0, leaves have left = right = -1 and carry num_instances, internal nodes carry
 *    num_instances = -1.  All trees are concatenated; node_off[t] is the first row of tree t.
 */
#ifndef IFB200_H
#define IFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define IFB_API
#else
#define IFB_API __attribute__((visibility("default")))
#endif

#define IFB_ABI_VERSION 1

typedef enum ifb_status {
    IFB_OK = 0,
    IFB_EINVAL = 1,  /* bad argument        -> IllegalArgumentException on the JVM side */
    IFB_ESTATE = 2,  /* bad state           -> IllegalStateException */
    IFB_ECUDA = 3,   /* CUDA runtime error  -> RuntimeException */
    IFB_ENOMEM = 4,  /* host/device OOM     -> RuntimeException */
    IFB_ENOGPU = 5,  /* no usable sm_100 device: the engine has NO CPU fallback */
    IFB_ENCCL = 6    /* NCCL missing or failed (ifb_comm_*, ifb_score_sharded)  -> RuntimeException */
} ifb_status;

typedef enum ifb_layout { IFB_COL_MAJOR = 0, IFB_ROW_MAJOR = 1 } ifb_layout;

typedef struct ifb_forest ifb_forest; /* opaque; owns device memory */

/* ---------------------------------------------------------------------------------------------- */
/* library / device                                                                                */
/* ---------------------------------------------------------------------------------------------- */
IFB_API int ifb_abi_version(void);
IFB_API const char *ifb_last_error(void);
IFB_API int ifb_device_count(int32_t *count);
/* Pinned host memory for staging buffers (JNI: wrap with NewDirectByteBuffer). */
IFB_API int ifb_host_alloc(size_t bytes, void **ptr);
IFB_API int ifb_host_free(void *ptr);
/* Device memory (plain cudaMalloc allocations on purpose: ifb_ipc_export can hand them to peer processes, which memory
 * from a stream-ordered pool does not allow). */
IFB_API int ifb_device_alloc(int32_t device, size_t bytes, void **ptr);
IFB_API int ifb_device_free(int32_t device, void *ptr);
/* Blocking copies between host memory and device memory obtained from ifb_device_alloc. */
IFB_API int ifb_copy_to_device(int32_t device, void *dst_device, const void *src_host, size_t bytes);
IFB_API int ifb_copy_to_host(int32_t device, void *dst_host, const void *src_device, size_t bytes);

/* ---------------------------------------------------------------------------------------------- */
/* forest handles                                                                                  */
/* ---------------------------------------------------------------------------------------------- */

/* Replaces the object graph Array[IsolationTree] of InternalNode/ExternalNode
 * (IF/Nodes.scala:25-66, IF/IsolationTree.scala:18) that IsolationForestModel broadcasts
 * (IF/IsolationForestModel.scala:129).  All arrays are host memory and are copied.
 * num_samples is the model's numSamples (normaliser c(numSamples), IF/IsolationForestModel.scala:128);
 * total_num_features is the training dimension or -1 when unknown (legacy models,
 * IF/IsolationForestModel.scala:52-61). */
IFB_API int ifb_forest_create_standard(int32_t device, int32_t num_trees, const int32_t *node_off /*[T+1]*/,
                                       const int32_t *left, const int32_t *right, const int32_t *feature,
                                       const double *threshold, const int64_t *num_instances,
                                       int32_t num_samples, int32_t total_num_features, ifb_forest **out);

/* Replaces Array[ExtendedIsolationTree] of ExtendedInternalNode(SplitHyperplane)/ExtendedExternalNode
 * (IF/extended/ExtendedNodes.scala:28-63, IF/extended/ExtendedUtils.scala:21-62).  Hyperplanes are CSR:
 * node g (global row) owns hp_idx/hp_w[hp_off[g] .. hp_off[g+1]); leaves own nothing.  The
 * SplitHyperplane invariants (non-empty, ascending distinct non-negative indices,
 * ExtendedUtils.scala:27-34) are validated and reported as IFB_EINVAL. */
IFB_API int ifb_forest_create_extended(int32_t device, int32_t num_trees, const int32_t *node_off,
                                       const int32_t *left, const int32_t *right, const int64_t *num_instances,
                                       const double *offset, const int64_t *hp_off /*[nodes+1]*/,
                                       const int32_t *hp_idx, const float *hp_w, int32_t num_samples,
                                       int32_t total_num_features, ifb_forest **out);

IFB_API int ifb_forest_destroy(ifb_forest *forest);

typedef struct ifb_forest_info {
    int32_t extended;            /* 0 standard, 1 extended */
    int32_t device;
    int32_t num_trees;
    int32_t num_samples;
    int32_t total_num_features;  /* -1 unknown */
    int32_t max_feature_index;   /* largest feature index any node reads */
    int32_t max_depth;           /* deepest leaf over all trees */
    int32_t max_nnz;             /* extended: widest hyperplane; standard: 1 */
    int64_t num_nodes;           /* all trees */
    int64_t num_hp_entries;      /* extended: total CSR entries */
    int64_t device_bytes;        /* bytes of the device-resident node tables */
} ifb_forest_info;
IFB_API int ifb_forest_get_info(const ifb_forest *forest, ifb_forest_info *info);

/* Node tables back to the host in the persisted layout (for MLWriter.save,
 * IF/IsolationForestModelReadWrite.scala:238-249).  Pass NULL for arrays of the other variant. */
IFB_API int ifb_forest_export(const ifb_forest *forest, int32_t *node_off, int32_t *left, int32_t *right,
                              int32_t *feature, double *threshold, int64_t *num_instances, double *offset,
                              int64_t *hp_off, int32_t *hp_idx, float *hp_w);

/* ---------------------------------------------------------------------------------------------- */
/* scoring: IsolationForestModel.transform UDF body (IF/IsolationForestModel.scala:131-139) and      */
/* ExtendedIsolationForestModel.transform (IF/extended/ExtendedIsolationForestModel.scala:114-120):  */
/* per row, f32 left-to-right sum over trees of pathLength (IF/IsolationTree.scala:196-230 /         */
/* IF/extended/ExtendedIsolationTree.scala:283-355), / numTrees, score = 2^(-E/c(numSamples)).       */
/* Optional outputs (may be NULL): depth_sum[r] = sum over trees of the integer leaf depth           */
/* (bit-exact parity handle), path_sum[r] = the f32 sum of path lengths before the division.         */
/* ---------------------------------------------------------------------------------------------- */

/* Device-resident input/output; asynchronous on `stream` (a cudaStream_t, NULL = default stream). */
IFB_API int ifb_score_device(const ifb_forest *forest, const float *X, int64_t n_rows, int32_t d, int64_t ld,
                             int32_t layout, double *scores, int32_t *depth_sum, float *path_sum, void *stream);

/* Host-resident input/output (the call a Spark mapPartitions task makes with one batch): copies the
 * batch in chunks host->device, scores it, copies scores device->host; copies and kernels of successive
 * chunks overlap on internal streams.  Blocks until the scores are in `scores`. */
IFB_API int ifb_score_host(const ifb_forest *forest, const float *X, int64_t n_rows, int32_t d, int64_t ld,
                           int32_t layout, double *scores, int32_t *depth_sum, float *path_sum);

/* Tree-sharded scoring (numEstimators split over GPUs): accumulate this forest shard's contribution into
 * path_sum / depth_sum (device, length n_rows; caller zeroes them), all-reduce them across ranks, then
 * ifb_finalize_scores_device turns the reduced sums into scores with the FULL ensemble size. */
IFB_API int ifb_score_partial_device(const ifb_forest *forest, const float *X, int64_t n_rows, int32_t d,
                                     int64_t ld, int32_t layout, float *path_sum, int32_t *depth_sum,
                                     void *stream);
IFB_API int ifb_finalize_scores_device(int32_t device, const float *path_sum, int64_t n_rows,
                                       int32_t total_num_trees, int32_t num_samples, double *scores,
                                       void *stream);

/* The same layout with the collective INSIDE the library, for hosts that have no collective library of their own (a JVM
 * executor): one communicator per GPU process, bootstrapped from a 128-byte id that one rank creates with
 * ifb_comm_unique_id and the host framework hands to every rank (Spark: a broadcast variable).  ifb_score_sharded scores
 * ALL n_rows rows against `forest` (this rank's slice of the ensemble, e.g. built with ifb_fit_device tree_begin/tree_end,
 * the reference's tree-parallel fit IF/core/SharedTrainLogic.scala:140-149,276-317), sums the per-row path lengths across
 * ranks with ONE NCCL collective on `stream`, and applies the score epilogue with total_num_trees:
 *   IFB_SHARD_ALLREDUCE       every rank receives all scores[0 .. n_rows);
 *   IFB_SHARD_REDUCE_SCATTER  rank r receives only scores[0 .. e-b) of rows [b, e) = [r*per, min(n_rows, (r+1)*per)),
 *                             per = ceil(n_rows / world); the slice is returned in slice_begin / slice_end.
 * NCCL is loaded at run time (libnccl.so.2); without it these calls return IFB_ENCCL. */
typedef struct ifb_comm ifb_comm;
enum { IFB_SHARD_ALLREDUCE = 0, IFB_SHARD_REDUCE_SCATTER = 1 };
IFB_API int ifb_comm_unique_id(void *id128 /* 128 bytes out */);
IFB_API int ifb_comm_init(int32_t device, int32_t world, int32_t rank, const void *id128, ifb_comm **out);
IFB_API int ifb_comm_destroy(ifb_comm *comm);
IFB_API int ifb_score_sharded(const ifb_forest *forest, ifb_comm *comm, const float *X, int64_t n_rows, int32_t d,
                              int64_t ld, int32_t layout, int32_t total_num_trees, int32_t mode, double *scores,
                              int64_t *slice_begin, int64_t *slice_end, void *stream);

/* Fused variant of the tree-sharded layout (no NCCL on the data path): rows are cut into `world` contiguous
 * ownership ranges row_cuts[0..world]; rank r's scoring kernel writes its partial path-length sum of every row
 * straight into the OWNER's buffer peer_partials[owner][r][row - row_cuts[owner]] -- plain stores to NVLink peer
 * memory from the kernel's epilogue, i.e. a reduce-scatter fused into the compute kernel.  peer_partials[o] is
 * rank o's buffer of world * rows_o floats, mapped into this process with ifb_ipc_open (o == rank: the local
 * pointer).  After every rank's kernel has finished (stream sync + barrier), each owner calls
 * ifb_finalize_gathered_device on its own buffer: partials are added in rank order (reproducible) and turned
 * into scores with the full ensemble size.  Standard forests, column-major input. */
IFB_API int ifb_ipc_export(int32_t device, void *device_ptr, void *handle64 /* 64 bytes out */);
IFB_API int ifb_ipc_open(int32_t device, const void *handle64, void **device_ptr);
IFB_API int ifb_ipc_close(int32_t device, void *device_ptr);
IFB_API int ifb_score_scatter_device(const ifb_forest *forest, const float *X, int64_t n_rows, int32_t d, int64_t ld,
                                     int32_t layout, int32_t world, int32_t rank, const int64_t *row_cuts /*[world+1]*/,
                                     float *const *peer_partials /*[world]*/, void *stream);
IFB_API int ifb_finalize_gathered_device(int32_t device, const float *partials, int32_t world, int64_t rows_local,
                                         int32_t total_num_trees, int32_t num_samples, double *scores, void *stream);
/* Device-side barrier for the fused layout (no collective library involved): every rank owns `world` 32-bit
 * flags in peer-visible memory.  ifb_peer_signal_device (enqueued after the scatter kernel) stores `epoch` into
 * slot [rank] of every peer's flag array with system scope; ifb_peer_wait_device (enqueued before the gathered
 * finalize) spins until all `world` local slots hold `epoch`.  Epochs must increase from call to call. */
IFB_API int ifb_peer_signal_device(int32_t device, int32_t world, int32_t rank, uint32_t *const *peer_flags /*[world]*/,
                                   uint32_t epoch, void *stream);
IFB_API int ifb_peer_wait_device(int32_t device, int32_t world, const uint32_t *local_flags, uint32_t epoch,
                                 void *stream);

/* Diagnostics of the tensor-core path of fully-extended forests (extensionLevel = d - 1; csrc/score_ext_tc.cu), which
 * evaluates every hyperplane of the forest (ExtendedUtils.scala:36-55) as one column of a tcgen05 GEMM and uses the
 * accumulators only as a proven filter in front of the reference's exact comparison.
 * ifb_ext_tc_info: padded hyperplane width and number of accumulator columns (0: the forest does not qualify and is
 * scored by the CUDA-core kernels).  ifb_ext_tc_probe: scores the first min(n_rows, 128) rows exactly like
 * ifb_score_device and additionally returns their raw f32 accumulators acc_device[row][n_columns] (device memory) and
 * the weight slot of every column col_slot_host[n_columns] (-1: padding), so that a test can measure the accumulation
 * error of the tensor cores that the bound constant assumes. */
IFB_API int ifb_ext_tc_info(const ifb_forest *forest, int32_t *k_padded, int32_t *n_columns);
IFB_API int ifb_ext_tc_probe(const ifb_forest *forest, const float *X, int64_t n_rows, int32_t d, int64_t ld,
                             int32_t layout, double *scores, float *acc_device, int32_t *col_slot_host, void *stream);

/* Diagnostic of the rank-word path of standard forests (csrc/score_std_rank.cu; opt-in with IFB_STD_RANK=1 in the
 * environment): matrices of d <= 32 features are scored on per-feature ranks (same decisions as
 * IsolationTree.pathLength, IF/IsolationTree.scala:196-230).  *n_chunks = number of forest chunks of that layout for a
 * matrix of d features (built on first use), 0 when the path is off or the forest / shape does not qualify and the f32
 * kernel of score_std.cu scores it. */
IFB_API int ifb_std_rank_info(const ifb_forest *forest, int32_t d, int32_t *n_chunks);

/* prediction column: (score >= threshold) ? 1.0 : 0.0, all 0.0 when threshold <= 0
 * (IF/IsolationForestModel.scala:143-148). */
IFB_API int ifb_predict_device(int32_t device, const double *scores, int64_t n_rows, double threshold,
                               double *labels, void *stream);

/* ---------------------------------------------------------------------------------------------- */
/* fit: IsolationForest.fit / ExtendedIsolationForest.fit tree building                              */
/* (IF/core/SharedTrainLogic.scala:99-153,266-320 sampling + per-tree setup;                          */
/*  IF/IsolationTree.scala:53-183; IF/extended/ExtendedIsolationTree.scala:67-270).                   */
/* One kernel launch builds trees [tree_begin, tree_end) of the ensemble; a tree depends only on       */
/* (random_seed, num_partitions, tree id, data), never on which GPU builds it.                         */
/* ---------------------------------------------------------------------------------------------- */
typedef struct ifb_fit_params {
    int32_t num_estimators;   /* numEstimators (ensemble size) */
    int32_t num_samples;      /* resolved numSamples  (SharedTrainLogic.scala:47-75)  */
    int32_t num_features;     /* resolved numFeatures (SharedTrainLogic.scala:33-45)  */
    int32_t bootstrap;        /* 0/1 */
    int64_t random_seed;      /* randomSeed param */
    int32_t num_partitions;   /* P in treeSeed = randomSeed + 2*(P+1) + treeId (IF/IsolationForest.scala:76-78) */
    int32_t extension_level;  /* -1 = standard IF; >= 0 = extended IF with min(level+1, numFeatures) non-zeros */
    int32_t tree_begin;       /* shard: build trees [tree_begin, tree_end) */
    int32_t tree_end;
} ifb_fit_params;

IFB_API int ifb_fit_device(int32_t device, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                           const ifb_fit_params *params, ifb_forest **out, void *stream);
IFB_API int ifb_fit_host(int32_t device, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                         const ifb_fit_params *params, ifb_forest **out);

/* ---------------------------------------------------------------------------------------------- */
/* threshold: exact order statistic of the scores, the value approxQuantile(scoreCol, [q], 0.0)      */
/* returns (IF/core/SharedTrainLogic.scala:191-198): element of 1-based rank ceil(q*n) of the sorted  */
/* scores.  Also returns the observed contamination  #(score >= value)/n  (:211-213).                */
/* ---------------------------------------------------------------------------------------------- */
/* The result is returned in HOST memory, so the call enqueues its 8 radix passes on `stream` and then blocks until they
 * have finished (the reference's approxQuantile is an action as well). */
IFB_API int ifb_quantile_device(int32_t device, const double *scores, int64_t n_rows, double q, double *value,
                                double *observed_fraction_ge, void *stream);

/* Scalar helper: Utils.avgPathLength (IF/core/Utils.scala:85-92), evaluated on the host exactly as the
 * forest tables are built. */
IFB_API float ifb_avg_path_length(int64_t num_instances);

/* Counters for bench.py: kernels launched by this library since load / since the last reset. */
IFB_API int64_t ifb_kernel_launch_count(int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* IFB200_H */
