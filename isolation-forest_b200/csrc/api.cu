// C-ABI entry points for scoring / threshold (include/ifb200.h).  Host-side orchestration only: argument
// checks with the reference's messages, staging transposes, the host<->device copy pipeline.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "ifb_internal.h"

namespace ifb {

int launch_transpose(const float *in, int64_t n, int32_t d, int64_t ld_in, float *out, int64_t ld_out,
                     cudaStream_t stream);
int launch_select(const double *scores, int64_t n, int64_t rank0, double *value, unsigned long long *count_ge,
                  cudaStream_t stream);

// keep stream-ordered allocations cached instead of returning them to the OS after every call
void tune_mempool(int device) {
    static std::mutex mu;
    static std::vector<int> done;
    std::lock_guard<std::mutex> lk(mu);
    if (std::find(done.begin(), done.end(), device) != done.end()) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done.push_back(device);
}

namespace {

// The require()s at the top of transform (IF/IsolationForestModel.scala:118-125,
// IF/extended/ExtendedIsolationForestModel.scala:100-107) and the per-row dimension check
// (IF/core/Utils.scala:67-72), with the reference's message texts.
int check_scoring_args(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout) {
    IFB_REQUIRE(f, "forest is null");
    IFB_REQUIRE(f->num_samples >= 2, "Cannot score with numSamples=%d; expected numSamples >= 2.", f->num_samples);
    IFB_REQUIRE(f->num_trees > 0, f->extended ? "Cannot score with an empty ExtendedIsolationForestModel."
                                              : "Cannot score with an empty IsolationForestModel.");
    IFB_REQUIRE(n_rows >= 0, "n_rows must be >= 0");
    IFB_REQUIRE(d >= 1, "d must be >= 1");
    IFB_REQUIRE(layout == IFB_COL_MAJOR || layout == IFB_ROW_MAJOR, "unknown layout %d", layout);
    IFB_REQUIRE(n_rows == 0 || X, "X is null");
    IFB_REQUIRE(layout == IFB_COL_MAJOR ? ld >= n_rows : ld >= d, "leading dimension %lld too small", (long long)ld);
    if (f->total_num_features != -1)
        IFB_REQUIRE(d == f->total_num_features,
                    "Input feature vector size %d did not match the model's training dimension %d.", d,
                    f->total_num_features);
    IFB_REQUIRE(f->max_feature_index < d, "Input feature vector size %d is smaller than the largest feature index %d "
                "the model reads.", d, f->max_feature_index);
    return IFB_OK;
}

int score_device_impl(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                      double *scores, int32_t *depth_sum, float *path_sum, bool accumulate_only,
                      cudaStream_t stream) {
    if (n_rows == 0) return IFB_OK;
    tune_mempool(f->device);
    if (f->extended)
        return launch_score_extended(f, X, n_rows, d, ld, layout, scores, depth_sum, path_sum, accumulate_only,
                                     stream);
    ifb_forest::StdPlan *plan = nullptr;
    int rc = get_std_plan(const_cast<ifb_forest *>(f), d, &plan);
    if (rc) return rc;
    if (!plan)  // no shared-memory plan for rows this wide: generic kernel, either layout
        return launch_score_standard_generic(f, X, n_rows, d, ld, layout, scores, depth_sum, path_sum, accumulate_only, stream);
    struct Scratch {   // stream-ordered scratch, released on every exit path
        cudaStream_t s;
        float *p = nullptr;
        ~Scratch() {
            if (p) cudaFreeAsync(p, s);
        }
    } xt_g{stream}, sum_g{stream};
    float *&xt = xt_g.p;
    float *&tmp_sum = sum_g.p;
    const float *Xc = X;
    int64_t ldc = ld;
    if (layout == IFB_ROW_MAJOR) {
        ldc = (n_rows + 3) & ~3LL;
        IFB_CUDA(cudaMallocAsync((void **)&xt, (size_t)ldc * d * 4, stream));
        rc = launch_transpose(X, n_rows, d, ld, xt, ldc, stream);
        if (rc) return rc;
        Xc = xt;
    }
    if (plan->chunks.size() > 1 && !path_sum && !accumulate_only) {
        IFB_CUDA(cudaMallocAsync((void **)&tmp_sum, (size_t)n_rows * 4, stream));
        path_sum = tmp_sum;
    }
    return launch_score_standard(f, plan, Xc, n_rows, d, ldc, IFB_COL_MAJOR, scores, depth_sum, path_sum, accumulate_only,
                                 stream);
}

}  // namespace
}  // namespace ifb

using namespace ifb;

extern "C" {

int ifb_score_device(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                     double *scores, int32_t *depth_sum, float *path_sum, void *stream) {
    int rc = check_scoring_args(f, X, n_rows, d, ld, layout);
    if (rc) return rc;
    IFB_REQUIRE(n_rows == 0 || scores, "scores is null");
    DeviceGuard dg(f->device);
    return score_device_impl(f, X, n_rows, d, ld, layout, scores, depth_sum, path_sum, false, (cudaStream_t)stream);
}

int ifb_score_partial_device(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld,
                             int32_t layout, float *path_sum, int32_t *depth_sum, void *stream) {
    int rc = check_scoring_args(f, X, n_rows, d, ld, layout);
    if (rc) return rc;
    IFB_REQUIRE(n_rows == 0 || path_sum, "path_sum is null");
    DeviceGuard dg(f->device);
    return score_device_impl(f, X, n_rows, d, ld, layout, nullptr, depth_sum, path_sum, true, (cudaStream_t)stream);
}

int ifb_finalize_scores_device(int32_t device, const float *path_sum, int64_t n_rows, int32_t total_num_trees,
                               int32_t num_samples, double *scores, void *stream) {
    IFB_REQUIRE(num_samples >= 2, "Cannot score with numSamples=%d; expected numSamples >= 2.", num_samples);
    IFB_REQUIRE(total_num_trees > 0, "Cannot score with an empty IsolationForestModel.");
    IFB_REQUIRE(n_rows == 0 || (path_sum && scores), "null buffer");
    DeviceGuard dg(device);
    return launch_finalize(path_sum, n_rows, total_num_trees, avg_path_length_host(num_samples), scores,
                           (cudaStream_t)stream);
}

// ---- fused tree-sharded scoring over peer memory ------------------------------------------------------
int ifb_ipc_export(int32_t device, void *device_ptr, void *handle64) {
    IFB_REQUIRE(device_ptr && handle64, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard dg(device);
    IFB_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), device_ptr));
    return IFB_OK;
}
int ifb_ipc_open(int32_t device, const void *handle64, void **device_ptr) {
    IFB_REQUIRE(device_ptr && handle64, "null argument");
    DeviceGuard dg(device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    IFB_CUDA(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return IFB_OK;
}
int ifb_ipc_close(int32_t device, void *device_ptr) {
    DeviceGuard dg(device);
    if (device_ptr) IFB_CUDA(cudaIpcCloseMemHandle(device_ptr));
    return IFB_OK;
}

int ifb_score_scatter_device(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                             int32_t world, int32_t rank, const int64_t *row_cuts, float *const *peer_partials,
                             void *stream) {
    int rc = check_scoring_args(f, X, n_rows, d, ld, layout);
    if (rc) return rc;
    IFB_REQUIRE(!f->extended, "ifb_score_scatter_device supports standard forests (use ifb_score_partial_device + "
                              "all-reduce for extended forests)");
    IFB_REQUIRE(layout == IFB_COL_MAJOR, "ifb_score_scatter_device expects a column-major matrix");
    IFB_REQUIRE(world >= 1 && world <= kMaxScatterRanks && rank >= 0 && rank < world, "bad world/rank %d/%d", world, rank);
    IFB_REQUIRE(row_cuts && peer_partials, "null argument");
    IFB_REQUIRE(row_cuts[0] == 0 && row_cuts[world] == n_rows, "row_cuts must cover [0, n_rows)");
    ScatterTarget st;
    st.world = world;
    st.rank = rank;
    for (int i = 0; i <= kMaxScatterRanks; i++) st.cut[i] = i <= world ? row_cuts[i] : n_rows;
    for (int i = 0; i < world; i++) IFB_REQUIRE(st.cut[i] <= st.cut[i + 1], "row_cuts must be non-decreasing");
    for (int i = 0; i < kMaxScatterRanks; i++) st.peer[i] = i < world ? peer_partials[i] : nullptr;
    if (n_rows == 0) return IFB_OK;
    DeviceGuard dg(f->device);
    tune_mempool(f->device);
    ifb_forest::StdPlan *plan = nullptr;
    rc = get_std_plan(const_cast<ifb_forest *>(f), d, &plan);
    if (rc) return rc;
    IFB_REQUIRE(plan != nullptr, "ifb_score_scatter_device: rows of %d features are too wide for the fused kernel", d);
    float *tmp_sum = nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    if (plan->chunks.size() > 1) IFB_CUDA(cudaMallocAsync((void **)&tmp_sum, (size_t)n_rows * 4, s));
    // all chunks but the last accumulate locally; the last one scatters the finished sums
    rc = launch_score_standard(f, plan, X, n_rows, d, ld, IFB_COL_MAJOR, nullptr, nullptr, tmp_sum, /*accumulate_only=*/false,
                               s, &st);
    if (tmp_sum) cudaFreeAsync(tmp_sum, s);
    return rc;
}

int ifb_finalize_gathered_device(int32_t device, const float *partials, int32_t world, int64_t rows_local,
                                 int32_t total_num_trees, int32_t num_samples, double *scores, void *stream) {
    IFB_REQUIRE(num_samples >= 2, "Cannot score with numSamples=%d; expected numSamples >= 2.", num_samples);
    IFB_REQUIRE(total_num_trees > 0 && world >= 1, "bad ensemble size / world");
    IFB_REQUIRE(rows_local == 0 || (partials && scores), "null buffer");
    DeviceGuard dg(device);
    return launch_finalize_gathered(partials, world, rows_local, total_num_trees, avg_path_length_host(num_samples), scores,
                                    (cudaStream_t)stream);
}

int ifb_peer_signal_device(int32_t device, int32_t world, int32_t rank, uint32_t *const *peer_flags, uint32_t epoch,
                           void *stream) {
    IFB_REQUIRE(world >= 1 && world <= kMaxScatterRanks && rank >= 0 && rank < world && peer_flags, "bad arguments");
    DeviceGuard dg(device);
    return launch_peer_signal(world, rank, peer_flags, epoch, (cudaStream_t)stream);
}
int ifb_peer_wait_device(int32_t device, int32_t world, const uint32_t *local_flags, uint32_t epoch, void *stream) {
    IFB_REQUIRE(world >= 1 && world <= kMaxScatterRanks && local_flags, "bad arguments");
    DeviceGuard dg(device);
    return launch_peer_wait(world, local_flags, epoch, (cudaStream_t)stream);
}

int ifb_ext_tc_info(const ifb_forest *f, int32_t *k_padded, int32_t *n_columns) {
    IFB_REQUIRE(f && k_padded && n_columns, "null argument");
    *k_padded = f->tc_ok ? f->tc_kp : 0;
    *n_columns = f->tc_ok ? f->tc_blocks * 256 : 0;
    return IFB_OK;
}

int ifb_std_rank_info(const ifb_forest *f, int32_t d, int32_t *n_chunks) {
    IFB_REQUIRE(f && n_chunks, "null argument");
    *n_chunks = 0;
    if (f->extended || d < 1 || d > 32 || f->num_trees == 0 || !std_rank_enabled()) return IFB_OK;
    DeviceGuard dg(f->device);
    RankPlan *rp = nullptr;
    int rc = get_rank_plan(const_cast<ifb_forest *>(f), d, &rp);
    if (rc) return rc;
    *n_chunks = rank_plan_chunks(rp);
    return IFB_OK;
}

int ifb_ext_tc_probe(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                     double *scores, float *acc_device, int32_t *col_slot_host, void *stream) {
    int rc = check_scoring_args(f, X, n_rows, d, ld, layout);
    if (rc) return rc;
    IFB_REQUIRE(f->extended && f->tc_ok, "the forest has no tensor-core layout");
    IFB_REQUIRE(scores && acc_device && col_slot_host && n_rows >= 1, "null argument");
    DeviceGuard dg(f->device);
    const int64_t rows = std::min<int64_t>(n_rows, 128);
    rc = launch_score_extended_tc(f, X, rows, d, ld, layout, scores, nullptr, nullptr, false, (cudaStream_t)stream, acc_device);
    IFB_REQUIRE(rc >= 0, "the tensor-core path refused this call (d = %d, forest width %d)", d, f->tc_k);
    if (rc) return rc;
    IFB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    IFB_CUDA(cudaMemcpy(col_slot_host, f->d_tc_col_slot, (size_t)f->tc_blocks * 256 * 4, cudaMemcpyDeviceToHost));
    return IFB_OK;
}

int ifb_predict_device(int32_t device, const double *scores, int64_t n_rows, double threshold, double *labels,
                       void *stream) {
    IFB_REQUIRE(n_rows == 0 || (scores && labels), "null buffer");
    DeviceGuard dg(device);
    return launch_predict(scores, n_rows, threshold, labels, (cudaStream_t)stream);
}

int ifb_quantile_device(int32_t device, const double *scores, int64_t n_rows, double q, double *value,
                        double *observed_fraction_ge, void *stream) {
    IFB_REQUIRE(scores && value, "null buffer");
    IFB_REQUIRE(n_rows >= 1, "quantile of an empty score vector");
    IFB_REQUIRE(q >= 0.0 && q <= 1.0, "quantile %g outside [0,1]", q);
    DeviceGuard dg(device);
    tune_mempool(device);
    // QuantileSummaries.query with relativeError 0: 1-based rank ceil(q*n), clamped to [1, n]
    int64_t rank1 = (int64_t)std::ceil(q * (double)n_rows);
    rank1 = std::min<int64_t>(std::max<int64_t>(rank1, 1), n_rows);
    unsigned long long cge = 0;
    int rc = launch_select(scores, n_rows, rank1 - 1, value, &cge, (cudaStream_t)stream);
    if (rc) return rc;
    if (observed_fraction_ge) *observed_fraction_ge = (double)cge / (double)n_rows;
    return IFB_OK;
}

// Host-buffer scoring: rows are cut into chunks; chunk i+1's H2D copy overlaps chunk i's kernel and chunk
// i-1's D2H copy on three round-robin streams.
int ifb_score_host(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                   double *scores, int32_t *depth_sum, float *path_sum) {
    int rc = check_scoring_args(f, X, n_rows, d, ld, layout);
    if (rc) return rc;
    IFB_REQUIRE(n_rows == 0 || scores, "scores is null");
    if (n_rows == 0) return IFB_OK;
    DeviceGuard dg(f->device);
    tune_mempool(f->device);
    constexpr int kSlots = 3;
    int64_t chunk = (int64_t)(96ll << 20) / (4ll * d);
    chunk = std::max<int64_t>(chunk, 1 << 16);
    chunk = (chunk + 1023) & ~1023ll;
    chunk = std::min<int64_t>(chunk, (n_rows + 3) & ~3ll);
    const int64_t n_chunks = (n_rows + chunk - 1) / chunk;
    const int slots = (int)std::min<int64_t>(kSlots, n_chunks);
    // Streams are cached per (host thread, device): a Spark task thread scores batch after batch, and creating and
    // destroying three streams per call was visible in small batches (config 1: 1,000 rows).  The stream-ordered
    // scratch comes from the device's memory pool (release threshold raised in tune_mempool) and is returned on every
    // exit path.  The streams of a thread live until the process ends.
    struct StreamSet {
        cudaStream_t st[kSlots] = {};
        bool ok = false;
    };
    thread_local StreamSet tl_streams[16];
    StreamSet *cached = (f->device >= 0 && f->device < 16) ? &tl_streams[f->device] : nullptr;
    StreamSet own;
    StreamSet *ss = cached ? cached : &own;
    if (!ss->ok) {
        for (int i = 0; i < kSlots; i++) IFB_CUDA(cudaStreamCreateWithFlags(&ss->st[i], cudaStreamNonBlocking));
        ss->ok = true;
    }
    struct Pipe {
        cudaStream_t st[kSlots] = {};
        bool owned = false;
        int slots = 0;
        float *dX[kSlots] = {};
        double *dS[kSlots] = {};
        int32_t *dD[kSlots] = {};
        float *dP[kSlots] = {};
        ~Pipe() {
            for (int i = 0; i < slots; i++) {
                if (dX[i]) cudaFreeAsync(dX[i], st[i]);
                if (dS[i]) cudaFreeAsync(dS[i], st[i]);
                if (dD[i]) cudaFreeAsync(dD[i], st[i]);
                if (dP[i]) cudaFreeAsync(dP[i], st[i]);
                cudaStreamSynchronize(st[i]);
            }
            if (owned)
                for (int i = 0; i < kSlots; i++) cudaStreamDestroy(st[i]);
        }
    } pipe;
    pipe.owned = cached == nullptr;
    pipe.slots = slots;
    for (int i = 0; i < kSlots; i++) pipe.st[i] = ss->st[i];
    cudaStream_t *st = pipe.st;
    float **dX = pipe.dX;
    double **dS = pipe.dS;
    int32_t **dD = pipe.dD;
    float **dP = pipe.dP;
    for (int i = 0; i < slots; i++) {
        IFB_CUDA(cudaMallocAsync((void **)&dX[i], (size_t)chunk * d * 4, st[i]));
        IFB_CUDA(cudaMallocAsync((void **)&dS[i], (size_t)chunk * 8, st[i]));
        if (depth_sum) IFB_CUDA(cudaMallocAsync((void **)&dD[i], (size_t)chunk * 4, st[i]));
        if (path_sum) IFB_CUDA(cudaMallocAsync((void **)&dP[i], (size_t)chunk * 4, st[i]));
    }
    for (int64_t c = 0; c < n_chunks && rc == IFB_OK; c++) {
        const int s = (int)(c % slots);
        const int64_t r0 = c * chunk;
        const int64_t rows = std::min<int64_t>(chunk, n_rows - r0);
        cudaError_t e;
        int64_t ldd;
        if (layout == IFB_COL_MAJOR) {
            ldd = chunk;
            e = cudaMemcpy2DAsync(dX[s], (size_t)ldd * 4, X + r0, (size_t)ld * 4, (size_t)rows * 4, (size_t)d,
                                  cudaMemcpyHostToDevice, st[s]);
        } else {
            ldd = d;
            if (ld == d)
                e = cudaMemcpyAsync(dX[s], X + r0 * ld, (size_t)rows * d * 4, cudaMemcpyHostToDevice, st[s]);
            else
                e = cudaMemcpy2DAsync(dX[s], (size_t)d * 4, X + r0 * ld, (size_t)ld * 4, (size_t)d * 4, (size_t)rows,
                                      cudaMemcpyHostToDevice, st[s]);
        }
        if (e != cudaSuccess) {
            set_error("host->device copy failed: %s", cudaGetErrorString(e));
            rc = IFB_ECUDA;
            break;
        }
        rc = score_device_impl(f, dX[s], rows, d, ldd, layout, dS[s], dD[s], dP[s], false, st[s]);
        if (rc) break;
        e = cudaMemcpyAsync(scores + r0, dS[s], (size_t)rows * 8, cudaMemcpyDeviceToHost, st[s]);
        if (e == cudaSuccess && depth_sum)
            e = cudaMemcpyAsync(depth_sum + r0, dD[s], (size_t)rows * 4, cudaMemcpyDeviceToHost, st[s]);
        if (e == cudaSuccess && path_sum)
            e = cudaMemcpyAsync(path_sum + r0, dP[s], (size_t)rows * 4, cudaMemcpyDeviceToHost, st[s]);
        if (e != cudaSuccess) {
            set_error("device->host copy failed: %s", cudaGetErrorString(e));
            rc = IFB_ECUDA;
        }
    }
    for (int i = 0; i < slots; i++) {
        cudaError_t e = cudaStreamSynchronize(st[i]);
        if (e != cudaSuccess && rc == IFB_OK) {
            set_error("scoring pipeline failed: %s", cudaGetErrorString(e));
            rc = IFB_ECUDA;
        }
    }
    return rc;
}

}  // extern "C"
