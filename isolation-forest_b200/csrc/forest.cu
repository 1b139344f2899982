// Forest handles: validation of the persisted node tables, relayout into the kernel layout, upload.
//
// Replaces the broadcast object graph of the reference (IF/Nodes.scala:25-66,
// IF/extended/ExtendedNodes.scala:28-63; broadcast at IF/IsolationForestModel.scala:129).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <queue>

#include "ifb_internal.h"

namespace ifb {

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

static std::atomic_llong g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Utils.avgPathLength, IF/core/Utils.scala:85-92: f32 arithmetic around an f64 log rounded to f32.
static float avg_path_length_compute(int64_t n) {
    if (n <= 1) return 0.0f;
    const float euler = 0.5772156649f;
    volatile float nf = (float)n;
    volatile float lg = (float)std::log((double)(nf - 1.0f));
    volatile float a = 2.0f * (lg + euler);
    volatile float b = (2.0f * (nf - 1.0f)) / nf;
    return a - b;
}
// Leaf sizes are bounded by maxSamples, so the same few hundred values are asked for once per leaf of every forest:
// a table for n < 4096 (filled once) takes the log() out of forest creation.
float avg_path_length_host(int64_t n) {
    constexpr int kTable = 4096;
    static std::once_flag once;
    static float table[kTable];
    if (n < 0 || n >= kTable) return avg_path_length_compute(n);
    std::call_once(once, [] {
        for (int i = 0; i < kTable; i++) table[i] = avg_path_length_compute(i);
    });
    return table[n];
}

int device_smem_optin(int device) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    return v;
}
int device_sm_count(int device) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    return v;
}

// Smallest f32 >= t.  For an f32 x:  (double)x < t  <=>  x < ceil32(t)   (proof in DESIGN.md).
static float ceil_to_f32(double t) {
    float f = (float)t;  // round to nearest
    if ((double)f < t) f = std::nextafterf(f, INFINITY);
    return f;
}

// Shape validation shared by both variants.  Mirrors what the reference's loader enforces
// (IF/IsolationForestModelReadWrite.scala:179-205: ids 0..n-1 in pre-order, children resolvable).
static int validate_shape(int32_t T, const int32_t *node_off, const int32_t *left, const int32_t *right,
                          const int64_t *num_instances, bool allow_empty_leaf) {
    IFB_REQUIRE(T >= 0, "num_trees must be >= 0, got %d", T);
    IFB_REQUIRE(node_off && (T == 0 || (left && right && num_instances)), "null node table");
    IFB_REQUIRE(node_off[0] == 0, "node_off[0] must be 0");
    for (int t = 0; t < T; t++) {
        int32_t n = node_off[t + 1] - node_off[t];
        IFB_REQUIRE(n >= 1, "tree %d has no nodes", t);
        const int32_t base = node_off[t];
        std::vector<uint8_t> seen(n, 0);
        seen[0] = 1;
        for (int32_t i = 0; i < n; i++) {
            int32_t l = left[base + i], r = right[base + i];
            if (l == -1 || r == -1) {
                IFB_REQUIRE(l == -1 && r == -1, "tree %d node %d: half-leaf (left=%d right=%d)", t, i, l, r);
                int64_t c = num_instances[base + i];
                IFB_REQUIRE(allow_empty_leaf ? c >= 0 : c > 0,
                            "tree %d node %d: leaf numInstances %lld is invalid", t, i, (long long)c);
            } else {
                // pre-order: left child is the next row, right child comes later
                IFB_REQUIRE(l == i + 1 && r > l && r < n, "tree %d node %d: children (%d,%d) not in pre-order", t, i,
                            l, r);
                IFB_REQUIRE(!seen[l] && !seen[r], "tree %d node %d: child visited twice", t, i);
                seen[l] = seen[r] = 1;
            }
        }
        for (int32_t i = 0; i < n; i++) IFB_REQUIRE(seen[i], "tree %d node %d is unreachable", t, i);
    }
    return IFB_OK;
}

// BFS order of one tree: order[k] = pre-order id of the k-th BFS node, children of a node adjacent.
static void bfs_order(const int32_t *left, const int32_t *right, int32_t n, std::vector<int32_t> &order,
                      std::vector<int32_t> &depth_of_bfs) {
    order.clear();
    depth_of_bfs.clear();
    order.reserve(n);
    order.push_back(0);
    depth_of_bfs.push_back(0);
    for (size_t h = 0; h < order.size(); h++) {
        int32_t id = order[h];
        if (left[id] != -1) {
            order.push_back(left[id]);
            depth_of_bfs.push_back(depth_of_bfs[h] + 1);
            order.push_back(right[id]);
            depth_of_bfs.push_back(depth_of_bfs[h] + 1);
        }
    }
}

int build_standard_tables(ifb_forest *f) {
    const int T = f->num_trees;
    const int64_t total = f->node_off[T];
    f->h_val.assign(total, 0.f);
    f->h_meta_feat.assign(total, 0xFFFFFFFFu);
    f->h_child.assign(total, -1);
    f->h_depth.assign(total, 0);
    f->bfs_off.assign(f->node_off.begin(), f->node_off.end());
    f->max_depth = 0;
    f->max_feature_index = -1;
    std::vector<int32_t> order, dep, pos;
    for (int t = 0; t < T; t++) {
        const int32_t base = f->node_off[t], n = f->node_off[t + 1] - base;
        bfs_order(&f->left[base], &f->right[base], n, order, dep);
        pos.assign(n, 0);
        for (int32_t k = 0; k < n; k++) pos[order[k]] = k;
        for (int32_t k = 0; k < n; k++) {
            const int32_t id = order[k];
            const int64_t g = base + k;
            IFB_REQUIRE(dep[k] <= 250, "tree %d deeper than 250 levels", t);
            f->h_depth[g] = (uint8_t)dep[k];
            if (f->left[base + id] == -1) {
                // leaf value = currentPathLength (exact small integer in f32) + avgPathLength(n):
                // one f32 add, exactly as IF/IsolationTree.scala:218.
                volatile float v = (float)dep[k] + avg_path_length_host(f->num_instances[base + id]);
                f->h_val[g] = v;
                f->max_depth = std::max(f->max_depth, dep[k]);
            } else {
                f->h_val[g] = ceil_to_f32(f->threshold[base + id]);
                f->h_meta_feat[g] = (uint32_t)f->feature[base + id];
                f->h_child[g] = pos[f->left[base + id]];
                f->max_feature_index = std::max(f->max_feature_index, f->feature[base + id]);
            }
        }
    }
    return IFB_OK;
}

// ---- device-resident hyperplanes (forests fitted on the device with k == d) ---------------------------------------
namespace {
// slot s of the scoring layout <- row slot_src[s] of the builder's weight array (one warp per slot); also the slot's
// sum |w| and ||w||_2 (f64, the bounds of the exact / f32 tiers) and the "weights are f32-tier safe" flag
__global__ void ext_gather_weights_kernel(const float *__restrict__ src, const int64_t *__restrict__ slot_src, int64_t n_slots,
                                          int k, float *__restrict__ w, double *__restrict__ wabs, double *__restrict__ wnorm,
                                          int32_t *__restrict__ unsafe) {
    const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n_slots) return;
    const float *in = src + slot_src[s] * k;
    float *out = w + s * k;
    double a = 0.0, q = 0.0;
    bool bad = false;
    for (int i = lane; i < k; i += 32) {
        const float v = in[i];
        out[i] = v;
        a += fabs((double)v);
        q += (double)v * (double)v;
        bad = bad || !(fabsf(v) <= 0x1p40f);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        // lane-parallel f64 sums differ from a sequential sum by <= k 2^-53 relative: covered by the inflations
        wabs[s] = a * 1.0000001;
        wnorm[s] = sqrt(q) * 1.0000002;
        if (bad) atomicExch(unsafe, 1);
    }
}
__global__ void ext_patch_wide_nodes_kernel(WideNode *nodes, int64_t n_nodes, const double *__restrict__ wnorm) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_nodes) return;
    const int32_t slot = nodes[g].slot;
    if (slot < 0) return;
    const double v = wnorm[slot];
    float wf = (float)v;
    if ((double)wf < v) wf = nextafterf(wf, INFINITY);
    nodes[g].wnorm = wf;
}
}  // namespace

int build_extended_tables(ifb_forest *f, const DeviceHyperplanes *dev) {
    const bool timing = getenv("IFB_FIT_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (timing)
            std::fprintf(stderr, "[ifb ext tables] %s at %.3f ms\n", what,
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    const int T = f->num_trees;
    const int64_t total = f->node_off[T];
    const int k = f->max_nnz;
    std::vector<double> off(total, 0.0);
    std::vector<float> leaf(total, 0.f);
    std::vector<int32_t> child(total, -1), hp(total, -1), len(total, 0);
    std::vector<uint8_t> depthv(total, 0);   // BFS-order depth of every node (leaf depth = internal nodes on its path)
    std::vector<int64_t> tree_node(f->node_off.begin(), f->node_off.end());
    int64_t internal = 0;
    for (int64_t g = 0; g < total; g++) internal += (f->left[g] != -1);
    // weights that already live on the device are gathered there (dev != nullptr): no host copy of w / idx
    std::vector<float> w(dev ? 0 : (size_t)internal * k, 0.f);
    std::vector<int32_t> idx(dev ? 0 : (size_t)internal * k, 0);
    std::vector<int64_t> slot_src(dev ? (size_t)internal : 0, 0);
    if (dev) f->lazy_slot_of_node.assign((size_t)total, -1);
    bool identity = true;
    f->max_depth = 0;
    std::vector<int32_t> order, dep, pos;
    int64_t slot = 0;
    for (int t = 0; t < T; t++) {
        const int32_t base = f->node_off[t], n = f->node_off[t + 1] - base;
        bfs_order(&f->left[base], &f->right[base], n, order, dep);
        pos.assign(n, 0);
        for (int32_t q = 0; q < n; q++) pos[order[q]] = q;
        for (int32_t q = 0; q < n; q++) {
            const int32_t id = order[q];
            const int64_t g = base + q, src = base + id;
            depthv[g] = (uint8_t)std::min(dep[q], 255);
            if (f->left[src] == -1) {
                volatile float v = (float)dep[q] + avg_path_length_host(f->num_instances[src]);
                leaf[g] = v;
                f->max_depth = std::max(f->max_depth, dep[q]);
            } else {
                off[g] = f->offset[src];
                child[g] = pos[f->left[src]];
                hp[g] = (int32_t)slot;
                if (dev) {
                    len[g] = k;
                    slot_src[(size_t)slot] = dev->src_row[src];
                    f->lazy_slot_of_node[(size_t)src] = (int32_t)slot;
                    slot++;
                    continue;
                }
                const int64_t b = f->hp_off[src], e = f->hp_off[src + 1];
                len[g] = (int32_t)(e - b);
                // Rows narrower than k are padded (index of the last term, weight 0); kernels stop at len.
                for (int64_t i = 0; i < k; i++) {
                    const int64_t s = b + std::min<int64_t>(i, e - b - 1);
                    idx[slot * k + i] = f->hp_idx[s];
                    w[slot * k + i] = (i < e - b) ? f->hp_w[s] : 0.0f;
                    if (idx[slot * k + i] != i) identity = false;
                }
                if (e - b != k) identity = false;
                slot++;
            }
        }
    }
    f->ext_internal_slots = internal;
    lap("BFS relayout, weights / indices");
    f->ext_dense_identity = identity && internal > 0;
    DeviceGuard dg(f->device);
    // Every table goes into ONE device allocation (a dozen cudaMallocs cost more than the uploads): `up` records
    // the request, `commit` allocates the arena, copies each piece and hands out the 256-byte aligned pointers.
    struct Piece {
        void **dptr;
        const void *src;
        size_t bytes, at;
    };
    std::vector<Piece> pieces;
    size_t arena_bytes = 0;
    auto up = [&](void **dptr, const void *src, size_t bytes) -> int {
        if (bytes == 0) bytes = 16;
        pieces.push_back(Piece{dptr, src, bytes, arena_bytes});
        arena_bytes += (bytes + 255) & ~(size_t)255;
        return IFB_OK;
    };
    auto commit = [&]() -> int {
        IFB_CUDA(cudaMalloc((void **)&f->d_ext_arena, arena_bytes));
        f->device_bytes += (int64_t)arena_bytes;
        for (const Piece &pc : pieces) {
            *pc.dptr = f->d_ext_arena + pc.at;
            if (pc.src) IFB_CUDA(cudaMemcpyAsync(*pc.dptr, pc.src, pc.bytes, cudaMemcpyHostToDevice, 0));
        }
        IFB_CUDA(cudaStreamSynchronize(0));
        return IFB_OK;
    };
    // host tables referenced by `pieces` must outlive commit()
    std::vector<double> wabs((size_t)internal, 0.0), wnorm((size_t)internal, 0.0);
    std::vector<WideNode> wn((size_t)total);
    std::vector<int32_t> tslot((size_t)std::max(T, 1), -1);
    int rc;
    if ((rc = up((void **)&f->d_ext_w, dev ? nullptr : w.data(), (size_t)internal * k * 4))) return rc;
    if (!f->ext_dense_identity)
        if ((rc = up((void **)&f->d_ext_idx, idx.data(), idx.size() * 4))) return rc;
    if ((rc = up((void **)&f->d_ext_off, off.data(), off.size() * 8))) return rc;
    if ((rc = up((void **)&f->d_ext_leaf, leaf.data(), leaf.size() * 4))) return rc;
    if ((rc = up((void **)&f->d_ext_child, child.data(), child.size() * 4))) return rc;
    if ((rc = up((void **)&f->d_ext_hp, hp.data(), hp.size() * 4))) return rc;
    if ((rc = up((void **)&f->d_ext_len, len.data(), len.size() * 4))) return rc;
    {
        for (int64_t sl = 0; sl < (dev ? 0 : internal); sl++) {
            double a = 0.0, q = 0.0;
            for (int i = 0; i < k; i++) {
                const double v = (double)w[(size_t)sl * k + i];
                a += std::fabs(v);
                q += v * v;
            }
            wabs[(size_t)sl] = a;
            wnorm[(size_t)sl] = std::sqrt(q) * 1.0000001;
        }
        if ((rc = up((void **)&f->d_ext_wabs, dev ? nullptr : wabs.data(), wabs.size() * 8))) return rc;
        if ((rc = up((void **)&f->d_ext_wnorm, dev ? nullptr : wnorm.data(), wnorm.size() * 8))) return rc;
        for (int t = 0; t < T; t++) {
            const int64_t base = f->node_off[t];
            const int n = f->node_off[t + 1] - f->node_off[t];
            tslot[t] = hp[base];
            for (int q = 0; q < n; q++) {
                const int64_t g = base + q;
                WideNode &r = wn[(size_t)g];
                r.off = off[g];
                r.leaf = leaf[g];
                r.cbase = child[g];
                r.slot = hp[g];
                r.slot_l = r.slot_r = -1;
                r.wnorm = 0.f;
                if (child[g] >= 0) {
                    r.slot_l = hp[base + child[g]];
                    r.slot_r = hp[base + child[g] + 1];
                    if (dev) continue;   // wnorm is patched on the device once the weights have been gathered
                    float wf = (float)wnorm[(size_t)hp[g]];
                    if ((double)wf < wnorm[(size_t)hp[g]]) wf = std::nextafterf(wf, INFINITY);
                    r.wnorm = wf;
                }
            }
        }
        if ((rc = up((void **)&f->d_ext_wide_nodes, wn.data(), wn.size() * sizeof(WideNode)))) return rc;
        if ((rc = up((void **)&f->d_ext_tree_slot, tslot.data(), tslot.size() * 4))) return rc;
    }
    if ((rc = up((void **)&f->d_ext_tree_node, tree_node.data(), tree_node.size() * 8))) return rc;

    {
        bool wsafe = true;   // finite and |w| <= 2^40: precondition of the f32 fast paths
        for (float wv : w)
            if (!(std::fabs(wv) <= 0x1p40f)) wsafe = false;
        f->ext_w_safe = wsafe;
    }
    lap("bounds, wide-node records");
    // (the per-tree blobs of the dense CUDA-core kernel are built on first use: ensure_ext_blob)
    rc = commit();
    lap("arena allocated, uploads queued");
    if (rc) return rc;
    if (dev && internal > 0) {
        int64_t *d_src = nullptr;
        int32_t *d_flag = nullptr;
        IFB_CUDA(cudaMalloc((void **)&d_src, (size_t)internal * 8 + 16));
        d_flag = reinterpret_cast<int32_t *>(d_src + internal);
        IFB_CUDA(cudaMemcpyAsync(d_src, slot_src.data(), (size_t)internal * 8, cudaMemcpyHostToDevice, 0));
        IFB_CUDA(cudaMemsetAsync(d_flag, 0, 4, 0));
        ext_gather_weights_kernel<<<(unsigned)((internal + 7) / 8), 256>>>(dev->w, d_src, internal, k, f->d_ext_w, f->d_ext_wabs,
                                                                            f->d_ext_wnorm, d_flag);
        ext_patch_wide_nodes_kernel<<<(unsigned)((total + 255) / 256), 256>>>(reinterpret_cast<WideNode *>(f->d_ext_wide_nodes),
                                                                                total, f->d_ext_wnorm);
        count_launch(2);
        int32_t unsafe = 0;
        cudaError_t e = cudaMemcpy(&unsafe, d_flag, 4, cudaMemcpyDeviceToHost);
        cudaFree(d_src);
        IFB_CUDA(e);
        IFB_CUDA(cudaGetLastError());
        f->ext_w_safe = unsafe == 0;
    }
    // tensor-core layout (fully-extended forests, and sparse hyperplanes as zero-padded columns; a forest that does not
    // qualify simply keeps tc_ok = false)
    if (internal > 0 && getenv("IFB_EXT_NO_TC") == nullptr && (f->ext_dense_identity || getenv("IFB_TC_NO_SPARSE") == nullptr)) {
        rc = build_ext_tc_tables(f, child, hp, leaf, off, depthv, len);
        if (rc) return rc;
        lap("tensor-core tables");
    }
    return IFB_OK;
}

// Per-tree blobs of score_ext_dense_kernel (fully-extended forests with k <= 64; the fallback when the tensor-core path
// does not take a call: IFB_EXT_NO_TC, a matrix wider than the forest, weights outside the fp16-safe range).  Built on
// first use from the device-resident tables, so that creating / fitting a forest never pays for them.
int ensure_ext_blob(ifb_forest *f) {
    std::lock_guard<std::mutex> lk(f->plan_mu);
    if (f->ext_blob_tried) return IFB_OK;
    f->ext_blob_tried = true;
    const int k = f->max_nnz;
    if (!(f->ext_dense_identity && k <= 64) || f->ext_internal_slots == 0) return IFB_OK;
    const int T = f->num_trees;
    const int64_t total = f->node_off[T];
    DeviceGuard dg(f->device);
    std::vector<int32_t> child((size_t)total), hp((size_t)total);
    std::vector<float> leaf((size_t)total), w((size_t)f->ext_internal_slots * k);
    std::vector<double> off((size_t)total);
    IFB_CUDA(cudaMemcpy(child.data(), f->d_ext_child, (size_t)total * 4, cudaMemcpyDeviceToHost));
    IFB_CUDA(cudaMemcpy(hp.data(), f->d_ext_hp, (size_t)total * 4, cudaMemcpyDeviceToHost));
    IFB_CUDA(cudaMemcpy(leaf.data(), f->d_ext_leaf, (size_t)total * 4, cudaMemcpyDeviceToHost));
    IFB_CUDA(cudaMemcpy(off.data(), f->d_ext_off, (size_t)total * 8, cudaMemcpyDeviceToHost));
    IFB_CUDA(cudaMemcpy(w.data(), f->d_ext_w, w.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<unsigned char> blob;
    std::vector<int64_t> boff((size_t)T + 1, 0);
    {
        const int D = k <= 8 ? 8 : k <= 16 ? 16 : k <= 32 ? 32 : 64;
        const int WS = D + 4;  // row stride in floats: 16-byte units odd => per-lane LDS.128 gathers spread over banks
        int64_t mx = 0;
        for (int t = 0; t < T; t++) {
            const int64_t base = f->node_off[t];
            const int n = f->node_off[t + 1] - f->node_off[t];
            int ni = 0;
            for (int q = 0; q < n; q++) ni += (child[base + q] >= 0);
            const int npad = (n + 3) & ~3, ipad = (ni + 1) & ~1;
            const size_t bytes = 16 + (size_t)npad * 12 + (size_t)ipad * 16 + (size_t)ni * WS * 4;
            const size_t bpad = (bytes + 15) & ~(size_t)15;
            const size_t at = blob.size();
            blob.resize(at + bpad, 0);
            unsigned char *B = blob.data() + at;
            int32_t *hdr = (int32_t *)B;
            hdr[0] = n; hdr[1] = ni; hdr[2] = npad; hdr[3] = ipad;
            int32_t *bchild = (int32_t *)(B + 16), *bslot = bchild + npad;
            float *bleaf = (float *)(bslot + npad);
            double *boffs = (double *)(bleaf + npad);
            double *bwn = boffs + ipad;            // ||w||_2 of the slot, inflated: bound of the f32 fast path
            float *bw = (float *)(bwn + ipad);
            int sl = 0;
            for (int q = 0; q < n; q++) {
                const int64_t g = base + q;
                bchild[q] = child[g];
                bleaf[q] = leaf[g];
                if (child[g] >= 0) {
                    bslot[q] = sl;
                    boffs[sl] = off[g];
                    const int64_t src = (int64_t)hp[g] * k;
                    double sq = 0.0;
                    for (int i = 0; i < k; i++) {
                        bw[(size_t)sl * WS + i] = w[src + i];
                        sq += (double)w[src + i] * (double)w[src + i];
                    }
                    bwn[sl] = std::sqrt(sq) * 1.0000001;
                    sl++;
                } else {
                    bslot[q] = -1;
                }
            }
            boff[t + 1] = (int64_t)blob.size();
            mx = std::max<int64_t>(mx, (int64_t)bpad);
        }
        unsigned char *dev_blob = nullptr;
        const size_t b_blob = (blob.size() + 255) & ~(size_t)255;
        IFB_CUDA(cudaMalloc((void **)&dev_blob, b_blob + boff.size() * 8));
        cudaError_t e = cudaMemcpy(dev_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dev_blob + b_blob, boff.data(), boff.size() * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(dev_blob);
            IFB_CUDA(e);
        }
        f->d_ext_blob = dev_blob;
        f->d_ext_blob_off = reinterpret_cast<int64_t *>(dev_blob + b_blob);
        f->device_bytes += (int64_t)(b_blob + boff.size() * 8);
        f->ext_blob_max = mx;
        f->ext_blob_D = D;   // last: launch_score_extended keys on it
    }
    return IFB_OK;
}

// ------------------------------------------------------------------------------------------------
// Standard-forest shared-memory plan for a given feature count d.
//   smem = [ val words | meta words ] of one chunk  +  stages * (d+1) * R floats of row tiles + barriers
// ------------------------------------------------------------------------------------------------
int get_std_plan(ifb_forest *f, int32_t d, ifb_forest::StdPlan **out) {
    std::lock_guard<std::mutex> lk(f->plan_mu);
    for (auto *p : f->std_plans)
        if (p->d == d) {
            *out = p;
            return IFB_OK;
        }
    IFB_REQUIRE(d >= 1, "d must be >= 1, got %d", d);
    if (d > 16382) {  // feature index no longer fits the 16-bit node encoding
        *out = nullptr;
        return IFB_OK;
    }
    IFB_REQUIRE(f->max_feature_index < d, "forest reads feature index %d but the matrix has only %d columns",
                f->max_feature_index, d);
    const int smem_max = device_smem_optin(f->device);
    IFB_REQUIRE(smem_max >= 200 * 1024, "device %d offers only %d bytes of shared memory per block", f->device,
                smem_max);
    const int T = f->num_trees;
    int64_t largest_tree = 0;
    for (int t = 0; t < T; t++) largest_tree = std::max<int64_t>(largest_tree, f->bfs_off[t + 1] - f->bfs_off[t]);
    const int64_t total_nodes = f->bfs_off[T];

    // pick the widest row tile that leaves room for the whole forest if possible, else for >= 48 KB of it
    auto *p = new ifb_forest::StdPlan();
    p->d = d;
    const int64_t overhead = 1024;  // barriers, tree-root table slack, alignment
    // Row-tile width R (= threads per CTA).  Prefer the widest tile that still leaves room for the whole
    // forest (single pass over the matrix); otherwise the widest tile <= 256 rows that leaves >= 64 KB for
    // one chunk of trees (several passes, sums carried in path_sum between them).
    const int64_t need_whole = (total_nodes + 1) * 8 + (int64_t)T * 4;
    const int64_t need_one = (largest_tree + 1) * 8 + 64;
    auto room_for = [&](int cand, int stages) { return (int64_t)smem_max - overhead - (int64_t)stages * (d + 1) * cand * 4; };
    // Candidates in order of preference: (rows per tile = threads per CTA, ring depth).  Wide tiles matter more than
    // double buffering: a tile keeps the CTA busy for tens of microseconds, its TMA fill costs ~1-3.
    static const bool try1024 = getenv("IFB_STD_NO_1024") == nullptr;
    const int kCandAll[][2] = {{1024, 1}, {512, 2}, {512, 1}, {256, 2}, {256, 1}, {128, 2}, {128, 1}, {64, 2}, {32, 2}};
    const int (*kCand)[2] = try1024 ? kCandAll : kCandAll + 1;
    const int nCand = try1024 ? 9 : 8;
    static const bool single_ok = getenv("IFB_STD_NO_SINGLE_STAGE") == nullptr;
    int R = 0, S = 2;
    for (int ci = 0; ci < nCand; ci++) {   // pass 1: the whole forest fits next to WIDE tiles (>= 8 warps per CTA)
        const int *c = kCand[ci];
        if (c[1] == 1 && !single_ok) continue;
        // a resident forest is not worth 32..128-row tiles (1-4 warps per SM: measured 4x slower per row-tree than
        // 256-row tiles with the forest cut into chunks -- 256 trees, d = 128: 95 ms vs ~24 ms for 25M rows)
        if (c[0] < 256) continue;
        if (room_for(c[0], c[1]) >= std::max(need_whole, need_one)) {
            R = c[0];
            S = c[1];
            break;
        }
    }
    if (R == 0)
        for (int ci = 0; ci < nCand; ci++) {   // pass 2: chunked forest, >= 64 KB of trees per pass, tiles of <= 256 rows
            const int *c = kCand[ci];
            if (c[0] > 256 || (c[1] == 1 && !single_ok)) continue;
            if (room_for(c[0], c[1]) >= std::max<int64_t>(64 * 1024, need_one)) {
                R = c[0];
                S = c[1];
                break;
            }
        }
    if (R == 0)
        for (int ci = 0; ci < nCand; ci++) {   // pass 3: narrow tiles, whole forest or as many trees as fit
            const int *c = kCand[ci];
            if (c[1] == 1 && !single_ok) continue;
            if (room_for(c[0], c[1]) >= std::max<int64_t>(std::min<int64_t>(need_whole, 48 * 1024), need_one)) {
                R = c[0];
                S = c[1];
                break;
            }
        }
    if (R == 0 && room_for(32, 2) >= need_one) {
        R = 32;
        S = 2;
    }
    if (R == 0) {  // rows too wide for a tile next to one tree: the generic kernel takes over
        delete p;
        *out = nullptr;
        return IFB_OK;
    }
    p->stages = S;
    p->rows_per_tile = R;
    const int64_t room = smem_max - overhead - (int64_t)S * (d + 1) * R * 4;
    // greedy chunking; each chunk: 1 pad word + nodes, plus 4 bytes per tree for the root table
    std::vector<float> val;
    std::vector<uint32_t> meta, roots(T, 0);
    int t = 0;
    while (t < T) {
        StdChunk c;
        c.tree_begin = t;
        c.node_begin = (int32_t)val.size();
        int64_t words = 1;  // pad slot 0 so that "self - 1" of a root leaf is valid
        int64_t trees = 0;
        while (t < T) {
            int64_t n = f->bfs_off[t + 1] - f->bfs_off[t];
            if (((words + n) * 8 + (trees + 1) * 4 > room || trees >= std_top_table_max_trees()) && trees > 0) break;
            words += n;
            trees++;
            t++;
        }
        c.tree_end = t;
        c.node_count = (int32_t)words;
        IFB_REQUIRE(words < (1 << 16), "chunk too large for 16-bit child indices");
        val.push_back(0.f);
        meta.push_back(0u);
        int64_t w = 1;
        for (int tt = c.tree_begin; tt < c.tree_end; tt++) {
            const int32_t base = f->bfs_off[tt], n = f->bfs_off[tt + 1] - base;
            roots[tt] = (uint32_t)(w * 4);  // byte offset of the root inside the chunk's val array
            for (int32_t q = 0; q < n; q++) {
                const int64_t g = base + q;
                const int64_t self = w + q;
                val.push_back(f->h_val[g]);
                if (f->h_child[g] < 0) {
                    meta.push_back(((uint32_t)(self - 1) << 16) | (uint32_t)d);
                } else {
                    meta.push_back(((uint32_t)(w + f->h_child[g]) << 16) | f->h_meta_feat[g]);
                }
            }
            w += n;
        }
        p->chunks.push_back(c);
    }
    // kernel-parameter tables (tree levels 0 and 1) per chunk
    p->h_top.assign(p->chunks.size() * std_top_table_bytes(), 0);
    for (size_t ci = 0; ci < p->chunks.size(); ci++) {
        const StdChunk &c = p->chunks[ci];
        std_fill_top_table(p->h_top.data() + ci * std_top_table_bytes(), val.data() + c.node_begin,
                           meta.data() + c.node_begin, roots.data() + c.tree_begin, c.tree_end - c.tree_begin,
                           std_rows_per_box(R, S));
    }
    p->total_words = (int64_t)val.size();
    DeviceGuard dg(f->device);
    // one allocation and one copy for the three tables (a cudaMalloc + blocking cudaMemcpy each was a quarter of
    // BASELINE config 1's fit + transform): [ val | meta | roots ], every part 256-byte aligned
    {
        auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t b_val = al(val.size() * 4), b_meta = al(meta.size() * 4), b_root = al(roots.size() * 4);
        std::vector<unsigned char> pack(b_val + b_meta + b_root, 0);
        std::memcpy(pack.data(), val.data(), val.size() * 4);
        std::memcpy(pack.data() + b_val, meta.data(), meta.size() * 4);
        std::memcpy(pack.data() + b_val + b_meta, roots.data(), roots.size() * 4);
        unsigned char *base = nullptr;
        IFB_CUDA(cudaMalloc((void **)&base, std::max<size_t>(256, pack.size())));
        p->d_val = reinterpret_cast<float *>(base);
        p->d_meta = reinterpret_cast<uint32_t *>(base + b_val);
        p->d_tree_root = reinterpret_cast<uint32_t *>(base + b_val + b_meta);
        IFB_CUDA(cudaMemcpy(base, pack.data(), pack.size(), cudaMemcpyHostToDevice));
        f->device_bytes += (int64_t)pack.size();
    }
    f->std_plans.push_back(p);
    *out = p;
    return IFB_OK;
}

int ensure_std_generic_tables(ifb_forest *f) {
    std::lock_guard<std::mutex> lk(f->plan_mu);
    if (f->d_gval) return IFB_OK;
    const size_t n = f->h_val.size();
    std::vector<int32_t> feat(n);
    for (size_t i = 0; i < n; i++) feat[i] = f->h_child[i] < 0 ? -1 : (int32_t)f->h_meta_feat[i];
    DeviceGuard dg(f->device);
    IFB_CUDA(cudaMalloc((void **)&f->d_gval, std::max<size_t>(16, n * 4)));
    IFB_CUDA(cudaMalloc((void **)&f->d_gfeat, std::max<size_t>(16, n * 4)));
    IFB_CUDA(cudaMalloc((void **)&f->d_gchild, std::max<size_t>(16, n * 4)));
    IFB_CUDA(cudaMalloc((void **)&f->d_groot, (f->bfs_off.size()) * 4));
    IFB_CUDA(cudaMemcpy(f->d_gval, f->h_val.data(), n * 4, cudaMemcpyHostToDevice));
    IFB_CUDA(cudaMemcpy(f->d_gfeat, feat.data(), n * 4, cudaMemcpyHostToDevice));
    IFB_CUDA(cudaMemcpy(f->d_gchild, f->h_child.data(), n * 4, cudaMemcpyHostToDevice));
    IFB_CUDA(cudaMemcpy(f->d_groot, f->bfs_off.data(), f->bfs_off.size() * 4, cudaMemcpyHostToDevice));
    f->device_bytes += (int64_t)(n * 12 + f->bfs_off.size() * 4);
    return IFB_OK;
}

}  // namespace ifb

namespace ifb {
// Forest from the device builder's output without a host copy of the hyperplanes (fit.cu, k == d): the node
// structure comes from the host tables, the weight rows are gathered on the device.  The builder guarantees the
// SplitHyperplane invariants (ExtendedUtils.scala:27-34): k distinct ascending indices 0..k-1.
int create_extended_from_device(int32_t device, int32_t num_trees, const int32_t *node_off, const int32_t *left,
                                const int32_t *right, const int64_t *num_instances, const double *offset, int32_t k,
                                const DeviceHyperplanes &dev, int32_t num_samples, int32_t total_num_features,
                                ifb_forest **out) {
    *out = nullptr;
    int rc = validate_shape(num_trees, node_off, left, right, num_instances, /*allow_empty_leaf=*/true);
    if (rc) return rc;
    const int64_t total = node_off[num_trees];
    auto *f = new ifb_forest();
    f->device = device;
    f->extended = true;
    f->num_trees = num_trees;
    f->num_samples = num_samples;
    f->total_num_features = total_num_features;
    f->avg_path_norm = avg_path_length_host(num_samples);
    f->max_nnz = k;
    f->max_feature_index = k - 1;
    f->node_off.assign(node_off, node_off + num_trees + 1);
    f->left.assign(left, left + total);
    f->right.assign(right, right + total);
    f->offset.assign(offset, offset + total);
    f->num_instances.assign(num_instances, num_instances + total);
    f->hp_off.assign(total + 1, 0);
    for (int64_t g = 0; g < total; g++) f->hp_off[g + 1] = f->hp_off[g] + (left[g] != -1 ? k : 0);
    f->hp_lazy = true;
    rc = build_extended_tables(f, &dev);
    if (rc) {
        delete f;
        return rc;
    }
    *out = f;
    return IFB_OK;
}
}  // namespace ifb

ifb_forest::~ifb_forest() {
    ifb::DeviceGuard dg(device);
    for (auto *p : std_plans) {
        cudaFree(p->d_val);   // d_meta and d_tree_root are slices of the same allocation
        delete p;
    }
    ifb::free_rank_plans(this);
    cudaFree(d_gval);
    cudaFree(d_gfeat);
    cudaFree(d_gchild);
    cudaFree(d_groot);
    cudaFree(d_ext_arena);   // every d_ext_* table is a slice of it ...
    cudaFree(d_ext_blob);    // ... except the lazily built dense-kernel blobs (d_ext_blob_off is a slice of this one)
    cudaFree(d_tc_arena);    // every d_tc_* table is a slice of it ...
    cudaFree(d_tc_slot_len); // ... except the per-slot term counts of sparse forests
}

using namespace ifb;

extern "C" {

int ifb_abi_version(void) { return IFB_ABI_VERSION; }
const char *ifb_last_error(void) { return g_last_error.c_str(); }
float ifb_avg_path_length(int64_t n) { return avg_path_length_host(n); }
int64_t ifb_kernel_launch_count(int32_t reset) {
    return reset ? (int64_t)g_launches.exchange(0) : (int64_t)g_launches.load();
}

int ifb_device_count(int32_t *count) {
    IFB_REQUIRE(count, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        set_error("no CUDA device: %s", cudaGetErrorString(e));
        return IFB_ENOGPU;
    }
    *count = n;
    return IFB_OK;
}

int ifb_host_alloc(size_t bytes, void **ptr) {
    IFB_REQUIRE(ptr, "ptr is null");
    IFB_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return IFB_OK;
}
int ifb_host_free(void *ptr) {
    if (ptr) IFB_CUDA(cudaFreeHost(ptr));
    return IFB_OK;
}
// Plain cudaMalloc / cudaFree on purpose: these buffers are what ifb_ipc_export hands to peer processes (the fused
// tree-sharded scatter), and cudaIpcGetMemHandle refuses memory from the stream-ordered pool.
int ifb_device_alloc(int32_t device, size_t bytes, void **ptr) {
    IFB_REQUIRE(ptr, "ptr is null");
    DeviceGuard dg(device);
    IFB_CUDA(cudaMalloc(ptr, bytes ? bytes : 1));
    return IFB_OK;
}
int ifb_device_free(int32_t device, void *ptr) {
    DeviceGuard dg(device);
    if (ptr) IFB_CUDA(cudaFree(ptr));
    return IFB_OK;
}

int ifb_copy_to_device(int32_t device, void *dst, const void *src, size_t bytes) {
    IFB_REQUIRE(bytes == 0 || (dst && src), "null buffer");
    DeviceGuard dg(device);
    IFB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return IFB_OK;
}
int ifb_copy_to_host(int32_t device, void *dst, const void *src, size_t bytes) {
    IFB_REQUIRE(bytes == 0 || (dst && src), "null buffer");
    DeviceGuard dg(device);
    IFB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return IFB_OK;
}

static int check_device(int32_t device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); this engine has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return IFB_ENOGPU;
    }
    IFB_REQUIRE(device >= 0 && device < n, "device %d out of range [0,%d)", device, n);
    // cudaGetDeviceProperties costs about a millisecond per call; two attributes are all that is needed
    int major = 0, minor = 0;
    IFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    IFB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, major, minor);
        return IFB_ENOGPU;
    }
    return IFB_OK;
}

int ifb_forest_create_standard(int32_t device, int32_t num_trees, const int32_t *node_off, const int32_t *left,
                               const int32_t *right, const int32_t *feature, const double *threshold,
                               const int64_t *num_instances, int32_t num_samples, int32_t total_num_features,
                               ifb_forest **out) {
    IFB_REQUIRE(out, "out is null");
    *out = nullptr;
    int rc = check_device(device);
    if (rc) return rc;
    // IsolationForestModel constructor requires (IF/IsolationForestModel.scala:61-78)
    IFB_REQUIRE(num_samples > 0, "parameter numSamples must be >0, but given invalid value %d", num_samples);
    IFB_REQUIRE(total_num_features == -1 || total_num_features > 0,
                "parameter totalNumFeatures must be >0 or UnknownTotalNumFeatures, but given invalid value %d",
                total_num_features);
    const bool timing = getenv("IFB_FIT_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (timing)
            std::fprintf(stderr, "[ifb create] %s at %.3f ms\n", what,
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    rc = validate_shape(num_trees, node_off, left, right, num_instances, /*allow_empty_leaf=*/false);
    if (rc) return rc;
    lap("shape validated");
    const int64_t total = node_off[num_trees];
    IFB_REQUIRE(total == 0 || (feature && threshold), "null split table");
    for (int64_t g = 0; g < total; g++) {
        if (left[g] != -1) {
            // InternalNode requires splitAttribute >= 0 (IF/Nodes.scala:55-59)
            IFB_REQUIRE(feature[g] >= 0, "node %lld: splitAttribute %d must be >= 0", (long long)g, feature[g]);
            IFB_REQUIRE(total_num_features == -1 || feature[g] < total_num_features,
                        "node %lld: splitAttribute %d outside the %d training features", (long long)g, feature[g],
                        total_num_features);
            IFB_REQUIRE(!std::isnan(threshold[g]), "node %lld: splitValue is NaN", (long long)g);
        }
    }
    auto *f = new ifb_forest();
    f->device = device;
    f->extended = false;
    f->num_trees = num_trees;
    f->num_samples = num_samples;
    f->total_num_features = total_num_features;
    f->avg_path_norm = avg_path_length_host(num_samples);
    f->node_off.assign(node_off, node_off + num_trees + 1);
    f->left.assign(left, left + total);
    f->right.assign(right, right + total);
    f->feature.assign(feature, feature + total);
    f->threshold.assign(threshold, threshold + total);
    f->num_instances.assign(num_instances, num_instances + total);
    lap("tables copied");
    rc = build_standard_tables(f);
    if (rc) {
        delete f;
        return rc;
    }
    lap("kernel layout built");
    *out = f;
    return IFB_OK;
}

int ifb_forest_create_extended(int32_t device, int32_t num_trees, const int32_t *node_off, const int32_t *left,
                               const int32_t *right, const int64_t *num_instances, const double *offset,
                               const int64_t *hp_off, const int32_t *hp_idx, const float *hp_w,
                               int32_t num_samples, int32_t total_num_features, ifb_forest **out) {
    IFB_REQUIRE(out, "out is null");
    *out = nullptr;
    int rc = check_device(device);
    if (rc) return rc;
    IFB_REQUIRE(num_samples > 0, "parameter numSamples must be >0, but given invalid value %d", num_samples);
    IFB_REQUIRE(total_num_features == -1 || total_num_features > 0,
                "parameter totalNumFeatures must be >0 or UnknownTotalNumFeatures, but given invalid value %d",
                total_num_features);
    // ExtendedExternalNode allows numInstances == 0 (IF/extended/ExtendedNodes.scala:28-38)
    rc = validate_shape(num_trees, node_off, left, right, num_instances, /*allow_empty_leaf=*/true);
    if (rc) return rc;
    const int64_t total = node_off[num_trees];
    IFB_REQUIRE(total == 0 || (offset && hp_off), "null hyperplane table");
    IFB_REQUIRE(total == 0 || hp_off[0] == 0, "hp_off[0] must be 0");
    int32_t max_nnz = 1, max_idx = -1;
    for (int64_t g = 0; g < total; g++) {
        const int64_t b = hp_off[g], e = hp_off[g + 1];
        IFB_REQUIRE(e >= b, "hp_off not monotone at node %lld", (long long)g);
        if (left[g] == -1) {
            IFB_REQUIRE(e == b, "leaf node %lld carries a hyperplane", (long long)g);
            continue;
        }
        // SplitHyperplane invariants (IF/extended/ExtendedUtils.scala:27-34)
        IFB_REQUIRE(e > b, "indices must be non-empty.");
        IFB_REQUIRE(e - b <= (1 << 20), "hyperplane too wide");
        IFB_REQUIRE(hp_idx && hp_w, "null hyperplane arrays");
        for (int64_t i = b; i < e; i++) {
            IFB_REQUIRE(hp_idx[i] >= 0, "indices must be non-negative.");
            IFB_REQUIRE(i == b || hp_idx[i] != hp_idx[i - 1], "indices must be distinct.");
            IFB_REQUIRE(i == b || hp_idx[i] > hp_idx[i - 1], "indices must be sorted in ascending order.");
            IFB_REQUIRE(total_num_features == -1 || hp_idx[i] < total_num_features,
                        "hyperplane index %d outside the %d training features", hp_idx[i], total_num_features);
            max_idx = std::max(max_idx, hp_idx[i]);
        }
        IFB_REQUIRE(!std::isnan(offset[g]), "node %lld: offset is NaN", (long long)g);
        max_nnz = std::max<int32_t>(max_nnz, (int32_t)(e - b));
    }
    auto *f = new ifb_forest();
    f->device = device;
    f->extended = true;
    f->num_trees = num_trees;
    f->num_samples = num_samples;
    f->total_num_features = total_num_features;
    f->avg_path_norm = avg_path_length_host(num_samples);
    f->max_nnz = max_nnz;
    f->max_feature_index = max_idx;
    f->node_off.assign(node_off, node_off + num_trees + 1);
    f->left.assign(left, left + total);
    f->right.assign(right, right + total);
    f->offset.assign(offset, offset + total);
    f->num_instances.assign(num_instances, num_instances + total);
    f->hp_off.assign(hp_off, hp_off + total + 1);
    const int64_t nh = total ? hp_off[total] : 0;
    f->hp_idx.assign(hp_idx, hp_idx + nh);
    f->hp_w.assign(hp_w, hp_w + nh);
    rc = build_extended_tables(f);
    if (rc) {
        delete f;
        return rc;
    }
    *out = f;
    return IFB_OK;
}

int ifb_forest_destroy(ifb_forest *forest) {
    delete forest;
    return IFB_OK;
}

int ifb_forest_get_info(const ifb_forest *f, ifb_forest_info *info) {
    IFB_REQUIRE(f && info, "null argument");
    info->extended = f->extended;
    info->device = f->device;
    info->num_trees = f->num_trees;
    info->num_samples = f->num_samples;
    info->total_num_features = f->total_num_features;
    info->max_feature_index = f->max_feature_index;
    info->max_depth = f->max_depth;
    info->max_nnz = f->max_nnz;
    info->num_nodes = f->node_off.empty() ? 0 : f->node_off.back();
    info->num_hp_entries = f->hp_lazy ? (f->hp_off.empty() ? 0 : f->hp_off.back()) : (int64_t)f->hp_idx.size();
    info->device_bytes = f->device_bytes;
    return IFB_OK;
}

int ifb_forest_export(const ifb_forest *f, int32_t *node_off, int32_t *left, int32_t *right, int32_t *feature,
                      double *threshold, int64_t *num_instances, double *offset, int64_t *hp_off,
                      int32_t *hp_idx, float *hp_w) {
    IFB_REQUIRE(f, "forest is null");
    auto cp = [](auto *dst, const auto &v) {
        if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
    };
    cp(node_off, f->node_off);
    cp(left, f->left);
    cp(right, f->right);
    cp(num_instances, f->num_instances);
    if (f->extended) {
        cp(offset, f->offset);
        cp(hp_off, f->hp_off);
        if (f->hp_lazy) {
            // the weights never left the device: gather them now (pre-order rows, k identity indices each)
            const int k = f->max_nnz;
            const int64_t total = f->node_off[f->num_trees], slots = f->ext_internal_slots;
            std::vector<float> wdev((size_t)slots * k);
            if (slots > 0 && (hp_w || hp_idx)) {
                DeviceGuard dg(f->device);
                if (hp_w) IFB_CUDA(cudaMemcpy(wdev.data(), f->d_ext_w, wdev.size() * 4, cudaMemcpyDeviceToHost));
                for (int64_t g = 0; g < total; g++) {
                    const int32_t sl = f->lazy_slot_of_node[(size_t)g];
                    if (sl < 0) continue;
                    const int64_t at = f->hp_off[g];
                    if (hp_w) std::memcpy(hp_w + at, wdev.data() + (size_t)sl * k, (size_t)k * 4);
                    if (hp_idx)
                        for (int i = 0; i < k; i++) hp_idx[at + i] = i;
                }
            }
        } else {
            cp(hp_idx, f->hp_idx);
            cp(hp_w, f->hp_w);
        }
    } else {
        cp(feature, f->feature);
        cp(threshold, f->threshold);
    }
    return IFB_OK;
}

}  // extern "C"
