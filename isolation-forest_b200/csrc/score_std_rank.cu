// Standard isolation-forest scoring on per-feature RANKS (matrices of <= 32 features) for sm_100a.
//
// Same contract as score_std.cu (IF/IsolationTree.scala:196-230 walk, IF/IsolationForestModel.scala:131-139 epilogue),
// different arithmetic: a visit only needs the ORDER of x[f] against the node's threshold, so
//   * the thresholds of a forest chunk are sorted per feature (the "cuts" c_0 < c_1 < ... of feature f);
//   * every feature value of a row tile is replaced IN PLACE by the word  (B + #{cuts <= x}) << 16 | 0xFFFF
//     (found with a 256-cell grid lookup + a short linear scan; exact for every float incl. NaN / +-inf);
//   * a node is ONE 32-bit word  (B + j + 1) << 16 | feature << 11 | byte offset of its child pair inside the
//     tree's 2 KB block, so  x < c_j  <=>  rank word < node word  (as unsigned integers and -- because both
//     are positive normal floats -- as f32, which lets FSET produce the branch mask in one instruction).
// A visit is then  LDS node word, LOP3 (feature field | row address), LDS rank word, FSET, LOP3 (block | pair field),
// LOP3 (| 4 when not less)  = 6 instructions and 2 shared-memory wavefronts instead of 8 and 3 in score_std.cu,
// which is what the profile of that kernel says it is bound by (shared-memory wavefronts 96.7 % of peak).
//
// Leaves map onto themselves without a pseudo feature: a leaf in a left slot carries a rank code above every row
// code (never "not less"), one in a right slot a code below every row code (always "not less"); the spare 13 bits
// of the code index a table of the chunk's distinct leaf values ((float)depth + c(numInstances), precomputed).
//
// Shared-memory map (absolute shared addresses; the feature field of a node word is OR-ed into the row address, so
// the two 64 KB sub-tiles must sit on 64 KB boundaries):
//   [a_w, 64K)            node words in 2 KB blocks (every tree inside one block), then the grid table (32 x 256 u16)
//   [64K, 192K)           two sub-tiles [feature 0..31][512 rows] of one 512-thread group each
//   [192K, end)           cuts (f32, per feature, +inf sentinel), leaf values
// The two groups of a CTA fetch (one 2 KB cp.async.bulk per feature column), convert, walk and refill their sub-tile
// independently (mbarrier + named barrier per group), so one group's refill hides behind the other's walk.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <type_traits>

#include "ifb_internal.h"

namespace ifb {

namespace {

constexpr int kG = 512;                          // rows per group
constexpr int kNG = 2;                           // groups per CTA
constexpr int kThreads = kG * kNG;
constexpr uint32_t kColBytes = kG * 4;           // one feature column of a sub-tile
constexpr uint32_t kFeatMask = 31u * kColBytes;  // 0xF800: feature field of a node word == byte offset of its column
constexpr uint32_t kSubBytes = 32 * kColBytes;   // 64 KB
constexpr uint32_t kTileAbs = 65536;
constexpr uint32_t kHighAbs = kTileAbs + kNG * kSubBytes;
constexpr uint32_t kBlockBytes = 2048, kBlockWords = 512;
constexpr int kCells = 256;
constexpr uint32_t kGridBytes = 32 * kCells * 2;
constexpr uint32_t kCodeBase = 8192;             // row codes are kCodeBase + rank; leaf payloads sit below / above
constexpr uint32_t kLeftLeaf = 0x4000;           // code bit of a leaf in a left slot (above every row code)
constexpr int kMaxCutsPerFeature = 8000;
constexpr int kMaxLeafValues = 8191;
constexpr int kMaxTrees = 256;
constexpr uint32_t kDynSmem = 232448;            // 227 KB: the opt-in maximum of sm_100

struct RankTop {
    uint32_t w0, f0;      // root word, byte offset of its feature column
    uint32_t wL, wR;      // the two level-1 words
    uint32_t aL, aR;      // absolute shared address of the pair each of them points to
    uint32_t pad[2];
};
struct RankTopTable {
    RankTop e[kMaxTrees];
};

struct RankParams {
    const float *X;
    int64_t n_rows, ld;
    int32_t d, use_bulk;
    const uint32_t *gw;
    const float *glv, *gcut;
    int32_t w_words, n_lv, cut_words;
    uint32_t a_w, a_lv, a_grid, a_cut, sbase;
    uint32_t cut_off[32], n_cut[32];
    float inv[32], c0[32];
    int32_t n_trees, max_depth, total_trees;
    float avg_path;
    int32_t first_chunk, finalize, finalize_scatter;
    double *scores;
    float *path_sum;
    int64_t n_gtiles;
    ScatterTarget scatter;
    uint32_t *probe_out;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
// Address of the next node: `pair` (bit 2 clear) when the rank word x is below the node word w, pair + 4 otherwise.
// Both words are positive floats (normal or denormal, never NaN), so the f32 compare is the unsigned integer compare.
// PTX pins the shape: FSETP + one predicated LOP3 (the compiler's own choice is FSETP + SEL + 2 LOP3).
__device__ __forceinline__ uint32_t pick_child(uint32_t pair, uint32_t x, uint32_t w) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, 4;\n\t}"
        : "+r"(pair)
        : "f"(__uint_as_float(x)), "f"(__uint_as_float(w)));
    return pair;
}
// (node & ~0x7FF) | (w & 0x7FF): the block of the current node + the pair offset stored in its word (bits 0..2 are 0)
__device__ __forceinline__ uint32_t pair_of(uint32_t node, uint32_t w) {
    uint32_t t;
    asm("lop3.b32 %0, %1, %2, 0x7FF, 0xD8;" : "=r"(t) : "r"(node), "r"(w));
    return t;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < 0x7fffffffu; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();   // a copy that never lands must surface as a launch failure, not as a hung GPU
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Grid cell of a (clamped, non-NaN) feature value: monotone non-decreasing in x, which is all the lookup needs.
// Explicit intrinsics: the same function classifies the cuts when the grid is built and the rows when they are ranked.
__device__ __forceinline__ uint32_t rank_cell(float xc, float inv, float c0) {
    const uint32_t q = __float2uint_rd(__fmaf_rn(xc, inv, c0));   // saturating: negative / NaN -> 0
    return q < (uint32_t)(kCells - 1) ? q : (uint32_t)(kCells - 1);
}

template <int DEEP>
__global__ void __launch_bounds__(kThreads, 1)
score_std_rank_kernel(const __grid_constant__ RankTopTable top, const __grid_constant__ RankParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    if (p.probe_out) {   // planner probe: where does the dynamic shared memory of THIS kernel start?
        if (threadIdx.x == 0 && blockIdx.x == 0) *p.probe_out = sbase;
        return;
    }
    if (sbase != p.sbase) __trap();
    const int tid = threadIdx.x;
    const int g = tid / kG, rl = tid % kG;
    const int d = p.d;
    const uint32_t bar = sbase + 8u * (uint32_t)g;
    unsigned char *const s0 = smem - sbase;   // s0 + absolute shared address == generic pointer

    if (tid == 0) {
        for (int i = 0; i < kNG; i++) mbar_init(sbase + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t stride = (int64_t)gridDim.x * kNG;
    int64_t gt = (int64_t)blockIdx.x * kNG + g;
    const uint32_t sub_abs = kTileAbs + (uint32_t)g * kSubBytes;
    auto is_bulk_tile = [&](int64_t t) { return p.use_bulk && (t + 1) * kG <= p.n_rows; };
    auto issue_fill = [&](int64_t t) {   // first warp of the group; one 2 KB column per lane
        if (rl == 0) mbar_expect_tx(bar, (uint32_t)d * kColBytes);
        __syncwarp();
        if (rl < d) bulk_g2s(sub_abs + (uint32_t)rl * kColBytes, p.X + (int64_t)rl * p.ld + t * kG, kColBytes, bar);
    };
    if (rl < 32 && gt < p.n_gtiles && is_bulk_tile(gt)) issue_fill(gt);

    // ---- chunk tables -> shared memory, then the grid of every feature --------------------------------------------
    {
        uint32_t *sw = reinterpret_cast<uint32_t *>(s0 + p.a_w);
        for (int i = tid; i < p.w_words; i += kThreads) sw[i] = p.gw[i];
        float *slv = reinterpret_cast<float *>(s0 + p.a_lv);
        for (int i = tid; i < p.n_lv; i += kThreads) slv[i] = p.glv[i];
        float *sc = reinterpret_cast<float *>(s0 + p.a_cut);
        for (int i = tid; i < p.cut_words; i += kThreads) sc[i] = p.gcut[i];
    }
    __syncthreads();
    {
        // grid[f][q] = 4 * #{cuts of f whose cell is < q}: every x of cell q is above all of them (rank_cell is
        // monotone), so the scan below may start there
        uint16_t *grid = reinterpret_cast<uint16_t *>(s0 + p.a_grid);
        for (int i = tid; i < d * kCells; i += kThreads) {
            const int f = i / kCells;
            const uint32_t q = (uint32_t)(i % kCells);
            const float *cf = reinterpret_cast<const float *>(s0 + p.a_cut + p.cut_off[f]);
            const float inv = p.inv[f], c0 = p.c0[f];
            uint32_t lo = 0, hi = p.n_cut[f];
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (rank_cell(cf[mid], inv, c0) < q) lo = mid + 1; else hi = mid;
            }
            grid[i] = (uint16_t)(lo * 4u);
        }
    }
    __syncthreads();

    const uint32_t xrow_s = sub_abs + (uint32_t)rl * 4u;
    const uint32_t code_k = (kCodeBase << 16) | 0xFFFFu;
    const int n_trees = p.n_trees;
    const int deep_levels = DEEP >= 0 ? DEEP : (p.max_depth > 2 ? p.max_depth - 2 : 0);

    for (int64_t k = 0; gt < p.n_gtiles; gt += stride, ++k) {
        const int64_t row = gt * kG + rl;
        const bool live = row < p.n_rows;
        const bool bulk = is_bulk_tile(gt);
        if (bulk) mbar_wait(bar, (uint32_t)(k & 1));

        // ---- feature values -> rank words, in place (a thread touches only its own row) --------------------------
#pragma unroll 4
        for (int f = 0; f < d; f++) {
            const uint32_t xa = xrow_s + (uint32_t)f * kColBytes;
            float x;
            if (bulk) x = lds_f32(xa);
            else x = live ? __ldg(p.X + (int64_t)f * p.ld + row) : 0.f;
            const float xc = fminf(x, FLT_MAX);   // +inf and NaN rank above every (finite) cut, like `x < c` being false
            const uint32_t q = rank_cell(xc, p.inv[f], p.c0[f]);
            uint32_t r4;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r4) : "r"(p.a_grid + (uint32_t)f * (kCells * 2) + q * 2u));
            const uint32_t cf = p.a_cut + p.cut_off[f];
            while (!(xc < lds_f32(cf + r4))) r4 += 4;   // ends at the +inf sentinel at the latest
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(xa), "r"((r4 << 14) + code_k) : "memory");
        }

        float s = 0.f;
        if (!p.first_chunk && live) s = p.path_sum[row];

        auto top_levels = [&](int t) -> uint32_t {
            const RankTop &e = top.e[t];
            const uint32_t x0 = lds_u32(xrow_s + e.f0);
            const bool lt0 = __uint_as_float(x0) < __uint_as_float(e.w0);
            const uint32_t w1 = lt0 ? e.wL : e.wR;
            const uint32_t a1 = lt0 ? e.aL : e.aR;
            const uint32_t x1 = lds_u32((w1 & kFeatMask) | xrow_s);
            return pick_child(a1, x1, w1);
        };
        auto walk_group = [&](int t0, auto cc_tag) {
            constexpr int CC = decltype(cc_tag)::value;
            uint32_t node[CC];
#pragma unroll
            for (int c = 0; c < CC; c++) node[c] = top_levels(t0 + c);
            auto level = [&]() {
#pragma unroll
                for (int c = 0; c < CC; c++) {
                    const uint32_t w = lds_u32(node[c]);
                    const uint32_t x = lds_u32((w & kFeatMask) | xrow_s);
                    node[c] = pick_child(pair_of(node[c], w), x, w);
                }
            };
            if constexpr (DEEP >= 0) {
#pragma unroll
                for (int lvl = 0; lvl < DEEP; lvl++) level();
            } else {
#pragma unroll 1
                for (int lvl = 0; lvl < deep_levels; lvl++) level();
            }
#pragma unroll
            for (int c = 0; c < CC; c++) {
                const uint32_t w = lds_u32(node[c]);
                s = s + lds_f32(p.a_lv + ((w >> 14) & 0x7FFCu));   // tree order, one f32 add per tree
            }
        };
        int t = 0;
        for (; t + 4 <= n_trees; t += 4) walk_group(t, std::integral_constant<int, 4>{});
        if (t + 2 <= n_trees) { walk_group(t, std::integral_constant<int, 2>{}); t += 2; }
        if (t < n_trees) walk_group(t, std::integral_constant<int, 1>{});

        if (live) {
            if (p.finalize) {
                // IF/IsolationForestModel.scala:137-138: Float sum / Int, -Float / Float, Math.pow(2, Double)
                const float e = __fdiv_rn(s, (float)p.total_trees);
                const float z = __fdiv_rn(-e, p.avg_path);
                p.scores[row] = exp2((double)z);
            }
            if (p.scatter.world > 0 && p.finalize_scatter) {
                int o = 0;
#pragma unroll
                for (int q = 1; q < kMaxScatterRanks; q++) o += (q < p.scatter.world && row >= p.scatter.cut[q]) ? 1 : 0;
                const int64_t r0 = p.scatter.cut[o], rows_o = p.scatter.cut[o + 1] - r0;
                p.scatter.peer[o][(int64_t)p.scatter.rank * rows_o + (row - r0)] = s;
            } else if (p.path_sum) {
                p.path_sum[row] = s;
            }
        }
        // every access of THIS GROUP's sub-tile is done before it is refilled (the rank words were written through
        // the generic proxy, the refill arrives through the async proxy); the other group is not waited for
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(kG) : "memory");
        if (rl < 32 && gt + stride < p.n_gtiles && is_bulk_tile(gt + stride)) issue_fill(gt + stride);
    }
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------------------------------

struct RankChunk {
    int32_t tree_begin = 0, tree_end = 0;
    int64_t w_off = 0, lv_off = 0, cut_off_words = 0;   // slices of the plan's device arrays (words)
    int32_t w_words = 0, n_lv = 0, cut_words = 0;
    uint32_t a_w = 0, a_lv = 0, a_grid = 0, a_cut = 0;
    uint32_t cut_off[32], n_cut[32];
    float inv[32], c0[32];
    RankTopTable top;
};
struct RankPlan {
    int32_t d = -1;
    bool ok = false;
    uint32_t sbase = 0;
    std::vector<RankChunk> chunks;
    uint32_t *d_w = nullptr;
    float *d_lv = nullptr, *d_cut = nullptr;
};

namespace {

template <int DEEP>
int launch_rank_variant(const RankTopTable &top, const RankParams &p, int grid, cudaStream_t stream) {
    auto kern = score_std_rank_kernel<DEEP>;
    IFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmem));
    kern<<<grid, kThreads, kDynSmem, stream>>>(top, p);
    IFB_CUDA(cudaGetLastError());
    return IFB_OK;
}

// Shared address at which the kernel's dynamic shared memory starts (per device; the kernel reports it itself).
int probe_smem_base(int device, uint32_t *out) {
    static std::mutex mu;
    static std::map<int, uint32_t> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(device);
    if (it != cache.end()) {
        *out = it->second;
        return IFB_OK;
    }
    uint32_t *dp = nullptr;
    IFB_CUDA(cudaMalloc((void **)&dp, 4));
    static RankTopTable top0;   // zero-initialised
    RankParams p;
    std::memset(&p, 0, sizeof p);
    p.probe_out = dp;
    uint32_t v = 0, v2 = 0;
    int rc = launch_rank_variant<6>(top0, p, 1, 0);
    cudaError_t e = rc ? cudaSuccess : cudaMemcpy(&v, dp, 4, cudaMemcpyDeviceToHost);
    if (!rc && e == cudaSuccess) {
        rc = launch_rank_variant<-1>(top0, p, 1, 0);
        if (!rc) e = cudaMemcpy(&v2, dp, 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(dp);
    if (rc) return rc;
    IFB_CUDA(e);
    if (v != v2) v = 0xFFFFFFFFu;   // never seen; disqualifies the kernel instead of trusting a layout assumption
    cache[device] = v;
    *out = v;
    return IFB_OK;
}

struct FloatLess {   // -0.0 and +0.0 are one cut (x < -0.0 <=> x < +0.0)
    bool operator()(float a, float b) const { return a < b; }
};

}  // namespace

void free_rank_plans(ifb_forest *f) {
    for (auto *rp : f->rank_plans) {
        cudaFree(rp->d_w);
        cudaFree(rp->d_lv);
        cudaFree(rp->d_cut);
        delete rp;
    }
    f->rank_plans.clear();
}

// Builds (once per feature count) the rank tables of the forest; (*out)->ok == false when the forest or the matrix
// shape does not qualify (the caller then uses score_std.cu).
int get_rank_plan(ifb_forest *f, int32_t d, RankPlan **out) {
    std::lock_guard<std::mutex> lk(f->plan_mu);
    for (auto *rp : f->rank_plans)
        if (rp->d == d) {
            *out = rp;
            return IFB_OK;
        }
    auto *rp = new RankPlan();
    rp->d = d;
    f->rank_plans.push_back(rp);
    *out = rp;
    const int T = f->num_trees;
    if (d > 32 || T == 0 || f->max_feature_index >= d || device_smem_optin(f->device) < (int)kDynSmem) return IFB_OK;
    for (int t = 0; t < T; t++)
        if (f->bfs_off[t + 1] - f->bfs_off[t] > (int)kBlockWords - 1) return IFB_OK;
    for (size_t g = 0; g < f->h_val.size(); g++)
        if (f->h_child[g] >= 0 && !std::isfinite(f->h_val[g])) return IFB_OK;   // +inf / NaN thresholds: score_std.cu
    DeviceGuard dg(f->device);
    uint32_t sbase = 0;
    int rc = probe_smem_base(f->device, &sbase);
    if (rc) return rc;
    rp->sbase = sbase;
    if (sbase > 4096u) return IFB_OK;
    const uint32_t a_w = (sbase + 64u + kBlockBytes - 1u) & ~(kBlockBytes - 1u);
    if (a_w + kBlockBytes + kGridBytes > kTileAbs || sbase + kDynSmem < kHighAbs + 4096u) return IFB_OK;
    const int max_blocks = (int)((kTileAbs - kGridBytes - a_w) / kBlockBytes);
    const int64_t high_bytes = (int64_t)sbase + kDynSmem - kHighAbs;

    std::vector<uint32_t> W;
    std::vector<float> LV, CUT;
    int t = 0;
    while (t < T) {
        RankChunk c;
        c.tree_begin = t;
        std::vector<int> block_fill;                 // words used per block
        std::vector<std::pair<int, int>> place;      // per tree of the chunk: (block, first word)
        std::set<float, FloatLess> cuts[32];
        std::map<uint32_t, int> lv_index;            // leaf value bits -> payload
        int64_t cut_total = 0;
        while (t < T && (t - c.tree_begin) < kMaxTrees) {
            const int32_t base = f->bfs_off[t], n = f->bfs_off[t + 1] - base;
            const int need = (n + 2) & ~1;           // pad word + nodes, even
            // tentative: new cuts / leaf values this tree adds
            std::set<float, FloatLess> add_c[32];
            std::set<uint32_t> add_l;
            for (int q = 0; q < n; q++) {
                const int64_t gq = base + q;
                if (f->h_child[gq] >= 0) {
                    const int ft = (int)f->h_meta_feat[gq];
                    if (!cuts[ft].count(f->h_val[gq])) add_c[ft].insert(f->h_val[gq]);
                } else {
                    uint32_t b;
                    std::memcpy(&b, &f->h_val[gq], 4);
                    if (!lv_index.count(b)) add_l.insert(b);
                }
            }
            int64_t new_cut_total = cut_total;
            bool fits = true;
            for (int ft = 0; ft < 32; ft++) {
                new_cut_total += (int64_t)add_c[ft].size();
                if ((int)(cuts[ft].size() + add_c[ft].size()) > kMaxCutsPerFeature) fits = false;
            }
            const int64_t new_lv = (int64_t)lv_index.size() + (int64_t)add_l.size();
            if (new_lv > kMaxLeafValues) fits = false;
            if ((new_cut_total + 32 + new_lv) * 4 > high_bytes) fits = false;
            int bi = -1;                             // best fit among the open blocks
            for (int b = 0; b < (int)block_fill.size(); b++)
                if (block_fill[b] + need <= (int)kBlockWords && (bi < 0 || block_fill[b] > block_fill[bi])) bi = b;
            if (bi < 0 && (int)block_fill.size() >= max_blocks) fits = false;
            if (!fits) break;
            if (bi < 0) {
                block_fill.push_back(0);
                bi = (int)block_fill.size() - 1;
            }
            place.push_back({bi, block_fill[bi]});
            block_fill[bi] += need;
            for (int ft = 0; ft < 32; ft++) cuts[ft].insert(add_c[ft].begin(), add_c[ft].end());
            for (uint32_t b : add_l) {
                const int id = (int)lv_index.size();
                lv_index[b] = id;
            }
            cut_total = new_cut_total;
            t++;
        }
        if (t == c.tree_begin) return IFB_OK;   // one tree alone does not fit: not a shape for this kernel
        c.tree_end = t;
        const int nb = (int)block_fill.size();
        c.a_w = a_w;
        c.a_grid = a_w + (uint32_t)nb * kBlockBytes;
        c.a_cut = kHighAbs;
        c.w_off = (int64_t)W.size();
        c.w_words = nb * (int)kBlockWords;
        c.lv_off = (int64_t)LV.size();
        c.n_lv = (int)lv_index.size();
        c.cut_off_words = (int64_t)CUT.size();
        // cuts: per feature ascending + sentinel; grid parameters
        std::vector<std::vector<float>> cv(32);
        uint32_t off = 0;
        for (int ft = 0; ft < 32; ft++) {
            cv[ft].assign(cuts[ft].begin(), cuts[ft].end());
            c.cut_off[ft] = off * 4u;
            c.n_cut[ft] = (uint32_t)cv[ft].size();
            for (float v : cv[ft]) CUT.push_back(v);
            CUT.push_back(INFINITY);
            off += (uint32_t)cv[ft].size() + 1u;
            float inv = 0.f, c0 = 1.f;
            if (cv[ft].size() >= 2) {
                const float lo = cv[ft].front(), hi = cv[ft].back();
                const float i2 = 254.f / (hi - lo);
                const float c2 = 1.f - lo * i2;
                if (std::isfinite(i2) && std::isfinite(c2) && i2 > 0.f) {
                    inv = i2;
                    c0 = c2;
                }
            }
            c.inv[ft] = inv;
            c.c0[ft] = c0;
        }
        c.cut_words = (int)off;
        c.a_lv = c.a_cut + off * 4u;
        LV.resize(LV.size() + lv_index.size());
        for (auto &kv : lv_index) std::memcpy(&LV[(size_t)c.lv_off + kv.second], &kv.first, 4);
        // node words
        W.resize(W.size() + (size_t)c.w_words, 0u);
        uint32_t *cw = W.data() + c.w_off;
        std::memset(&c.top, 0, sizeof c.top);
        for (int tt = c.tree_begin; tt < c.tree_end; tt++) {
            const int32_t base = f->bfs_off[tt], n = f->bfs_off[tt + 1] - base;
            const int blk = place[tt - c.tree_begin].first, w0 = place[tt - c.tree_begin].second;
            // BFS node q lives at block word w0 + 1 + q: the root in a right slot, every child pair 8-byte aligned
            auto word_of = [&](int q) { return w0 + 1 + q; };
            for (int q = 0; q < n; q++) {
                const int64_t gq = base + q;
                const int self = word_of(q);
                uint32_t w;
                if (f->h_child[gq] >= 0) {
                    const int ft = (int)f->h_meta_feat[gq];
                    const auto &v = cv[ft];
                    const uint32_t j = (uint32_t)(std::lower_bound(v.begin(), v.end(), f->h_val[gq], FloatLess()) - v.begin());
                    const uint32_t pair = (uint32_t)word_of(f->h_child[gq]) * 4u;   // even word => bit 2 clear
                    w = ((kCodeBase + j + 1u) << 16) | ((uint32_t)ft * kColBytes) | pair;
                } else {
                    uint32_t b;
                    std::memcpy(&b, &f->h_val[gq], 4);
                    const uint32_t payload = (uint32_t)lv_index[b];
                    const uint32_t pair = (uint32_t)(self & ~1) * 4u;
                    w = (((self & 1) ? payload : (kLeftLeaf | payload)) << 16) | pair;
                }
                cw[(size_t)blk * kBlockWords + self] = w;
            }
            RankTop &e = c.top.e[tt - c.tree_begin];
            const uint32_t blk_abs = a_w + (uint32_t)blk * kBlockBytes;
            const uint32_t wr = cw[(size_t)blk * kBlockWords + word_of(0)];
            e.w0 = wr;
            e.f0 = wr & kFeatMask;
            if (f->h_child[base] >= 0) {
                const int l = word_of(f->h_child[base]);
                e.wL = cw[(size_t)blk * kBlockWords + l];
                e.wR = cw[(size_t)blk * kBlockWords + l + 1];
            } else {
                e.wL = e.wR = wr;   // a root leaf sits in a right slot: "not less" keeps it on itself
            }
            e.aL = blk_abs + (e.wL & (kBlockBytes - 1u));
            e.aR = blk_abs + (e.wR & (kBlockBytes - 1u));
        }
        rp->chunks.push_back(c);
    }
    IFB_CUDA(cudaMalloc((void **)&rp->d_w, std::max<size_t>(16, W.size() * 4)));
    IFB_CUDA(cudaMalloc((void **)&rp->d_lv, std::max<size_t>(16, LV.size() * 4)));
    IFB_CUDA(cudaMalloc((void **)&rp->d_cut, std::max<size_t>(16, CUT.size() * 4)));
    IFB_CUDA(cudaMemcpy(rp->d_w, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    IFB_CUDA(cudaMemcpy(rp->d_lv, LV.data(), LV.size() * 4, cudaMemcpyHostToDevice));
    IFB_CUDA(cudaMemcpy(rp->d_cut, CUT.data(), CUT.size() * 4, cudaMemcpyHostToDevice));
    f->device_bytes += (int64_t)((W.size() + LV.size() + CUT.size()) * 4);
    rp->ok = true;
    return IFB_OK;
}

int rank_plan_chunks(const RankPlan *rp) { return rp && rp->ok ? (int)rp->chunks.size() : 0; }

int launch_score_standard_rank(const ifb_forest *f, RankPlan *rp, const float *X, int64_t n_rows, int32_t d, int64_t ld,
                               double *scores, float *path_sum, bool accumulate_only, cudaStream_t stream,
                               const ScatterTarget *scatter) {
    if (n_rows == 0) return IFB_OK;
    const size_t n_chunks = rp->chunks.size();
    IFB_REQUIRE(n_chunks <= 1 || path_sum != nullptr || accumulate_only,
                "internal: multi-chunk scoring needs a path_sum scratch buffer");
    const int sms = device_sm_count(f->device);
    const int64_t n_gtiles = (n_rows + kG - 1) / kG;
    const int grid = (int)std::min<int64_t>((n_gtiles + kNG - 1) / kNG, sms);
    for (size_t ci = 0; ci < n_chunks; ci++) {
        const RankChunk &c = rp->chunks[ci];
        RankParams p;
        std::memset(&p, 0, sizeof p);
        p.X = X;
        p.n_rows = n_rows;
        p.ld = ld;
        p.d = d;
        p.use_bulk = ((reinterpret_cast<uintptr_t>(X) & 15u) == 0 && ld % 4 == 0) ? 1 : 0;
        p.gw = rp->d_w + c.w_off;
        p.glv = rp->d_lv + c.lv_off;
        p.gcut = rp->d_cut + c.cut_off_words;
        p.w_words = c.w_words;
        p.n_lv = c.n_lv;
        p.cut_words = c.cut_words;
        p.a_w = c.a_w;
        p.a_lv = c.a_lv;
        p.a_grid = c.a_grid;
        p.a_cut = c.a_cut;
        p.sbase = rp->sbase;
        std::memcpy(p.cut_off, c.cut_off, sizeof p.cut_off);
        std::memcpy(p.n_cut, c.n_cut, sizeof p.n_cut);
        std::memcpy(p.inv, c.inv, sizeof p.inv);
        std::memcpy(p.c0, c.c0, sizeof p.c0);
        p.n_trees = c.tree_end - c.tree_begin;
        p.max_depth = f->max_depth;
        p.total_trees = f->num_trees;
        p.avg_path = f->avg_path_norm;
        p.first_chunk = (ci == 0 && !accumulate_only) ? 1 : 0;
        p.finalize = (ci + 1 == n_chunks && !accumulate_only && !scatter) ? 1 : 0;
        p.finalize_scatter = (scatter && ci + 1 == n_chunks) ? 1 : 0;
        p.scores = scores;
        p.path_sum = path_sum;
        p.n_gtiles = n_gtiles;
        if (scatter) p.scatter = *scatter; else p.scatter.world = 0;
        p.probe_out = nullptr;
        int rc = f->max_depth == 8 ? launch_rank_variant<6>(c.top, p, grid, stream)
                                   : launch_rank_variant<-1>(c.top, p, grid, stream);
        if (rc) return rc;
        count_launch();
    }
    return IFB_OK;
}

}  // namespace ifb
