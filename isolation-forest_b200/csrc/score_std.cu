// Standard isolation-forest scoring for sm_100a.
//
// Replaces the per-row UDF of IsolationForestModel.transform (IF/IsolationForestModel.scala:131-139) and
// the tail-recursive walk IsolationTree.pathLength (IF/IsolationTree.scala:196-230).
//
// Shape of the kernel (DESIGN.md "score_standard"):
//   * persistent grid, one CTA per SM, R threads; thread r owns row r of the current row tile;
//   * the node tables of one forest chunk (val[] / meta[] words, see ifb_internal.h) are copied into
//     shared memory once per CTA;
//   * row tiles are filled by TMA (cp.async.bulk.tensor.2d on a tensor map over the column-major matrix;
//     box = RB rows x <=256 features, landing as [feature][RB] so that lane r reads bank r%32 whatever
//     feature its node asks for) and signalled through mbarriers.  Wide single-stage tiles (512 / 1024 rows)
//     are split into 256-row groups that wait, walk and re-arm their own sub-tile independently, so one
//     group's fill overlaps the other groups' walks; narrow tiles use a 2-stage ring or whole-tile refills
//     (the planner in forest.cu picks tile width and depth);
//   * levels 0 and 1 of every tree come from a kernel-parameter table (constant bank, warp-uniform);
//   * a walk is a fixed number of steps (the forest's max depth); leaves map onto themselves through the
//     NaN pseudo feature, so there is no per-level branch; C trees are walked at once per thread for ILP;
//   * per-row path lengths are added in tree order in f32 (Array[Float].sum), then
//     score = exp2(-(s/T)/c(numSamples)) in f64.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "ifb_internal.h"

namespace ifb {

// Rows per TMA box (== row stride of a shared-memory sub-tile): 256, or the whole tile when it is narrower.
// Single-stage tiles wider than one box are split into row GROUPS of 256 threads that load, wait and refill
// independently (see kGroupPipe in the kernel): 1024 -> 4 x 256, 512 -> 2 x 256.  Splitting 256-row tiles into
// 2 x 128 was measured and LOSES 17 % (d = 128: 128-row boxes halve the contiguous run of every TMA row), so tiles of
// <= 256 rows keep one box and whole-tile refills.  IFB_STD_NO_GROUPS=1 (A/B hook) restores whole-tile refills.
bool std_rank_enabled() {
    const char *e = getenv("IFB_STD_RANK");   // read on every call: tests switch it inside one process
    return e != nullptr && e[0] == '1';
}
bool std_grouped() {
    static const bool v = getenv("IFB_STD_NO_GROUPS") == nullptr;
    return v;
}
__host__ __device__ constexpr int std_rows_per_box_c(int R, int stages, bool grouped) {
    (void)stages;
    (void)grouped;
    return R < 256 ? R : 256;
}
int std_rows_per_box(int R, int stages) { return std_rows_per_box_c(R, stages, std_grouped()); }

namespace {

constexpr int kMaxStages = 2;
constexpr int kMaxTopTrees = 256;   // trees per chunk whose two top levels ride in the kernel parameters

// Levels 0 and 1 of every tree of the chunk, passed as a kernel parameter (constant bank): all 32 lanes of a
// warp walk the same tree, so these loads are warp-uniform and cost no shared-memory wavefronts.
struct TopEntry {
    float thr0;          // root: val word
    uint32_t f0;         // root: byte offset of its feature column inside a row sub-tile (feature * RB * 4)
    uint32_t c0;         // root: byte offset of its left child in val[]
    float thrL, thrR;    // the two level-1 candidates (nodes at c0 and c0 + 4)
    uint32_t fL, fR;
    uint32_t cL, cR;
    uint32_t pad[3];
};
struct TopTable {
    TopEntry e[kMaxTopTrees];
};

struct ScoreStdParams {
    const float *X;          // column-major matrix (fallback loader) -- also the TMA source
    int64_t n_rows;
    int64_t ld;
    int32_t d;
    int32_t box_d;           // features per TMA box (<= 256)
    int32_t rem_d;           // features in the trailing partial box (0 = none)
    const float *val;        // chunk tables in global memory
    const uint32_t *meta;
    int32_t chunk_words;
    int32_t n_trees;         // trees in this chunk
    int32_t max_depth;
    int32_t total_trees;     // ensemble size (divisor of the mean)
    float avg_path;          // c(numSamples)
    int32_t first_chunk;     // start sums at 0 instead of reading path_sum / depth_sum
    int32_t finalize;        // write scores
    int32_t finalize_scatter;  // last chunk of a scatter launch
    double *scores;
    float *path_sum;         // may be null when first_chunk && finalize
    int32_t *depth_sum;      // may be null
    int64_t n_tiles;
    ScatterTarget scatter;   // tree-sharded multi-GPU: where the finished per-row sums of the LAST chunk go
    int32_t l2_prefetch;     // single-stage whole-tile refills: prefetch the next tile into L2 during the walk
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Bounded wait: a TMA that never lands (bad descriptor, lost transaction) must surface as a launch failure,
// not as a hung GPU.  ~2^31 polls of a few ns each is tens of seconds, far beyond any legitimate wait.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
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
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int32_t c0, int32_t c1,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], "
        "[%4];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// L2 prefetch of a box (no shared-memory destination, no barrier): used by single-stage tiles that cannot overlap
// their refill with the walk -- the next tile is pulled into L2 while this one is walked, so the refill itself runs
// at L2 speed instead of waiting on HBM.
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1)
                 : "memory");
}

// Shared-memory carve-up (byte offsets from the dynamic smem base, which is 1024-aligned).
struct SmemLayout {
    uint32_t bars;    // kStages mbarriers
    uint32_t roots;   // n_trees u32
    uint32_t val;     // chunk_words f32
    uint32_t meta;    // chunk_words u32
    uint32_t tiles;   // kStages * R * (d+1) f32, 128-byte aligned
    uint32_t total;
};
__host__ __device__ inline SmemLayout make_layout(int n_trees, int chunk_words, int R, int d, int kStages) {
    SmemLayout L;
    L.bars = 0;
    L.roots = 64;
    L.val = 64;
    L.meta = L.val + (uint32_t)chunk_words * 4;
    L.tiles = (L.meta + (uint32_t)chunk_words * 4 + 127u) & ~127u;
    L.total = L.tiles + (uint32_t)kStages * (uint32_t)R * (uint32_t)(d + 1) * 4u;
    return L;
}

// One generic level of one walk.  PTX pins the instruction mix: 3 LDS + {LOP3, SHF} on the ALU pipe +
// {IMAD, IMAD} on the FMA pipe + FSET, so that neither math pipe becomes the limiter.
//   meta = (left-child WORD index << 16) | feature;   next = child*4 + 4*!(x < val)
__device__ __forceinline__ uint32_t walk_step(uint32_t node, uint32_t val_s, uint32_t meta_s, uint32_t xrow_s,
                                              uint32_t col_bytes, uint32_t d, int32_t &depth, bool want_depth) {
    float v, x;
    uint32_t m, feat, xaddr, cb, ge, next;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(val_s + node));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m) : "r"(meta_s + node));
    asm("and.b32 %0, %1, 0xFFFF;" : "=r"(feat) : "r"(m));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(xaddr) : "r"(feat), "r"(col_bytes), "r"(xrow_s));
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(xaddr));
    asm("shr.u32 %0, %1, 14;" : "=r"(cb) : "r"(m));            // (child << 2) | (feat >> 14) == child*4
    asm("set.geu.u32.f32 %0, %1, %2;" : "=r"(ge) : "f"(x), "f"(v));   // 0xFFFFFFFF when !(x < v) (incl. NaN)
    asm("mad.lo.s32 %0, %1, -4, %2;" : "=r"(next) : "r"(ge), "r"(cb));
    if (want_depth) depth += (feat != d) ? 1 : 0;
    return next;
}

template <int R, int C, bool USE_TMA, bool WANT_DEPTH, int DEEP, int kStages, bool GROUPED>
__global__ void __launch_bounds__(R, 1)
score_std_kernel(const __grid_constant__ CUtensorMap tmap_main, const __grid_constant__ CUtensorMap tmap_rem,
                 const __grid_constant__ TopTable top, const ScoreStdParams p) {
    constexpr int RB = std_rows_per_box_c(R, kStages, GROUPED);  // rows per TMA box == row stride of the smem tile
    constexpr int NSUB = R / RB;
    // Group pipelining (single-stage TMA tiles): every row group of RB threads owns its sub-tile, its mbarrier and a
    // named barrier; a group refills its sub-tile the moment ITS rows are done, so the groups drift out of phase and
    // one group's TMA fill hides behind the other groups' walks -- double buffering without a second buffer.
    constexpr bool kGroupPipe = GROUPED && USE_TMA && kStages == 1 && NSUB > 1;
    constexpr int NBARS = kGroupPipe ? NSUB : kStages;
    constexpr uint32_t COL_BYTES = RB * 4;
    extern __shared__ __align__(1024) unsigned char smem[];
    const SmemLayout L = make_layout(p.n_trees, p.chunk_words, R, p.d, kStages);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
    const int tid = threadIdx.x;
    const int d = p.d;
    const uint32_t tile_floats = (uint32_t)R * (uint32_t)(d + 1);
    const uint32_t sub_floats = (uint32_t)RB * (uint32_t)(d + 1);

    // ---- one-time setup: barriers, forest chunk, NaN pseudo columns -------------------------------
    if (tid == 0) {
        for (int s = 0; s < NBARS; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        float *sval = reinterpret_cast<float *>(smem + L.val);
        uint32_t *smeta = reinterpret_cast<uint32_t *>(smem + L.meta);
        for (int i = tid; i < p.chunk_words; i += R) {
            sval[i] = p.val[i];
            smeta[i] = p.meta[i];
        }
        float *tiles = reinterpret_cast<float *>(smem + L.tiles);
        const float qnan = __int_as_float(0x7fc00000);
        for (int i = tid; i < kStages * NSUB * RB; i += R) {
            const int s = i / (NSUB * RB), rem = i % (NSUB * RB);
            const int sub = rem / RB, r = rem % RB;
            tiles[(uint32_t)s * tile_floats + (uint32_t)sub * sub_floats + (uint32_t)d * RB + r] = qnan;
        }
    }
    __syncthreads();

    const uint32_t tile_bytes = (uint32_t)R * (uint32_t)d * 4u;  // bytes one stage receives from TMA
    auto issue_tile = [&](int64_t tile, int stage) {
        // called by thread 0 only (TMA path)
        float *dst = reinterpret_cast<float *>(smem + L.tiles) + (uint32_t)stage * tile_floats;
        mbar_expect_tx(&bars[stage], tile_bytes);
        const int64_t row0 = tile * R;
#pragma unroll
        for (int sub = 0; sub < NSUB; sub++) {
            float *sdst = dst + (uint32_t)sub * sub_floats;
            const int32_t r0 = (int32_t)(row0 + (int64_t)sub * RB);
            int f = 0;
            for (; f + p.box_d <= d; f += p.box_d) tma_load_2d(sdst + (uint32_t)f * RB, &tmap_main, r0, f, &bars[stage]);
            if (p.rem_d) tma_load_2d(sdst + (uint32_t)f * RB, &tmap_rem, r0, f, &bars[stage]);
        }
    };
    auto load_tile_plain = [&](int64_t tile, int stage) {
        // all threads; column-major source: lanes run along rows => coalesced
        float *dst = reinterpret_cast<float *>(smem + L.tiles) + (uint32_t)stage * tile_floats;
        const int64_t row0 = tile * R;
        const int sub = tid / RB, r = tid % RB;
        const int64_t row = row0 + tid;
        float *sdst = dst + (uint32_t)sub * sub_floats + r;
        if (row < p.n_rows) {
            const float *src = p.X + row;
#pragma unroll 4
            for (int f = 0; f < d; f++) sdst[(uint32_t)f * RB] = __ldg(src + (int64_t)f * p.ld);
        } else {
            for (int f = 0; f < d; f++) sdst[(uint32_t)f * RB] = 0.f;
        }
    };

    auto prefetch_tile = [&](int64_t tile) {
        // thread 0 only: the boxes issue_tile(tile, .) will fetch, pulled into L2 ahead of time
        const int64_t row0 = tile * R;
#pragma unroll
        for (int sub = 0; sub < NSUB; sub++) {
            const int32_t r0 = (int32_t)(row0 + (int64_t)sub * RB);
            int f = 0;
            for (; f + p.box_d <= d; f += p.box_d) tma_prefetch_2d(&tmap_main, r0, f);
            if (p.rem_d) tma_prefetch_2d(&tmap_rem, r0, f);
        }
    };

    auto issue_group = [&](int64_t tile, int sub) {
        // called by the first thread of row group `sub` (kGroupPipe): refill this group's sub-tile only
        float *sdst = reinterpret_cast<float *>(smem + L.tiles) + (uint32_t)sub * sub_floats;
        mbar_expect_tx(&bars[sub], (uint32_t)RB * (uint32_t)d * 4u);
        const int32_t r0 = (int32_t)(tile * R + (int64_t)sub * RB);
        int f = 0;
        for (; f + p.box_d <= d; f += p.box_d) tma_load_2d(sdst + (uint32_t)f * RB, &tmap_main, r0, f, &bars[sub]);
        if (p.rem_d) tma_load_2d(sdst + (uint32_t)f * RB, &tmap_rem, r0, f, &bars[sub]);
    };

    int64_t tile = blockIdx.x;
    const int64_t stride = gridDim.x;
    if constexpr (kGroupPipe) {
        if (tid % RB == 0 && tile < p.n_tiles) issue_group(tile, tid / RB);
    } else if constexpr (USE_TMA) {
        if (tid == 0) {
            if (tile < p.n_tiles) issue_tile(tile, 0);
            if (kStages > 1 && tile + stride < p.n_tiles) issue_tile(tile + stride, 1);
        }
    }

    const uint32_t val_s = smem_u32(smem + L.val), meta_s = smem_u32(smem + L.meta);
    const uint32_t tiles_s = smem_u32(smem + L.tiles);
    const int sub = tid / RB, rl = tid % RB;
    const int n_trees = p.n_trees;
    // levels 0 and 1 come from `top`; DEEP >= 0 fixes the remaining level count at compile time (full unroll)
    const int deep_levels = DEEP >= 0 ? DEEP : (p.max_depth > 2 ? p.max_depth - 2 : 0);
    const uint32_t leaf_col = (uint32_t)d * COL_BYTES;

    for (int64_t k = 0; tile < p.n_tiles; tile += stride, ++k) {
        const int stage = kStages > 1 ? (int)(k & 1) : 0;
        if constexpr (kGroupPipe) {
            mbar_wait(&bars[sub], (uint32_t)(k & 1));
        } else if constexpr (USE_TMA) {
            mbar_wait(&bars[stage], kStages > 1 ? (uint32_t)((k >> 1) & 1) : (uint32_t)(k & 1));
        } else {
            load_tile_plain(tile, stage);
            __syncthreads();
        }
        if constexpr (USE_TMA && kStages == 1 && !kGroupPipe) {
            if (p.l2_prefetch && tid == 0 && tile + stride < p.n_tiles) prefetch_tile(tile + stride);
        }
        const uint32_t xrow_s = tiles_s + ((uint32_t)stage * tile_floats + (uint32_t)sub * sub_floats + (uint32_t)rl) * 4u;
        const int64_t row = tile * R + tid;
        const bool live = row < p.n_rows;

        float s = 0.f;
        int32_t dsum = 0;
        if (!p.first_chunk && live) {
            s = p.path_sum[row];
            if (WANT_DEPTH) dsum = p.depth_sum[row];
        }

        // levels 0 and 1 from the constant bank, then `deep_levels` generic steps, C trees in flight
        auto top_levels = [&](int t) -> uint32_t {
            const TopEntry &e = top.e[t];
            float x0, x1;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x0) : "r"(xrow_s + e.f0));
            const bool lt0 = x0 < e.thr0;
            const float thr1 = lt0 ? e.thrL : e.thrR;
            const uint32_t f1 = lt0 ? e.fL : e.fR;
            const uint32_t c1 = lt0 ? e.cL : e.cR;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x1) : "r"(xrow_s + f1));
            if (WANT_DEPTH) dsum += (e.f0 != leaf_col ? 1 : 0) + (f1 != leaf_col ? 1 : 0);
            return c1 + ((x1 < thr1) ? 0u : 4u);
        };

        // walk CC trees at once (independent dependency chains), add their path lengths in tree order
        auto walk_group = [&](int t0, auto cc_tag) {
            constexpr int CC = decltype(cc_tag)::value;
            uint32_t node[CC];
#pragma unroll
            for (int c = 0; c < CC; c++) node[c] = top_levels(t0 + c);
            if constexpr (DEEP >= 0) {
#pragma unroll
                for (int lvl = 0; lvl < DEEP; lvl++) {
#pragma unroll
                    for (int c = 0; c < CC; c++)
                        node[c] = walk_step(node[c], val_s, meta_s, xrow_s, COL_BYTES, (uint32_t)d, dsum, WANT_DEPTH);
                }
            } else {
#pragma unroll 1
                for (int lvl = 0; lvl < deep_levels; lvl++) {
#pragma unroll
                    for (int c = 0; c < CC; c++)
                        node[c] = walk_step(node[c], val_s, meta_s, xrow_s, COL_BYTES, (uint32_t)d, dsum, WANT_DEPTH);
                }
            }
#pragma unroll
            for (int c = 0; c < CC; c++) {
                float lv;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lv) : "r"(val_s + node[c]));
                s = s + lv;
            }
        };
        int t = 0;
        for (; t + C <= n_trees; t += C) walk_group(t, std::integral_constant<int, C>{});
        if constexpr (C >= 16) if (t + 8 <= n_trees) { walk_group(t, std::integral_constant<int, 8>{}); t += 8; }
        if constexpr (C >= 8) if (t + 4 <= n_trees) { walk_group(t, std::integral_constant<int, 4>{}); t += 4; }
        if constexpr (C >= 4) if (t + 2 <= n_trees) { walk_group(t, std::integral_constant<int, 2>{}); t += 2; }
        for (; t < n_trees; t++) walk_group(t, std::integral_constant<int, 1>{});

        if (live) {
            if (p.finalize) {
                // IF/IsolationForestModel.scala:137-138: Float sum / Int, -Float / Float, Math.pow(2, Double)
                const float e = __fdiv_rn(s, (float)p.total_trees);
                const float z = __fdiv_rn(-e, p.avg_path);
                p.scores[row] = exp2((double)z);
            }
            if (p.scatter.world > 0 && p.finalize_scatter) {
                // fused reduce-scatter: this rank's partial sum of row `row` lands in the owning rank's buffer
                // (plain coalesced stores to NVLink peer memory; slot [rank][row - first row of the owner])
                int o = 0;
#pragma unroll
                for (int q = 1; q < kMaxScatterRanks; q++) o += (q < p.scatter.world && row >= p.scatter.cut[q]) ? 1 : 0;
                const int64_t r0 = p.scatter.cut[o], rows_o = p.scatter.cut[o + 1] - r0;
                p.scatter.peer[o][(int64_t)p.scatter.rank * rows_o + (row - r0)] = s;
            } else if (p.path_sum) {
                p.path_sum[row] = s;
            }
            if (WANT_DEPTH) p.depth_sum[row] = dsum;
        }
        if constexpr (kGroupPipe) {
            // every read of THIS GROUP's sub-tile is done before it is refilled; other groups are not waited for
            asm volatile("bar.sync %0, %1;" ::"r"(sub + 1), "n"(RB) : "memory");
            if (rl == 0 && tile + stride < p.n_tiles) issue_group(tile + stride, sub);
        } else {
            __syncthreads();  // every read of this stage is done before it is refilled
            if constexpr (USE_TMA) {
                if (tid == 0 && tile + kStages * stride < p.n_tiles) issue_tile(tile + kStages * stride, stage);
            }
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

int make_tmap(CUtensorMap *map, const float *X, int64_t n_rows, int32_t d, int64_t ld, int box_rows, int box_d) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return IFB_ECUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)n_rows, (cuuint64_t)d};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_d};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(X), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (n_rows=%lld d=%d ld=%lld box=%dx%d)", (int)r,
                  (long long)n_rows, d, (long long)ld, box_rows, box_d);
        return IFB_ECUDA;
    }
    return IFB_OK;
}

template <int R, int C, int S>
int launch_variant(bool use_tma, bool want_depth, const CUtensorMap &m0, const CUtensorMap &m1, const TopTable &top,
                   const ScoreStdParams &p, int grid, size_t smem, cudaStream_t stream) {
    auto go = [&](auto kern) -> int {
        IFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, R, smem, stream>>>(m0, m1, top, p);
        IFB_CUDA(cudaGetLastError());
        count_launch();
        return IFB_OK;
    };
    const bool deep6 = p.max_depth == 8 && !want_depth && use_tma;
    auto pick = [&](auto g_tag) -> int {
        constexpr bool G = decltype(g_tag)::value;
        if (deep6) return go(score_std_kernel<R, C, true, false, 6, S, G>);
        if (use_tma)
            return want_depth ? go(score_std_kernel<R, C, true, true, -1, S, G>)
                              : go(score_std_kernel<R, C, true, false, -1, S, G>);
        return want_depth ? go(score_std_kernel<R, C, false, true, -1, S, G>)
                          : go(score_std_kernel<R, C, false, false, -1, S, G>);
    };
    if constexpr (S == 1) {
        if (std_grouped()) return pick(std::true_type{});
    }
    return pick(std::false_type{});
}

}  // namespace

// Generic fallback: one thread per row, node tables and features straight from global memory (L2).  Used only when
// a row tile of the matrix cannot share shared memory with even one tree (d beyond ~850 features).
namespace {
__global__ void score_std_generic_kernel(const float *__restrict__ X, int64_t n_rows, int64_t ld, int layout,
                                         const float *__restrict__ val, const int32_t *__restrict__ feat,
                                         const int32_t *__restrict__ child, const int32_t *__restrict__ root, int num_trees,
                                         int total_trees, float avg_path, int accumulate_only, double *scores,
                                         float *path_sum, int32_t *depth_sum) {
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < n_rows; row += (int64_t)gridDim.x * blockDim.x) {
        float s = accumulate_only ? path_sum[row] : 0.f;
        int32_t dsum = (accumulate_only && depth_sum) ? depth_sum[row] : 0;
        const int64_t rs = layout == IFB_COL_MAJOR ? 1 : ld, cs = layout == IFB_COL_MAJOR ? ld : 1;
        for (int t = 0; t < num_trees; t++) {
            const int32_t base = root[t];
            int32_t node = 0;
            int32_t c = __ldg(child + base);
            while (c >= 0) {
                const float x = __ldg(X + row * rs + (int64_t)__ldg(feat + base + node) * cs);
                node = c + ((x < __ldg(val + base + node)) ? 0 : 1);
                c = __ldg(child + base + node);
                dsum++;
            }
            s = s + __ldg(val + base + node);
        }
        if (!accumulate_only) {
            const float e = __fdiv_rn(s, (float)total_trees);
            const float z = __fdiv_rn(-e, avg_path);
            scores[row] = exp2((double)z);
        }
        if (path_sum) path_sum[row] = s;
        if (depth_sum) depth_sum[row] = dsum;
    }
}
}  // namespace

int launch_score_standard_generic(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                                  double *scores, int32_t *depth_sum, float *path_sum, bool accumulate_only,
                                  cudaStream_t stream) {
    if (n_rows == 0) return IFB_OK;
    int rc = ensure_std_generic_tables(const_cast<ifb_forest *>(f));
    if (rc) return rc;
    const int grid = (int)std::min<int64_t>((n_rows + 127) / 128, (int64_t)device_sm_count(f->device) * 16);
    score_std_generic_kernel<<<grid, 128, 0, stream>>>(X, n_rows, ld, layout, f->d_gval, f->d_gfeat, f->d_gchild, f->d_groot,
                                                       f->num_trees, f->num_trees, f->avg_path_norm, accumulate_only ? 1 : 0,
                                                       scores, path_sum, depth_sum);
    IFB_CUDA(cudaGetLastError());
    count_launch();
    return IFB_OK;
}

size_t std_top_table_bytes() { return sizeof(TopTable); }
int std_top_table_max_trees() { return kMaxTopTrees; }

// Fill the kernel-parameter table of one chunk from its val/meta words (host side, called by the planner).
void std_fill_top_table(void *dst, const float *val, const uint32_t *meta, const uint32_t *root_byte, int n_trees,
                        int rows_per_box) {
    TopTable *tt = reinterpret_cast<TopTable *>(dst);
    std::memset(tt, 0, sizeof(TopTable));
    const uint32_t col = (uint32_t)rows_per_box * 4u;
    for (int t = 0; t < n_trees; t++) {
        TopEntry &e = tt->e[t];
        const uint32_t r = root_byte[t] / 4;
        e.thr0 = val[r];
        e.f0 = (meta[r] & 0xFFFFu) * col;
        e.c0 = (meta[r] >> 16) * 4u;
        const uint32_t l = e.c0 / 4, rr = l + 1;
        e.thrL = val[l];
        e.thrR = val[rr];
        e.fL = (meta[l] & 0xFFFFu) * col;
        e.fR = (meta[rr] & 0xFFFFu) * col;
        e.cL = (meta[l] >> 16) * 4u;
        e.cR = (meta[rr] >> 16) * 4u;
    }
}

int launch_score_standard(const ifb_forest *f, ifb_forest::StdPlan *plan, const float *X, int64_t n_rows, int32_t d,
                          int64_t ld, int32_t layout, double *scores, int32_t *depth_sum, float *path_sum,
                          bool accumulate_only, cudaStream_t stream, const ScatterTarget *scatter) {
    IFB_REQUIRE(layout == IFB_COL_MAJOR, "launch_score_standard expects a column-major matrix");
    if (n_rows == 0) return IFB_OK;
    // IFB_STD_RANK=1 (opt-in, measured alternative): narrow matrices walk on per-feature ranks (score_std_rank.cu).
    // Not the default: 2.53 ms against 2.48 ms for config 2 (DESIGN.md 4.1b has the ncu breakdown).
    if (!depth_sum && d <= 32 && std_rank_enabled()) {
        RankPlan *rp = nullptr;
        int rc = get_rank_plan(const_cast<ifb_forest *>(f), d, &rp);
        if (rc) return rc;
        const int nc = rank_plan_chunks(rp);
        if (nc == 1 || (nc > 1 && (path_sum != nullptr || accumulate_only)))
            return launch_score_standard_rank(f, rp, X, n_rows, d, ld, scores, path_sum, accumulate_only, stream, scatter);
    }
    const int R = plan->rows_per_tile;
    const int RB = std_rows_per_box(R, plan->stages);
    const bool want_depth = depth_sum != nullptr;
    const size_t n_chunks = plan->chunks.size();
    IFB_REQUIRE(n_chunks <= 1 || path_sum != nullptr || accumulate_only,
                "internal: multi-chunk scoring needs a path_sum scratch buffer");

    // TMA needs a 16-byte aligned base and a row pitch that is a multiple of 16 bytes
    bool use_tma = ((reinterpret_cast<uintptr_t>(X) & 15u) == 0) && (ld % 4 == 0) && n_rows < (1LL << 31);
    CUtensorMap m0, m1;
    std::memset(&m0, 0, sizeof m0);
    std::memset(&m1, 0, sizeof m1);
    const int box_d = std::min(d, 256);
    const int rem_d = d % box_d;
    if (use_tma) {
        int rc = make_tmap(&m0, X, n_rows, d, ld, RB, box_d);
        if (rc) return rc;
        if (rem_d) {
            rc = make_tmap(&m1, X, n_rows, d, ld, RB, rem_d);
            if (rc) return rc;
        } else {
            m1 = m0;
        }
    }
    const int sms = device_sm_count(f->device);
    const int64_t n_tiles = (n_rows + R - 1) / R;
    const int grid = (int)std::min<int64_t>(n_tiles, sms);

    for (size_t ci = 0; ci < n_chunks; ci++) {
        const StdChunk &c = plan->chunks[ci];
        ScoreStdParams p;
        p.X = X;
        p.n_rows = n_rows;
        p.ld = ld;
        p.d = d;
        p.box_d = box_d;
        p.rem_d = rem_d;
        p.val = plan->d_val + c.node_begin;
        p.meta = plan->d_meta + c.node_begin;
        p.chunk_words = c.node_count;
        p.n_trees = c.tree_end - c.tree_begin;
        p.max_depth = f->max_depth;
        p.total_trees = f->num_trees;
        p.avg_path = f->avg_path_norm;
        p.first_chunk = (ci == 0 && !accumulate_only) ? 1 : 0;
        p.finalize = (ci + 1 == n_chunks && !accumulate_only && !scatter) ? 1 : 0;
        p.scores = scores;
        p.path_sum = path_sum;
        p.depth_sum = depth_sum;
        p.n_tiles = n_tiles;
        if (scatter) p.scatter = *scatter; else p.scatter.world = 0;
        p.finalize_scatter = (scatter && ci + 1 == n_chunks) ? 1 : 0;
        static const bool no_prefetch = getenv("IFB_STD_NO_L2_PREFETCH") != nullptr;
        p.l2_prefetch = no_prefetch ? 0 : 1;
        const int S = plan->stages;
        const SmemLayout L = make_layout(p.n_trees, p.chunk_words, R, d, S);
        const TopTable &top = *reinterpret_cast<const TopTable *>(plan->h_top.data() + (size_t)ci * sizeof(TopTable));
        int rc;
        switch (R) {
            case 1024: {
                static const int c1024 = getenv("IFB_STD_1024") ? atoi(getenv("IFB_STD_1024")) : 4;
                rc = c1024 == 4 ? launch_variant<1024, 4, 1>(use_tma, want_depth, m0, m1, top, p, grid, L.total, stream)
                   : c1024 == 3 ? launch_variant<1024, 3, 1>(use_tma, want_depth, m0, m1, top, p, grid, L.total, stream)
                                : launch_variant<1024, 2, 1>(use_tma, want_depth, m0, m1, top, p, grid, L.total, stream);
                break;
            }
            case 512:
                rc = S == 2 ? launch_variant<512, 4, 2>(use_tma, want_depth, m0, m1, top, p, grid, L.total, stream)
                            : launch_variant<512, 4, 1>(use_tma, want_depth, m0, m1, top, p, grid, L.total, stream);
                break;
            case 256: {
                static const int c256 = getenv("IFB_STD_256") ? atoi(getenv("IFB_STD_256")) : 16;
                rc = S == 2     ? launch_variant<256, 8, 2>(use_tma, want_depth, m0, m1, top, p, grid, L.total, stream)
                   : c256 == 16 ? launch_variant<256, 16, 1>(use_tma, want_depth, m0, m1, top, p, grid, L.total, stream)
                                : launch_variant<256, 8, 1>(use_tma, want_depth, m0, m1, top, p, grid, L.total, stream);
                break;
            }
            case 128:
                rc = S == 2 ? launch_variant<128, 16, 2>(use_tma, want_depth, m0, m1, top, p, grid, L.total, stream)
                            : launch_variant<128, 16, 1>(use_tma, want_depth, m0, m1, top, p, grid, L.total, stream);
                break;
            case 64: rc = launch_variant<64, 16, 2>(use_tma, want_depth, m0, m1, top, p, grid, L.total, stream); break;
            default: rc = launch_variant<32, 16, 2>(use_tma, want_depth, m0, m1, top, p, grid, L.total, stream); break;
        }
        if (rc) return rc;
    }
    return IFB_OK;
}

}  // namespace ifb
