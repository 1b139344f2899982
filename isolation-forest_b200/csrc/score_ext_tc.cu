// Tensor-core scoring of extended isolation forests for sm_100a.
//
// Replaces ExtendedIsolationTree.pathLength (IF/extended/ExtendedIsolationTree.scala:283-355) and
// SplitHyperplane.dot (IF/extended/ExtendedUtils.scala:36-55): BASELINE configs 3 (d = 64) and 5 (d = 1024) with
// extensionLevel = d-1, and sparse hyperplanes (extensionLevel < d-1 or a feature subspace) as zero-padded columns.
// The per-row dot products of a walk are a contraction
//     S[row][node] = sum_i x[row][i] * w[node][i]
// so ALL hyperplanes of the forest are evaluated as one [rows x nodes x d] GEMM on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM), and the walk itself only compares accumulators with offsets.
//
// Exactness.  The reference decides `sum < offset` with sum = f64 sequential sum of f32-rounded products.  The tensor
// cores are only a FILTER: operands are split  x' = xh + xl,  w' = wh + wl  into fp16 pairs after an exact power-of-two
// scaling per row (||x'||_2 in [0.5, 1)) / per node (max |w'| in [0.5, 1)), S' = xh.wh + xl.wh + xh.wl is accumulated
// in f32 by three MMAs per k-step, and a visit is accepted only when
//     |S' - offset'| > c_k * ||w'||_2 * 1  >=  c_k * ||w'||_2 * ||x'||_2      (c_k: bound constant)
// which proves that the reference's comparison has the same outcome.  Every other visit ("stuck" lane, ~1e-5 of the
// visits at d = 64, ~1e-3 at d = 1024, and every visit of a row with non-finite / out-of-range features) is decided by
// the warp cooperatively with the reference's exact arithmetic on the STORED terms (f32 product, f64 sum; lane-parallel
// re-association with its own proven bound, else the sequential order).  Decisions are bit-identical to the reference's.
//
// Kernels
//   ext_tc_densify        sparse hyperplanes only: stored terms -> zero-filled rows of the matrix width (once per forest).
//   ext_tc_prepare_cols   one warp per hyperplane: scale, fp16 hi/lo split, ||w'||, offset' and bound coefficient
//                         (once per forest).
//   ext_tc_prepare_rows   per call: row scaling, fp16 hi/lo split, row norm, row-major f32 copy for the exact path.
//   score_ext_tc_kernel   persistent, warp-specialised, 640 threads, one CTA per SM in clusters of 2 or 4:
//     warp 0   one lane: TMA producer of the operand stages (A = rows, B = hyperplanes, 64-byte swizzled K-major tiles;
//              the B tiles are fetched once per cluster with TMA multicast);
//     warp 1   one lane: tcgen05.mma issuer (128 x 256 x 16 fp16, f32 accumulators, two 256-column TMEM buffers);
//     warps 2-17  epilogue, four per TMEM lane quarter (= per scheduler), working as two TEAMS on alternate blocks.
//              Drain, split by COLUMNS: each warp tcgen05.ld's its 32-column chunks and turns every accumulator into a
//              "left?" bit appended to per-lane mask words, with ONE chunk-wide ambiguity test (smallest |dlt| against
//              the largest bound); per-column "ambiguous?" bits only for chunks that fail it.  Walk, split by TREES, on
//              the masks (up to 4 chains per lane, leaves point at themselves); ambiguous visits are resolved exactly;
//              leaf values and depths go to shared memory;
//     warp 18  adds every block's leaf values in tree order (the reference's sequential f32 sum) and writes the results;
//     warp 19  one lane: feeds the ring of block descriptors (node tables), independent of the operand stages.
//   Measured alternatives kept as opt-ins: IFB_TC_CG=2, the two CTAs of a cluster as one cta_group::2
//   pair (one MMA for 2 x 128 rows, half of the hyperplane tile per CTA); IFB_TC_BK=16, six 24 KB stages of 32-byte rows.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "ifb_internal.h"

namespace ifb {

namespace tc {

constexpr int BM = 128;        // rows per tile (TMEM lanes)
constexpr int BN = 256;        // hyperplanes (accumulator columns) per block
constexpr int BK = 32;         // padding unit of K (the widest K chunk per stage, see Geo)
constexpr int MAX_TREES_PER_BLOCK = 16;
constexpr int MAX_LEAVES_PER_BLOCK = 256;   // node ids stay below 512: bit 8 alone tells a leaf from an internal node
#ifndef IFB_TC_PW
#define IFB_TC_PW 2
#endif
constexpr int PW = IFB_TC_PW;    // warps per team: a lane quarter (= scheduler) has two teams working on alternate blocks
constexpr int EPI_WARPS = 8 * PW;
constexpr int META_RING = 3;     // block descriptors in flight: the producer runs ahead of the epilogue by up to 3 blocks
constexpr int THREADS = (2 + EPI_WARPS + 2) * 32;   // producer, MMA issuer, 16 epilogue warps, summing warp, descriptor loader

// Node ids inside a block: [0, 256) = internal nodes (= accumulator columns), 256 + i = leaf i of the block.
constexpr uint32_t LEAF0 = BN;
constexpr int MAX_NODES_PER_BLOCK = BN + MAX_LEAVES_PER_BLOCK;
struct BlockMeta {
    float2 ne[BN];            // per column (= internal node): {-offset', bound}: offset' = offset * 2^-s_node rounded to
                              // f32; the visit is certain iff |S' - offset' * 2^-e_row| >= bound  (bound > c_k ||w'||)
    float nthr[BN];           // -offset' again, densely packed for the drain's fast path (+inf at padding columns)
    float emax[BN / 32];      // largest bound of each 32-column chunk
    uint2 next[MAX_NODES_PER_BLOCK];   // {left, right} node ids; a leaf points at itself, so a walk is a fixed number of
                                       // branch-free steps
    int32_t slot[BN];         // weight slot (row of d_ext_w) for the exact path; -1: padding column
    float leafv[MAX_LEAVES_PER_BLOCK];   // (float)depth + c(numInstances)
    uint8_t leafd[MAX_LEAVES_PER_BLOCK]; // depth of the leaf = internal nodes visited on the way (depth_sum output)
    uint16_t root[MAX_TREES_PER_BLOCK];
    int32_t n_trees;          // consecutive trees of the ensemble, columns assigned in tree order
    int32_t n_cols;           // used columns
    int32_t tree0;
    int32_t pad[5];
};
static_assert(sizeof(BlockMeta) % 16 == 0, "bulk copies need 16-byte multiples");
constexpr uint32_t META_BYTES = sizeof(BlockMeta);

// Operand stage geometry and shared-memory carve-up (offsets from a 1024-aligned base) for a K chunk of BKT elements per
// stage.  BKT = 32: 64-byte rows (64B swizzle), two k-steps per stage, 3 stages of 48 KB.  BKT = 16: 32-byte rows (32B
// swizzle), one k-step per stage, 6 stages of 24 KB -- the same bytes, but five stages instead of two in flight while one
// is consumed.  Built to test whether wide hyperplanes are bound by the LATENCY of the operand feed (the MMA issuer waits
// for operands 31 % of the time with 3 x 48 KB): they are not -- 400K x 1024 rows, 256 trees: 47.0 ms against 41.9 ms,
// operand waits 35 % -- so BKT = 32 stays the default and IFB_TC_BK=16 the experiment.
// CG = 2: the two CTAs of a cluster form one tcgen05 `cta_group::2` pair -- ONE MMA instruction (issued by the even CTA)
// multiplies the 2 x 128 rows of both CTAs with the block's 256 hyperplanes, each CTA holding only HALF of the
// hyperplane tile in its shared memory (the tensor cores of the pair read both halves): 32 KB instead of 48 KB written
// into a CTA's shared memory per stage, four stages instead of three.
template <int BKT, int CG = 1>
struct Geo {
    static_assert(BKT == 16 || BKT == 32, "K chunk of one or two 16-wide k-steps");
    static_assert(CG == 1 || (CG == 2 && BKT == 32), "the CTA-pair variant uses 64-byte rows");
    static constexpr int BK = BKT;
    static constexpr int STAGES = CG == 2 ? 4 : (BKT == 32 ? 3 : 6);
    static constexpr uint32_t A_BYTES = BM * BKT * 2;
    static constexpr uint32_t B_BYTES = (BN / CG) * BKT * 2;
    static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // xh, xl, wh, wl
    static constexpr uint32_t OFF_STAGES = 0;
    static constexpr uint32_t OFF_MASKS = OFF_STAGES + STAGES * STAGE_BYTES;   // [2 teams][4 quarters]: {left, ambiguous} words [8][32]
    static constexpr uint32_t OFF_LV = OFF_MASKS + 8 * 2048;                   // [2][MAX_TREES_PER_BLOCK][128] leaf values
    static constexpr uint32_t OFF_LD = OFF_LV + 2 * MAX_TREES_PER_BLOCK * BM * 4;   // [2][MAX_TREES_PER_BLOCK][128] leaf depths (u8)
    static constexpr uint32_t OFF_META = OFF_LD + 2 * MAX_TREES_PER_BLOCK * BM;
    static constexpr uint32_t OFF_BARS = (OFF_META + META_RING * META_BYTES + 15u) & ~15u;
    static constexpr int NBARS = 2 * STAGES + 4 + 2 * META_RING + 4;
    static constexpr uint32_t OFF_TMEMPTR = OFF_BARS + NBARS * 8;
    static constexpr uint32_t SMEM_BYTES = OFF_TMEMPTR + 16 + 1024;   // + alignment slack
    static_assert(SMEM_BYTES <= 232448, "stage ring does not fit the 227 KB of an sm_100 CTA");
};

struct Params {
    const unsigned char *meta;     // [n_blocks] BlockMeta
    int32_t n_blocks;
    int32_t kp;                    // padded K
    int32_t k;                     // hyperplane width (== d)
    int64_t n_rows;
    const float *rscale;           // [n_rows] 2^-e_row
    const uint8_t *rflag;          // [n_rows] 0: ordinary row, 1: all-zero row (never ambiguous), 2: non-finite / out-of-range
                                   //          features (every visit takes the exact path)
    const float *xr;               // [n_rows][kp] row-major f32 copy of the rows (exact path)
    const float *w;                // d_ext_w [slots][ks]: the hyperplane weights as stored (exact path)
    const int32_t *idx;            // d_ext_idx [slots][ks] feature index of every term; nullptr: term i reads feature i
    const int32_t *slot_len;       // [slots] number of terms of the slot; nullptr: every slot has k terms
    int32_t ks;                    // row stride of w / idx (the forest's widest hyperplane)
    const double *wabs;            // d_ext_wabs [slots]
    const double *col_off;         // [n_blocks*256]
    int32_t max_depth;
    int32_t total_trees;
    float avg_path;
    int32_t accumulate_only;
    float eb_scale;                // test hook: multiplies the bound (1 = product behaviour)
    double *scores;
    float *path_sum;
    int32_t *depth_sum;
    float *probe;                  // diagnostic: raw accumulators of row tile 0, [128][n_blocks*256]
    unsigned long long *stats;     // [0] stuck visits resolved exactly (diagnostic)
};

// ---- PTX helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a lost transaction / bad descriptor must surface as a launch failure, never as a hung GPU
// (wall-clock bound of ~4 s on %globaltimer; the longest legitimate wait is one block of MMAs, tens of microseconds).
// The polling loop itself is three instructions (try_wait blocks in hardware for a system-dependent time before it
// reports "not yet"): an earlier version polled through a 13-instruction C loop and its spinning warps ate 30 % of the
// issue slots -- and, being ALU instructions, of the pipe the epilogue is bound by (profiles/r02_score_ext_tc_v10_ncu.md).
__device__ __forceinline__ bool mbar_poll(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, 0x10000;\n"
        "IFB_TC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "@p bra IFB_TC_DONE_%=;\n"
        "sub.u32 n, n, 1;\n"
        "setp.ne.u32 p, n, 0;\n"
        "@p bra IFB_TC_WAIT_%=;\n"
        "setp.ne.u32 p, n, 0;\n"          // n == 0: gave up after 65536 polls -> p = false
        "IFB_TC_DONE_%=:\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// the same with cluster scope: for a barrier that threads of the PEER CTA arrive on (release.cluster)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    unsigned long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if ((spin & 0xFFFFu) == 0xFFFFu) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            else if (t1 - t0 > 4000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_poll(bar, parity)) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!mbar_poll(bar, parity)) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) __trap();
    }
}
// diagnostic variant (IFB_TC_STATS): adds the cycles spent in the wait to *acc
__device__ __forceinline__ void mbar_wait_timed(uint32_t bar, uint32_t parity, unsigned long long *acc, bool on) {
    if (!on) {
        mbar_wait(bar, parity);
        return;
    }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    *acc += (unsigned long long)(clock64() - t0);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int32_t c0, int32_t c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// The same load delivered to the same shared-memory offset of every CTA of the cluster named in `mask`; each
// destination CTA's mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *map, int32_t c0, int32_t c1, uint32_t bar,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, "
        "%3}], [%4], %5;" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
        : "memory");
}
// cta_group::2 flavours (CUTLASS sm100: SM100_TMA_2SM_LOAD, SM100_MMA_F16BF16_2x1SM_SS, umma_arrive_multicast_2x1SM).
// The load signals the mbarrier at `bar`, which may live in the PEER CTA of the pair (shared::cluster address).
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap *map, int32_t c0, int32_t c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], "
        "[%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// shared::cluster address of the same shared-memory offset in the EVEN CTA of the pair (bit 24 of a shared window
// address is the CTA's rank parity inside its pair: cute::Sm100MmaPeerBitMask)
__device__ __forceinline__ uint32_t pair_leader_addr(uint32_t a) { return a & 0xFEFFFFFFu; }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, f32 accumulate; issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the pair's MMA: D of BOTH CTAs (+)= A (each CTA's own rows) * B (half in each CTA's shared memory)
__device__ __forceinline__ void umma_f16_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_cg2_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// the same arrive on the mbarrier at this offset in every CTA of the cluster named in `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
// K-major operand tile with 64-byte swizzle: rows of 64 bytes, 8-row groups 512 bytes apart (SBO), LBO unused,
// descriptor version 1 (sm_100), layout type 4 = SWIZZLE_64B   (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// the same with 32-byte rows: 8-row groups 256 bytes apart, layout type 6 = SWIZZLE_32B
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format = F32 (bits 4-5 = 1), a/b format F16 (0), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// min(|a|, |b|, |c|) in one instruction (FMNMX3 with absolute-value modifiers)
__device__ __forceinline__ float fmin3_abs(float a, float b, float c) {
    float r;
    asm("min.abs.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ---- the exact path: one hyperplane decision for one row, evaluated by the whole warp -------------------------------
// Returns (warp-uniform) whether the reference goes LEFT:  sum_{i ascending} (double) fl32(w_i * x_i)  <  off.
//   tier 2: the exact addends p_i summed per lane (i = lane, lane+32, ...) and by a shuffle tree -- a re-association,
//           |S2 - S_ref| <= 2 gamma_{k-1} sum|p_i| <= 4 k 2^-53 max|x| sum|w| =: E2;   |S2 - off| > E2  =>  same outcome;
//   tier 3: the reference's sequential order (every lane redundantly).
// `ix` (nullptr = identity) lists the feature of every term, ascending as the reference stores them (ExtendedUtils.scala:27-34).
__device__ __forceinline__ bool exact_left(const float *__restrict__ xr, const float *__restrict__ wr,
                                           const int32_t *__restrict__ ix, int k, double off, double wabs, int lane) {
    double acc = 0.0;
    float mx = 0.f;
    bool bad = false;
    for (int i = lane; i < k; i += 32) {
        const float xv = __ldg(xr + (ix ? __ldg(ix + i) : i)), wv = __ldg(wr + i);
        const float a = fabsf(xv);
        bad = bad || !(a <= 3.0e38f);
        mx = fmaxf(mx, a);
        acc = __dadd_rn(acc, (double)__fmul_rn(wv, xv));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    bad = __any_sync(0xffffffffu, bad);
    // xor-butterfly: every lane holds the same bits (f64 addition is commutative, the tree is symmetric)
    const double E2 = 4.0 * (double)k * 0x1.0p-53 * ((double)mx * wabs * 1.0000002);
    if (!bad && fabs(acc - off) > E2) return acc < off;
    double sq = 0.0;
    for (int i = 0; i < k; i++) sq = __dadd_rn(sq, (double)__fmul_rn(__ldg(wr + i), __ldg(xr + (ix ? __ldg(ix + i) : i))));
    return sq < off;
}

// ---- main kernel ------------------------------------------------------------------------------------------------------
// CL = CTAs per cluster.  The hyperplane (B) tiles are the same for every row tile, so the CTAs of a cluster fetch each
// B tile from L2 ONCE: CTA r loads rows [r*256/CL, (r+1)*256/CL) of wh / wl and TMA-multicasts them into the same stage
// of all CL CTAs (the operand feed, not the MMA, bounds this kernel: 48 KB per stage per SM from L2 without sharing).
// A stage may be refilled once the MMAs of ALL CL CTAs have read it: tcgen05.commit arrives on every CTA's empty barrier.
template <bool HOOK, int CL, int BKT, int CG>
__global__ void __launch_bounds__(THREADS, 1)
score_ext_tc_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                    const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl, const Params p) {
    using G = Geo<BKT, CG>;
    constexpr bool PAIR = CG == 2;   // the cluster's two CTAs are one cta_group::2 pair (see Geo)
    static_assert(!PAIR || CL == 2, "a pair is a cluster of two");
    constexpr int BK = G::BK, STAGES = G::STAGES;
    constexpr uint32_t A_BYTES = G::A_BYTES, B_BYTES = G::B_BYTES, STAGE_BYTES = G::STAGE_BYTES, OFF_STAGES = G::OFF_STAGES,
                       OFF_MASKS = G::OFF_MASKS, OFF_LV = G::OFF_LV, OFF_LD = G::OFF_LD, OFF_META = G::OFF_META,
                       OFF_BARS = G::OFF_BARS, OFF_TMEMPTR = G::OFF_TMEMPTR;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (s32(smem_raw) + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (base - s32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bars = base + OFF_BARS;
    auto bar_full = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto bar_empty = [&](int s) { return bars + 8u * (uint32_t)(STAGES + s); };
    auto bar_tfull = [&](int b) { return bars + 8u * (uint32_t)(2 * STAGES + b); };
    auto bar_tempty = [&](int b) { return bars + 8u * (uint32_t)(2 * STAGES + 2 + b); };
    auto bar_mfull = [&](int b) { return bars + 8u * (uint32_t)(2 * STAGES + 4 + b); };
    auto bar_mempty = [&](int b) { return bars + 8u * (uint32_t)(2 * STAGES + 4 + META_RING + b); };
    auto bar_lvfull = [&](int b) { return bars + 8u * (uint32_t)(2 * STAGES + 4 + 2 * META_RING + b); };
    auto bar_lvempty = [&](int b) { return bars + 8u * (uint32_t)(2 * STAGES + 4 + 2 * META_RING + 2 + b); };
    uint32_t *tmem_ptr_s = reinterpret_cast<uint32_t *>(sm + OFF_TMEMPTR);

    const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;
    constexpr uint16_t kClusterMask = (uint16_t)((1u << CL) - 1u);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), PAIR ? 1 : CL);   // one commit per CTA of the cluster (one for the whole pair)
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(bar_tfull(b), 1);
            // the eight warps of the team that owns this accumulator buffer -- of BOTH CTAs when the pair shares the MMA
            mbar_init(bar_tempty(b), PAIR ? EPI_WARPS : EPI_WARPS / 2);
            // every LANE arrives on the leaf-buffer barriers, releasing its own shared-memory accesses (no reliance on a
            // warp-level sync in front of a single arrive: also what compute-sanitizer's racecheck can follow)
            mbar_init(bar_lvfull(b), EPI_WARPS / 2 * 32);
            mbar_init(bar_lvempty(b), 32);              // the summing warp
        }
        for (int b = 0; b < META_RING; b++) {
            mbar_init(bar_mfull(b), 1);
            mbar_init(bar_mempty(b), EPI_WARPS / 2 + 1);   // the block's eight epilogue warps + the summing warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: all 512 columns (two 256-column accumulator buffers); this warp also frees them
        if constexpr (PAIR) {   // the same warp of both CTAs, the same destination offset (cute::TMEM::Allocator2Sm)
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_ptr_s)), "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_ptr_s)), "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();   // every CTA's barriers are initialised before a peer signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    // every CTA of a cluster runs the same number of (tile, block, chunk) iterations: CTAs without a real tile left
    // process an all-out-of-range one (TMA zero fill, dead lanes) so that the shared B pipeline stays in lockstep
    const int64_t n_tiles = ((p.n_rows + BM - 1) / BM + gridDim.x - 1) / gridDim.x * gridDim.x;
    const int KC = p.kp / BK;
    const int NB = p.n_blocks;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            const bool st_on = p.stats != nullptr;
            unsigned long long w_mempty = 0, w_empty = 0;
            const long long t_begin = clock64();
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int b = 0; b < NB; b++, it++) {
                    for (int kc = 0; kc < KC; kc++) {
                        mbar_wait_timed(bar_empty(stage), phase ^ 1u, &w_empty, st_on);
                        const uint32_t st = base + OFF_STAGES + (uint32_t)stage * STAGE_BYTES;
                        if constexpr (PAIR) {
                            // both CTAs' loads report to the EVEN CTA's barrier, which expects the bytes of both; each CTA
                            // brings its own rows and its own half of the block's hyperplanes
                            constexpr int HN = BN / 2;
                            const uint32_t fb = pair_leader_addr(bar_full(stage));
                            if (cta_rank == 0) mbar_expect_tx(bar_full(stage), 2 * STAGE_BYTES);
                            tma_load_2d_cg2(st, &map_xh, kc * BK, (int32_t)(tile * BM), fb);
                            tma_load_2d_cg2(st + A_BYTES, &map_xl, kc * BK, (int32_t)(tile * BM), fb);
                            tma_load_2d_cg2(st + 2 * A_BYTES, &map_wh, kc * BK, b * BN + (int)cta_rank * HN, fb);
                            tma_load_2d_cg2(st + 2 * A_BYTES + B_BYTES, &map_wl, kc * BK, b * BN + (int)cta_rank * HN, fb);
                            if (++stage == STAGES) {
                                stage = 0;
                                phase ^= 1u;
                            }
                            continue;
                        }
                        mbar_expect_tx(bar_full(stage), STAGE_BYTES);
                        tma_load_2d(st, &map_xh, kc * BK, (int32_t)(tile * BM), bar_full(stage));
                        tma_load_2d(st + A_BYTES, &map_xl, kc * BK, (int32_t)(tile * BM), bar_full(stage));
                        if constexpr (CL == 1) {
                            tma_load_2d(st + 2 * A_BYTES, &map_wh, kc * BK, b * BN, bar_full(stage));
                            tma_load_2d(st + 2 * A_BYTES + B_BYTES, &map_wl, kc * BK, b * BN, bar_full(stage));
                        } else {
                            // my slice of the block's hyperplanes, delivered to every CTA of the cluster
                            constexpr int SL = BN / CL;
                            const uint32_t so = cta_rank * (uint32_t)(SL * BK * 2);
                            tma_load_2d_mc(st + 2 * A_BYTES + so, &map_wh, kc * BK, b * BN + (int)cta_rank * SL, bar_full(stage),
                                           kClusterMask);
                            tma_load_2d_mc(st + 2 * A_BYTES + B_BYTES + so, &map_wl, kc * BK, b * BN + (int)cta_rank * SL,
                                           bar_full(stage), kClusterMask);
                        }
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
            if (st_on) {
                atomicAdd(p.stats + 1, w_mempty);
                atomicAdd(p.stats + 2, w_empty);
                atomicAdd(p.stats + 3, (unsigned long long)(clock64() - t_begin));
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread issues for the whole CTA =====
        if (lane == 0 && (!PAIR || cta_rank == 0)) {   // the pair's MMAs are issued by its even CTA alone
            const uint32_t idesc = umma_idesc_f16(PAIR ? 2 * BM : BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            const bool st_on = p.stats != nullptr;
            unsigned long long w_tempty = 0, w_full = 0;
            const long long t_begin = clock64();
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int b = 0; b < NB; b++, it++) {
                    const int buf = (int)(it & 1u);
                    // the epilogue (of both CTAs of a pair) has drained this accumulator buffer
                    if constexpr (PAIR) mbar_wait_cluster(bar_tempty(buf), ((it >> 1) & 1u) ^ 1u);
                    else mbar_wait_timed(bar_tempty(buf), ((it >> 1) & 1u) ^ 1u, &w_tempty, st_on);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)buf * BN;
                    for (int kc = 0; kc < KC; kc++) {
                        mbar_wait_timed(bar_full(stage), phase, &w_full, st_on);
                        tc_fence_after();
                        const uint32_t st = base + OFF_STAGES + (uint32_t)stage * STAGE_BYTES;
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ks++) {
                            auto desc = [](uint32_t a) { return BKT == 32 ? umma_desc_sw64(a) : umma_desc_sw32(a); };
                            const uint64_t a_h = desc(st + ks * 32);
                            const uint64_t a_l = desc(st + A_BYTES + ks * 32);
                            const uint64_t b_h = desc(st + 2 * A_BYTES + ks * 32);
                            const uint64_t b_l = desc(st + 2 * A_BYTES + B_BYTES + ks * 32);
                            if constexpr (PAIR) {
                                umma_f16_cg2(d_tmem, a_h, b_h, idesc, (kc | ks) != 0 ? 1u : 0u);
                                umma_f16_cg2(d_tmem, a_l, b_h, idesc, 1u);
                                umma_f16_cg2(d_tmem, a_h, b_l, idesc, 1u);
                            } else {
                                umma_f16(d_tmem, a_h, b_h, idesc, (kc | ks) != 0 ? 1u : 0u);
                                umma_f16(d_tmem, a_l, b_h, idesc, 1u);
                                umma_f16(d_tmem, a_h, b_l, idesc, 1u);
                            }
                        }
                        // the stage may be refilled once these MMAs -- and the peers' -- have read it
                        if constexpr (PAIR) umma_commit_cg2_mc(bar_empty(stage), kClusterMask);
                        else if constexpr (CL == 1) umma_commit(bar_empty(stage));
                        else umma_commit_mc(bar_empty(stage), kClusterMask);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    // accumulators of this block are complete (in both CTAs of a pair)
                    if constexpr (PAIR) umma_commit_cg2_mc(bar_tfull(buf), kClusterMask);
                    else umma_commit(bar_tfull(buf));
                }
            }
            if (st_on) {
                atomicAdd(p.stats + 4, w_tempty);
                atomicAdd(p.stats + 5, w_full);
                atomicAdd(p.stats + 6, (unsigned long long)(clock64() - t_begin));
            }
        }
    } else if (warp == 3 + EPI_WARPS) {
        // ===== descriptor loader: block descriptors (node tables of the walk) into their ring, independent of the operand
        // stages -- a descriptor slot is only released after the block's walk, which must not hold back operand loads =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int b = 0; b < NB; b++, it++) {
                    const int mb = (int)(it % META_RING);
                    mbar_wait(bar_mempty(mb), ((it / META_RING) & 1u) ^ 1u);
                    mbar_expect_tx(bar_mfull(mb), META_BYTES);
                    bulk_g2s(base + OFF_META + (uint32_t)mb * META_BYTES, p.meta + (size_t)b * META_BYTES, META_BYTES,
                             bar_mfull(mb));
                }
            }
        }
    } else if (warp == 2 + EPI_WARPS) {
        // ===== summing warp: adds the leaf values of every block in tree order, owns the per-row results =====
        // lane l owns rows l, l + 32, l + 64, l + 96 of the tile.  Taking the sums out of the walking warps lets the two
        // halves of a lane quarter's warps work on alternate blocks without handing a running sum back and forth.
        const float *lvbuf = reinterpret_cast<const float *>(sm + OFF_LV);
        const uint8_t *ldbuf = reinterpret_cast<const uint8_t *>(sm + OFF_LD);
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            float s[4];
            int32_t ds[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int64_t row = tile * BM + j * 32 + lane;
                const bool live = row < p.n_rows;
                s[j] = (p.accumulate_only && live) ? p.path_sum[row] : 0.f;
                ds[j] = (p.accumulate_only && live && p.depth_sum) ? p.depth_sum[row] : 0;
            }
            for (int b = 0; b < NB; b++, it++) {
                const int buf = (int)(it & 1u);
                const int mb = (int)(it % META_RING);
                mbar_wait(bar_mfull(mb), (it / META_RING) & 1u);
                const int nt = reinterpret_cast<const BlockMeta *>(sm + OFF_META + (uint32_t)mb * META_BYTES)->n_trees;
                mbar_wait(bar_lvfull(buf), (it >> 1) & 1u);
                const float *lv = lvbuf + (size_t)buf * (MAX_TREES_PER_BLOCK * BM) + lane;
                const uint8_t *ld = ldbuf + (size_t)buf * (MAX_TREES_PER_BLOCK * BM) + lane;
                for (int t = 0; t < nt; t++) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        s[j] = s[j] + lv[t * BM + j * 32];   // tree order, one f32 add per tree (Array[Float].sum)
                        ds[j] += (int32_t)ld[t * BM + j * 32];
                    }
                }
                mbar_arrive(bar_lvempty(buf));
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_mempty(mb));
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int64_t row = tile * BM + j * 32 + lane;
                if (row < p.n_rows) {
                    if (!p.accumulate_only) {
                        // IF/extended/ExtendedIsolationForestModel.scala:116-119: Float sum / Int, -Float / Float, Math.pow(2, Double)
                        const float e = __fdiv_rn(s[j], (float)p.total_trees);
                        const float z = __fdiv_rn(-e, p.avg_path);
                        p.scores[row] = exp2((double)z);
                    }
                    if (p.path_sum) p.path_sum[row] = s[j];
                    if (p.depth_sum) p.depth_sum[row] = ds[j];
                }
            }
        }
    } else {
        // ===== epilogue warps: TMEM -> registers -> decision masks -> tree walks =====
        const int ew = warp - 2;                 // 0..15
        const int q = warp & 3;                  // TMEM lane quarter this warp may access (rows q*32 .. q*32+31 of the tile);
                                                 // it is also the warp's scheduler, so the four warps of a quarter share one
        const int gi = ew >> 2;                  // position among the 2 * PW warps of the quarter
        const int pair = gi / PW;                // the quarter's warps work in two TEAMS on alternate blocks: while one team
        const int pi = gi % PW;                  // walks (latency-bound), the other drains (issue-bound) on the same scheduler
        // per pair and quarter: [8 chunks][32 lanes] {"left" bits, "ambiguous" bits}; column j of a chunk = bit 31 - j
        uint2 *mq = reinterpret_cast<uint2 *>(sm + OFF_MASKS + (uint32_t)(pair * 4 + q) * 2048u);
        float *lvbuf = reinterpret_cast<float *>(sm + OFF_LV);
        uint8_t *ldbuf = reinterpret_cast<uint8_t *>(sm + OFF_LD);
        const int pair_bar = 1 + q * 2 + pair;   // named barrier of the two warps of this pair and quarter
        const int buf = pair;                    // blocks with (it & 1) == pair: accumulator / leaf-value buffer `pair`
        uint32_t it = 0;
        const bool st_on = p.stats != nullptr;   // diagnostic cycle accounts (IFB_TC_STATS), reported by warp 2 of each CTA
        unsigned long long w_mfull = 0, w_tfull = 0, w_lvempty = 0, c_drain = 0, c_walk = 0, c_ldtm = 0;
        const long long t_begin = st_on ? clock64() : 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t row = tile * BM + q * 32 + lane;
            const bool live = row < p.n_rows;
            const float scr = live ? __ldg(p.rscale + row) : 1.f;
            const uint32_t flag = live ? (uint32_t)__ldg(p.rflag + row) : 1u;   // dead lanes behave like zero rows
            const uint32_t amb_or = flag == 2u ? 0xFFFFFFFFu : 0u;               // out-of-range row: everything ambiguous
            const uint32_t amb_and = flag == 1u ? 0u : 0xFFFFFFFFu;              // zero row: S' = 0 exactly, never ambiguous
            for (int b = 0; b < NB; b++, it++) {
                if ((int)(it & 1u) != pair) continue;
                const int mb = (int)(it % META_RING);
                mbar_wait_timed(bar_mfull(mb), (it / META_RING) & 1u, &w_mfull, st_on);
                mbar_wait_timed(bar_tfull(buf), (it >> 1) & 1u, &w_tfull, st_on);
                tc_fence_after();
                const long long t_drain0 = st_on ? clock64() : 0;
                const BlockMeta *M = reinterpret_cast<const BlockMeta *>(sm + OFF_META + (uint32_t)mb * META_BYTES);
                const int nt = M->n_trees;
                const int nchunks = (M->n_cols + 31) >> 5;
                // ---- drain, split by columns: this warp's four 32-column chunks -> two bits per accumulator ----
                for (int cc = pi; cc < nchunks; cc += PW) {   // chunks pi, pi + PW, ...
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + cc * 32);
                    uint32_t v[32];
                    const long long t_ld0 = st_on ? clock64() : 0;
                    tmem_ld32(taddr, v);
                    if (st_on) c_ldtm += (unsigned long long)(clock64() - t_ld0);
                    // Fast path (2.75 instructions per column): the "left" bit of every accumulator, and ONE chunk-wide
                    // test for ambiguity -- the smallest |dlt| of the 32 columns against the largest bound of the chunk.
                    // Only chunks that fail it (rare at d <= 64) compute the per-column "ambiguous" bits.
                    const float4 *nt4 = reinterpret_cast<const float4 *>(M->nthr + cc * 32);
                    uint32_t Lq[4] = {0, 0, 0, 0};   // four short dependency chains
                    float mn = INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 c = nt4[j >> 2];                                  // warp-uniform (broadcast) 16-byte load
                        const float d0 = fmaf(c.x, scr, __uint_as_float(v[j]));        // S' - offset' * 2^-e_row
                        const float d1 = fmaf(c.y, scr, __uint_as_float(v[j + 1]));
                        const float d2 = fmaf(c.z, scr, __uint_as_float(v[j + 2]));
                        const float d3 = fmaf(c.w, scr, __uint_as_float(v[j + 3]));
                        Lq[j >> 3] = __funnelshift_l(__float_as_uint(d0), Lq[j >> 3], 1);   // append the sign bits
                        Lq[j >> 3] = __funnelshift_l(__float_as_uint(d1), Lq[j >> 3], 1);
                        Lq[j >> 3] = __funnelshift_l(__float_as_uint(d2), Lq[j >> 3], 1);
                        Lq[j >> 3] = __funnelshift_l(__float_as_uint(d3), Lq[j >> 3], 1);
                        mn = fmin3_abs(d0, d1, mn);
                        mn = fmin3_abs(d2, d3, mn);
                    }
                    // column j of the chunk ends up in bit 31 - j
                    const uint32_t L = (Lq[0] << 24) | (Lq[1] << 16) | (Lq[2] << 8) | Lq[3];
                    uint32_t A = 0;
                    const float emax = HOOK ? M->emax[cc] * p.eb_scale : M->emax[cc];
                    if (__any_sync(0xffffffffu, !(mn >= emax))) {
                        const float2 *ne = M->ne + cc * 32;
                        uint32_t Aq[4] = {0, 0, 0, 0};
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const float2 c = ne[j];
                            const float dlt = fmaf(c.x, scr, __uint_as_float(v[j]));
                            const float tt = fabsf(dlt) - (HOOK ? c.y * p.eb_scale : c.y);   // < 0: the visit is ambiguous
                            Aq[j >> 3] = __funnelshift_l(__float_as_uint(tt), Aq[j >> 3], 1);
                        }
                        A = (Aq[0] << 24) | (Aq[1] << 16) | (Aq[2] << 8) | Aq[3];
                    }
                    mq[cc * 32 + lane] = make_uint2(L, (A | amb_or) & amb_and);
                    if (p.probe && tile == 0 && live) {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            p.probe[(size_t)row * ((size_t)NB * BN) + (size_t)b * BN + cc * 32 + j] = __uint_as_float(v[j]);
                    }
                }
                // the accumulator buffer is free as soon as every warp's chunks are in its registers / masks
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (PAIR) mbar_arrive_cluster(pair_leader_addr(bar_tempty(buf)));   // the issuing CTA's barrier
                    else mbar_arrive(bar_tempty(buf));
                }
                if (st_on) c_drain += (unsigned long long)(clock64() - t_drain0);
                asm volatile("bar.sync %0, %1;" ::"r"(pair_bar), "n"(32 * PW) : "memory");   // the team's 8 mask words are complete
                // the summing warp has consumed the leaf values this buffer held two blocks ago
                mbar_wait_timed(bar_lvempty(buf), ((it >> 1) & 1u) ^ 1u, &w_lvempty, st_on);
                const long long t_walk0 = st_on ? clock64() : 0;
                // ---- walk, split by trees: trees pi, pi + PW, pi + 2 PW, ... of the block, up to four chains per lane ----
                float *lv = lvbuf + (size_t)buf * (MAX_TREES_PER_BLOCK * BM) + q * 32 + lane;
                uint8_t *ld = ldbuf + (size_t)buf * (MAX_TREES_PER_BLOCK * BM) + q * 32 + lane;
                auto walk_group = [&](auto nch_tag, int t0) {   // trees t0 + pi + PW * c, c < NCH
                    constexpr int NCH = decltype(nch_tag)::value;
                    uint32_t cur[NCH], amb[NCH];
#pragma unroll
                    for (int c = 0; c < NCH; c++) {
                        cur[c] = live ? (uint32_t)M->root[t0 + pi + PW * c] : LEAF0;
                        amb[c] = 0;
                    }
                    // fast path: branch-free levels that ignore ambiguity and only remember (bit 31 of amb) whether an
                    // ambiguous INTERNAL node was visited; leaves point at themselves.  13 instructions per level.
#pragma unroll 1
                    for (int lvl = 0; lvl < p.max_depth; lvl++) {
#pragma unroll
                        for (int c = 0; c < NCH; c++) {
                            const uint32_t cu = cur[c];
                            const uint32_t j = cu & 31u;
                            const uint2 m = mq[(cu & 0xE0u) + (uint32_t)lane];
                            const uint2 nx = M->next[cu];
                            amb[c] |= (m.y << j) & ~(cu << 23);          // cu >= 256 (a leaf) clears bit 31
                            cur[c] = ((int32_t)(m.x << j) < 0) ? nx.x : nx.y;
                        }
                    }
                    uint32_t any_amb = 0;
#pragma unroll
                    for (int c = 0; c < NCH; c++) any_amb |= amb[c];
                    if (__any_sync(0xffffffffu, (int32_t)any_amb < 0)) {
                        // careful path (rare at d <= 64, ~1/3 of the groups at d = 1024 where the epilogue is hidden behind
                        // the MMA): walk again from the roots, stop at ambiguous nodes and decide them exactly, one stuck
                        // (lane, chain) at a time, the whole warp cooperating
                        uint32_t stuck = 0;
#pragma unroll
                        for (int c = 0; c < NCH; c++) cur[c] = live ? (uint32_t)M->root[t0 + pi + PW * c] : LEAF0;
                        while (true) {
#pragma unroll 1
                            for (int lvl = 0; lvl < p.max_depth; lvl++) {
#pragma unroll
                                for (int c = 0; c < NCH; c++) {
                                    const uint32_t cu = cur[c];
                                    const uint32_t j = cu & 31u;
                                    const uint2 m = mq[(cu & 0xE0u) + (uint32_t)lane];
                                    const uint2 nx = M->next[cu];
                                    const uint32_t a = (((m.y << j) & ~(cu << 23)) >> 31) | ((stuck >> c) & 1u);
                                    stuck |= a << c;
                                    cur[c] = a ? cu : (((int32_t)(m.x << j) < 0) ? nx.x : nx.y);
                                }
                            }
                            uint32_t sm_mask = __ballot_sync(0xffffffffu, stuck != 0);
                            if (!sm_mask) break;
                            while (sm_mask) {
                                const int Ls = __ffs(sm_mask) - 1;
                                sm_mask &= sm_mask - 1;
                                const uint32_t stL = __shfl_sync(0xffffffffu, stuck, Ls);
                                const int64_t rowL = __shfl_sync(0xffffffffu, row, Ls);
#pragma unroll
                                for (int c = 0; c < NCH; c++) {
                                    if ((stL >> c) & 1u) {   // warp-uniform
                                        const uint32_t col = __shfl_sync(0xffffffffu, cur[c], Ls);
                                        const uint2 nx = M->next[col];
                                        const int32_t slot = M->slot[col];
                                        const double off = __ldg(p.col_off + (size_t)b * BN + col);
                                        const bool left = exact_left(p.xr + (size_t)rowL * p.kp, p.w + (size_t)slot * p.ks,
                                                                     p.idx ? p.idx + (size_t)slot * p.ks : nullptr,
                                                                     p.slot_len ? __ldg(p.slot_len + slot) : p.k, off,
                                                                     __ldg(p.wabs + slot), lane);
                                        if (lane == Ls) {
                                            cur[c] = left ? nx.x : nx.y;
                                            stuck &= ~(1u << c);
                                        }
                                        if (p.stats && lane == 0) atomicAdd(p.stats, 1ull);
                                    }
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < NCH; c++) {
                        const uint32_t lf = cur[c] - LEAF0;
                        lv[(t0 + pi + PW * c) * BM] = M->leafv[lf];
                        ld[(t0 + pi + PW * c) * BM] = M->leafd[lf];
                    }
                };
                const int my_trees = nt > pi ? (nt - pi + PW - 1) / PW : 0;   // <= MAX_TREES_PER_BLOCK / PW
                for (int done = 0; done < my_trees;) {
                    const int left_trees = my_trees - done;
                    const int t0 = done * PW;
                    if (left_trees >= 4) {
                        walk_group(std::integral_constant<int, 4>{}, t0);
                        done += 4;
                    } else if (left_trees == 3) {
                        walk_group(std::integral_constant<int, 3>{}, t0);
                        done += 3;
                    } else if (left_trees == 2) {
                        walk_group(std::integral_constant<int, 2>{}, t0);
                        done += 2;
                    } else {
                        walk_group(std::integral_constant<int, 1>{}, t0);
                        done += 1;
                    }
                }
                if (st_on) c_walk += (unsigned long long)(clock64() - t_walk0);
                // leaf values and depths of this warp's trees are in shared memory: hand them to the summing warp
                mbar_arrive(bar_lvfull(buf));
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_mempty(mb));
                // the pair's mask words may be overwritten (next drain) only when both warps are done walking on them
                asm volatile("bar.sync %0, %1;" ::"r"(pair_bar), "n"(32 * PW) : "memory");
            }
        }
        if (st_on && ew == 0 && lane == 0) {
            atomicAdd(p.stats + 7, w_mfull);
            atomicAdd(p.stats + 8, w_tfull);
            atomicAdd(p.stats + 9, w_lvempty);
            atomicAdd(p.stats + 10, c_drain);
            atomicAdd(p.stats + 11, c_walk);
            atomicAdd(p.stats + 12, (unsigned long long)(clock64() - t_begin));
            atomicAdd(p.stats + 13, c_ldtm);
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();   // no CTA exits while a peer may still multicast into it / signal it
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- per-forest column preparation --------------------------------------------------------------------------------------
// Sparse hyperplanes (extensionLevel < d - 1, or a feature subspace): slot s, terms i < len[s] -> dense row of D weights,
// zeros elsewhere.  A zero weight contributes exact zeros to every product and partial sum of the tensor-core
// accumulation, so the filter's bound (derived for D terms) holds a fortiori; the exact path keeps using the stored terms.
__global__ void ext_tc_densify(const float *__restrict__ w, const int32_t *__restrict__ idx, const int32_t *__restrict__ len,
                               int64_t n_slots, int ks, int D, float *__restrict__ dense) {
    const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n_slots) return;
    const int n = len[s];
    for (int i = lane; i < n; i += 32) dense[s * D + idx[s * ks + i]] = w[s * ks + i];
}

// One warp per accumulator column: power-of-two scaling of the weight row (max |w'| in [0.5, 1)), fp16 hi/lo split,
// ||w'||_2, the node's offset in its scaled domain and the bound of the filter.
__global__ void ext_tc_prepare_cols(const float *__restrict__ w, const int32_t *__restrict__ col_slot,
                                    const double *__restrict__ col_off, int n_cols, int k, int kp, double ck,
                                    __half *__restrict__ wh, __half *__restrict__ wl, unsigned char *__restrict__ meta,
                                    int32_t *__restrict__ flag) {
    const int col = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (col >= n_cols) return;
    const int slot = col_slot[col];
    __half *oh = wh + (size_t)col * kp, *ol = wl + (size_t)col * kp;
    if (slot < 0) {
        for (int i = lane; i < kp; i += 32) oh[i] = ol[i] = __float2half_rn(0.f);
        return;
    }
    const float *wr = w + (size_t)slot * k;
    float mx = 0.f;
    bool bad = false;
    for (int i = lane; i < k; i += 32) {
        const float a = fabsf(wr[i]);
        bad = bad || !(a <= 0x1p60f);
        mx = fmaxf(mx, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    bad = __any_sync(0xffffffffu, bad) || (mx > 0.f && mx < 0x1p-60f);
    int e = 0;
    if (mx > 0.f) (void)frexpf(mx, &e);        // mx = m * 2^e, m in [0.5, 1)
    const float sc = ldexpf(1.f, -e);
    float sq = 0.f;
    for (int i = lane; i < kp; i += 32) {
        float hi = 0.f, lo = 0.f;
        if (i < k && !bad) {
            const float v = wr[i] * sc;        // exact (power of two; |v| < 1)
            const __half h = __float2half_rn(v);
            hi = __half2float(h);
            lo = v - hi;                       // exact in f32
            sq = fmaf(v, v, sq);
        }
        oh[i] = __float2half_rn(hi);
        ol[i] = __float2half_rn(lo);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) {
        BlockMeta *M = reinterpret_cast<BlockMeta *>(meta + (size_t)(col / BN) * META_BYTES);
        const double off = col_off[col];
        const double offs = ldexp(off, -e);                    // exact unless it leaves the f64 range
        float thr = (float)offs;                               // round to nearest; its error is part of c_k
        // an offset beyond the f32 range still decides every ordinary row (|S'| <= k): clamp instead of +-inf so
        // that inf - inf can never appear in the filter
        thr = fminf(fmaxf(thr, -3.0e38f), 3.0e38f);
        // a non-zero offset that rounds to zero would lose its sign for all-zero rows: such forests keep the CUDA-core path
        if (off != 0.0 && thr == 0.f) bad = true;
        // ||w'||_2 from an f32 sum of <= kp squares: relative error <= kp 2^-24, covered by the 2^-10 inflation;
        // the row factor ||x'||_2 is < 1 by construction (ext_tc_prepare_rows)
        const double bnd = ck * sqrt((double)sq) * (1.0 + 0x1.0p-10);
        float bf = (float)bnd;
        if ((double)bf <= bnd) bf = nextafterf(bf, INFINITY);  // strictly above: the filter tests |dlt| - bound >= 0
        M->ne[col % BN] = make_float2(-thr, bf);
        M->nthr[col % BN] = -thr;
        atomicMax(reinterpret_cast<int *>(&M->emax[(col % BN) / 32]), __float_as_int(bf));   // positive floats order like ints
        if (bad) atomicExch(flag, 1);
    }
}

// ---- per-call row preparation ---------------------------------------------------------------------------------------------
// 32 rows per CTA staged through shared memory ([row][kp + 1] f32), one warp per 4 rows: power-of-two row scaling that
// brings ||x'||_2 into [0.5, 1), fp16 hi/lo split, and a row-major f32 copy for the exact path.
__global__ void __launch_bounds__(256) ext_tc_prepare_rows(const float *__restrict__ X, int64_t n_rows, int d, int64_t ld,
                                                           int layout, int kp, __half *__restrict__ xh,
                                                           __half *__restrict__ xl, float *__restrict__ xr,
                                                           float *__restrict__ rscale, uint8_t *__restrict__ rflag) {
    extern __shared__ float tile[];
    const int pitch = kp + 1;
    const int64_t row0 = (int64_t)blockIdx.x * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (layout == IFB_COL_MAJOR) {
        // lanes run along rows (coalesced), warps along columns
        const int64_t row = row0 + lane;
        for (int c = warp; c < d; c += 8) tile[lane * pitch + c] = row < n_rows ? __ldg(X + (int64_t)c * ld + row) : 0.f;
    } else {
        for (int r = warp; r < 32; r += 8) {
            const int64_t row = row0 + r;
            for (int c = lane; c < d; c += 32) tile[r * pitch + c] = row < n_rows ? __ldg(X + row * ld + c) : 0.f;
        }
    }
    __syncthreads();
    for (int r = warp; r < 32; r += 8) {
        const int64_t row = row0 + r;
        if (row >= n_rows) break;
        const float *t = tile + r * pitch;
        float mx = 0.f;
        bool bad = false;
        for (int c = lane; c < d; c += 32) {
            const float a = fabsf(t[c]);
            bad = bad || !(a <= 0x1p60f);
            mx = fmaxf(mx, a);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        bad = __any_sync(0xffffffffu, bad) || (mx > 0.f && mx < 0x1p-60f);
        // step 1: max |x| into [0.5, 1); step 2: the norm of that (in [0.5, sqrt(d)]) into [0.5, 1)
        int e1 = 0;
        if (mx > 0.f && !bad) (void)frexpf(mx, &e1);
        const float sc1 = ldexpf(1.f, -e1);
        float sq = 0.f;
        for (int c = lane; c < d; c += 32) {
            const float v = bad ? 0.f : t[c] * sc1;
            sq = fmaf(v, v, sq);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        int e2 = 0;
        // f32 sum of <= d squares in [0,1): relative error <= d 2^-24; the 2^-10 inflation keeps the true norm below 2^e2
        if (sq > 0.f) (void)frexpf(sqrtf(sq) * (1.0f + 0x1.0p-10f), &e2);
        const float sc = ldexpf(1.f, -(e1 + e2));
        __half *oh = xh + (size_t)row * kp, *ol = xl + (size_t)row * kp;
        float *of = xr + (size_t)row * kp;
        for (int c = lane; c < kp; c += 32) {
            const float x = c < d ? t[c] : 0.f;
            float hi = 0.f, lo = 0.f;
            if (!bad) {
                const float v = x * sc;        // exact power-of-two scaling (tiny elements may flush: absolute error < 2^-126)
                const __half h = __float2half_rn(v);
                hi = __half2float(h);
                lo = v - hi;
            }
            oh[c] = __float2half_rn(hi);
            ol[c] = __float2half_rn(lo);
            of[c] = x;
        }
        if (lane == 0) {
            rscale[row] = bad ? 1.f : sc;
            rflag[row] = bad ? 2 : (mx == 0.f ? 1 : 0);
        }
    }
}

}  // namespace tc

// ---- host side ----------------------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tc_encode_fn() {
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

// fp16 matrix [rows][kp] (K contiguous) -> tensor map with a (bk x box_rows) box and the swizzle (64B / 32B rows) UMMA expects
int make_tc_tmap(CUtensorMap *map, const void *ptr, int64_t rows, int32_t kp, int box_rows, int bk) {
    EncodeTiledFn enc = tc_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return IFB_ECUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (fp16 %lld x %d, box %d x %d) failed with CUresult %d", (long long)rows, kp, bk,
                  box_rows, (int)r);
        return IFB_ECUDA;
    }
    return IFB_OK;
}

// Bound constant of the tensor-core filter (scaled domain):
//   representation   3 * 2^-22 + 4.1 * sqrt(k) * 2^-25     (two-term fp16 splits of both operands, dropped lo*lo term,
//                                                            absolute 2^-25 resolution of fp16 below 2^-14; both
//                                                            ||x'|| and ||w'|| are >= 0.4995)
//   reference        2^-23                                   ((u + k 2^-53)(1 + u) of the f32 products / f64 sum, and the
//                                                            f32 rounding of the scaled offset)
//   accumulation     nsteps * u_acc * 1.01                   (nsteps = 3 kp / 16 MMA accumulation steps, each assumed
//                                                            accurate to u_acc relative to the sum of magnitudes; u_acc
//                                                            defaults to 2^-22 = 4 ulp of f32 -- measured on the part, see
//                                                            profiles/, IFB_TC_UACC_LOG2 overrides)
double tc_bound_constant(int k, int kp) {
    static const double u_acc = getenv("IFB_TC_UACC_LOG2") ? std::ldexp(1.0, atoi(getenv("IFB_TC_UACC_LOG2"))) : 0x1.0p-22;
    const double nsteps = 3.0 * (double)(kp / 16);
    double ck = 3.0 * 0x1.0p-22 + 4.1 * std::sqrt((double)k) * 0x1.0p-25 + 0x1.0p-23 + nsteps * u_acc * 1.01;
    return ck * (1.0 + 0x1.0p-12);
}

}  // namespace

int build_ext_tc_tables(ifb_forest *f, const std::vector<int32_t> &child, const std::vector<int32_t> &hp,
                        const std::vector<float> &leaf, const std::vector<double> &off, const std::vector<uint8_t> &depth,
                        const std::vector<int32_t> &len) {
    using namespace tc;
    const int T = f->num_trees;
    // Width of the GEMM's K dimension = width of the scored matrix.  Fully-extended forests: the hyperplane width;
    // sparse ones: the model's feature count (or, for tables that do not carry it, the largest index read + 1).
    const bool sparse = !f->ext_dense_identity;
    const int k = sparse ? (f->total_num_features > 0 ? f->total_num_features : f->max_feature_index + 1) : f->max_nnz;
    // k <= 1536: ext_tc_prepare_rows stages 32 rows x (k_pad + 1) floats in shared memory (197 KB at the limit); wider
    // hyperplanes keep the CUDA-core wide kernel
    if (T == 0 || k < 1 || k > 1536 || !f->ext_w_safe || f->max_depth > 255 || f->ext_internal_slots == 0) return IFB_OK;
    if (sparse && (f->max_feature_index >= k || (size_t)f->ext_internal_slots * (size_t)k * 4 > ((size_t)1 << 30))) return IFB_OK;
    const int kp = (k + BK - 1) / BK * BK;
    // ---- pack whole trees into 128-column halves of 256-column blocks, in tree order ----
    std::vector<BlockMeta> metas;
    std::vector<int32_t> col_slot;
    std::vector<double> col_off;
    auto new_block = [&](int tree0) {
        metas.emplace_back();
        BlockMeta &m = metas.back();
        std::memset(&m, 0, sizeof m);
        for (int i = 0; i < BN; i++) {
            m.ne[i] = make_float2(0.f, 0.f);
            m.nthr[i] = INFINITY;   // padding columns never look ambiguous to the drain's chunk-wide test
            m.slot[i] = -1;
        }
        for (int i = 0; i < BN / 32; i++) m.emax[i] = 0.f;
        for (int i = 0; i < MAX_NODES_PER_BLOCK; i++) m.next[i] = make_uint2((uint32_t)i, (uint32_t)i);
        m.tree0 = tree0;
        col_slot.resize(metas.size() * BN, -1);
        col_off.resize(metas.size() * BN, 0.0);
    };
    // internal-node / leaf counts per tree
    std::vector<int> tm(T), tl(T);
    for (int t = 0; t < T; t++) {
        const int64_t base = f->node_off[t];
        const int n = f->node_off[t + 1] - f->node_off[t];
        int m = 0;
        for (int q = 0; q < n; q++) m += child[base + q] >= 0;
        if (m > BN || n - m > MAX_LEAVES_PER_BLOCK) return IFB_OK;   // a tree wider than one block: CUDA-core kernels
        tm[t] = m;
        tl[t] = n - m;
    }
    std::vector<int32_t> ref_of;
    for (int ta = 0; ta < T;) {
        // consecutive trees, columns in tree order, as many as fit 256 columns / 16 trees / 512 leaves
        int cnt = 0, cols = 0, leaves = 0;
        while (ta + cnt < T && cnt < MAX_TREES_PER_BLOCK && cols + tm[ta + cnt] <= BN &&
               leaves + tl[ta + cnt] <= MAX_LEAVES_PER_BLOCK) {
            cols += tm[ta + cnt];
            leaves += tl[ta + cnt];
            cnt++;
        }
        new_block(ta);
        BlockMeta *M = &metas.back();
        const int blk = (int)metas.size() - 1;
        M->n_trees = cnt;
        M->n_cols = cols;
        int used = 0, n_leaves = 0;
        for (int i = 0; i < cnt; i++) {
            const int t = ta + i;
            const int64_t base = f->node_off[t];
            const int n = f->node_off[t + 1] - f->node_off[t];
            ref_of.assign(n, 0);
            int ci = 0, li = 0;
            for (int q = 0; q < n; q++) {
                if (child[base + q] >= 0) ref_of[q] = used + ci++;
                else ref_of[q] = (int)LEAF0 + n_leaves + li++;
            }
            for (int q = 0; q < n; q++) {
                const int64_t g = base + q;
                if (child[g] >= 0) {
                    const int col = ref_of[q];
                    M->next[col] = make_uint2((uint32_t)ref_of[child[g]], (uint32_t)ref_of[child[g] + 1]);
                    M->slot[col] = hp[g];
                    col_slot[(size_t)blk * BN + col] = hp[g];
                    col_off[(size_t)blk * BN + col] = off[g];
                } else {
                    M->leafv[ref_of[q] - (int)LEAF0] = leaf[g];
                    M->leafd[ref_of[q] - (int)LEAF0] = depth[g];
                }
            }
            M->root[i] = (uint16_t)ref_of[0];
            used += tm[t];
            n_leaves += tl[t];
        }
        ta += cnt;
    }
    const int NB = (int)metas.size();
    const size_t ncols = (size_t)NB * BN;
    // ---- one arena: wh | wl | meta | col_slot | col_off | flag ----
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_w = al(ncols * kp * 2), b_meta = al((size_t)NB * META_BYTES), b_slot = al(ncols * 4), b_off = al(ncols * 8);
    DeviceGuard dg(f->device);
    IFB_CUDA(cudaMalloc((void **)&f->d_tc_arena, 2 * b_w + b_meta + b_slot + b_off + 256));
    unsigned char *a = f->d_tc_arena;
    f->d_tc_wh = a;
    f->d_tc_wl = a + b_w;
    f->d_tc_meta = a + 2 * b_w;
    f->d_tc_col_slot = reinterpret_cast<int32_t *>(a + 2 * b_w + b_meta);
    f->d_tc_col_off = reinterpret_cast<double *>(a + 2 * b_w + b_meta + b_slot);
    f->d_tc_flag = reinterpret_cast<int32_t *>(a + 2 * b_w + b_meta + b_slot + b_off);
    f->device_bytes += (int64_t)(2 * b_w + b_meta + b_slot + b_off + 256);
    IFB_CUDA(cudaMemcpyAsync(f->d_tc_meta, metas.data(), (size_t)NB * META_BYTES, cudaMemcpyHostToDevice, 0));
    IFB_CUDA(cudaMemcpyAsync(f->d_tc_col_slot, col_slot.data(), ncols * 4, cudaMemcpyHostToDevice, 0));
    IFB_CUDA(cudaMemcpyAsync(f->d_tc_col_off, col_off.data(), ncols * 8, cudaMemcpyHostToDevice, 0));
    IFB_CUDA(cudaMemsetAsync(f->d_tc_flag, 0, 4, 0));
    const double ck = tc_bound_constant(k, kp);
    const int warps_per_cta = 8;
    const float *w_cols = f->d_ext_w;   // [slots][k] rows the columns are prepared from
    float *dense = nullptr;
    if (sparse) {
        // per-slot term counts (exact path) + the zero-filled dense rows (column preparation only, freed below)
        const int64_t slots = f->ext_internal_slots;
        std::vector<int32_t> slot_len((size_t)slots, 0);
        for (size_t g = 0; g < hp.size(); g++)
            if (hp[g] >= 0) slot_len[(size_t)hp[g]] = len[g];
        IFB_CUDA(cudaMalloc((void **)&f->d_tc_slot_len, (size_t)slots * 4));
        f->device_bytes += slots * 4;
        IFB_CUDA(cudaMemcpyAsync(f->d_tc_slot_len, slot_len.data(), (size_t)slots * 4, cudaMemcpyHostToDevice, 0));
        IFB_CUDA(cudaMalloc((void **)&dense, (size_t)slots * k * 4));
        cudaError_t e = cudaMemsetAsync(dense, 0, (size_t)slots * k * 4, 0);
        if (e == cudaSuccess) {
            ext_tc_densify<<<(unsigned)((slots + 7) / 8), 256>>>(f->d_ext_w, f->d_ext_idx, f->d_tc_slot_len, slots, f->max_nnz, k,
                                                                  dense);
            e = cudaGetLastError();
            count_launch();
        }
        if (e != cudaSuccess) {
            cudaFree(dense);
            IFB_CUDA(e);
        }
        w_cols = dense;
    }
    ext_tc_prepare_cols<<<(unsigned)((ncols + warps_per_cta - 1) / warps_per_cta), warps_per_cta * 32>>>(
        w_cols, f->d_tc_col_slot, f->d_tc_col_off, (int)ncols, k, kp, ck, reinterpret_cast<__half *>(f->d_tc_wh),
        reinterpret_cast<__half *>(f->d_tc_wl), f->d_tc_meta, f->d_tc_flag);
    cudaError_t le = cudaGetLastError();
    count_launch();
    int32_t flag = 0;
    if (le == cudaSuccess) le = cudaMemcpy(&flag, f->d_tc_flag, 4, cudaMemcpyDeviceToHost);   // also waits for the uploads above
    cudaFree(dense);
    IFB_CUDA(le);
    f->tc_sparse = sparse;
    f->tc_k = k;
    f->tc_kp = kp;
    f->tc_blocks = NB;
    f->tc_ok = flag == 0;
    return IFB_OK;
}

int launch_score_extended_tc(const ifb_forest *f, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                             double *scores, int32_t *depth_sum, float *path_sum, bool accumulate_only, cudaStream_t stream,
                             float *probe_out) {
    using namespace tc;
    if (!f->tc_ok || d != f->tc_k || getenv("IFB_EXT_NO_TC") != nullptr) return -1;
    if (n_rows == 0) return IFB_OK;
    const int kp = f->tc_kp;
    // test hooks (read per call): IFB_TC_SCALE / IFB_EXT_FAST_SCALE multiply the bound (1e30: every visit takes the exact path)
    const char *sc_env = getenv("IFB_TC_SCALE") ? getenv("IFB_TC_SCALE") : getenv("IFB_EXT_FAST_SCALE");
    const float eb_scale = sc_env ? (float)atof(sc_env) : 1.0f;
    const bool want_stats = getenv("IFB_TC_STATS") != nullptr;
    // rows are processed in chunks so that the per-call scratch (8 bytes per element) stays below ~8 GB
    int64_t chunk = ((int64_t)1 << 33) / ((int64_t)kp * 8);
    chunk = std::max<int64_t>(chunk & ~(int64_t)127, 128 * 1024);
    chunk = std::min<int64_t>(chunk, (n_rows + 127) & ~(int64_t)127);
    struct Scratch {
        cudaStream_t s;
        void *p = nullptr;
        ~Scratch() {
            if (p) cudaFreeAsync(p, s);
        }
    } scr{stream};
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_h = al((size_t)chunk * kp * 2), b_f = al((size_t)chunk * kp * 4), b_n = al((size_t)chunk * 4);
    IFB_CUDA(cudaMallocAsync(&scr.p, 2 * b_h + b_f + 2 * b_n + 512, stream));
    unsigned char *a = reinterpret_cast<unsigned char *>(scr.p);
    __half *xh = reinterpret_cast<__half *>(a), *xl = reinterpret_cast<__half *>(a + b_h);
    float *xr = reinterpret_cast<float *>(a + 2 * b_h);
    float *rscale = reinterpret_cast<float *>(a + 2 * b_h + b_f);
    uint8_t *rflag = reinterpret_cast<uint8_t *>(a + 2 * b_h + b_f + b_n);
    unsigned long long *stats = reinterpret_cast<unsigned long long *>(a + 2 * b_h + b_f + 2 * b_n);
    if (want_stats) IFB_CUDA(cudaMemsetAsync(stats, 0, 128, stream));

    // CTAs per cluster sharing the hyperplane tiles by TMA multicast (IFB_TC_CLUSTER = 1 | 2 | 4 overrides).  Wide
    // hyperplanes are operand-feed bound (measured, 1M x 1024, 256 trees: 96.6 / 91.2 / 84.6 ms with 1 / 2 / 4 CTAs
    // per cluster) and take clusters of four; narrow ones are bound by the epilogue's instruction count (10M x 64,
    // 200 trees: 66.5 / 66.6 / 74.4 ms -- clusters of four strand a few SMs per GPC) and take clusters of two, which
    // pack the 148 SMs exactly.
    // Final kernel (teams, summing warp, descriptor loader), 4M x 64 rows / 200 trees: 20.7 / 21.3 / 24.0 ms with 1 / 2 / 4
    // CTAs per cluster -- an epilogue-bound kernel gains nothing from sharing operand loads and loses a little to the
    // cluster's lockstep; 400K x 1024: 42.5 / 42.1 ms with 2 / 4.
    const int cl_env = getenv("IFB_TC_CLUSTER") ? atoi(getenv("IFB_TC_CLUSTER")) : (kp >= 256 ? 4 : 1);
    int CL = (cl_env == 2 || cl_env == 4) ? cl_env : 1;
    // K chunk per operand stage (IFB_TC_BK = 16 selects the 6 x 24 KB ring, measured SLOWER: the feed is bound by the
    // bytes written into shared memory per MMA cycle, not by the latency of a stage, and 32-byte rows cost TMA efficiency)
    const int bk_env = getenv("IFB_TC_BK") ? atoi(getenv("IFB_TC_BK")) : 32;
    const int bk = bk_env == 16 ? 16 : 32;
    // cta_group::2 pairs (IFB_TC_CG = 1 | 2 overrides): wide hyperplanes are bound by the bytes written into a CTA's
    // shared memory per MMA cycle; a pair halves the hyperplane tile per CTA
    const int cg_env = getenv("IFB_TC_CG") ? atoi(getenv("IFB_TC_CG")) : 1;
    const int cg = (cg_env == 2 && bk == 32) ? 2 : 1;
    const uint32_t smem_bytes = cg == 2 ? Geo<32, 2>::SMEM_BYTES : (bk == 16 ? Geo<16>::SMEM_BYTES : Geo<32>::SMEM_BYTES);
    if (cg == 2) CL = 2;
    CUtensorMap m_wh, m_wl;
    int rc = make_tc_tmap(&m_wh, f->d_tc_wh, (int64_t)f->tc_blocks * BN, kp, BN / CL, bk);
    if (rc) return rc;
    rc = make_tc_tmap(&m_wl, f->d_tc_wl, (int64_t)f->tc_blocks * BN, kp, BN / CL, bk);
    if (rc) return rc;
    const int sms = device_sm_count(f->device);
    const size_t prep_smem = (size_t)32 * (kp + 1) * 4;
    IFB_CUDA(cudaFuncSetAttribute(ext_tc_prepare_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
    // kernel instantiation for (test hook, cluster size)
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const Params);
    const bool hook = eb_scale != 1.0f;
    KernelFn kern = nullptr;
    if (cg == 2) {
        kern = hook ? score_ext_tc_kernel<true, 2, 32, 2> : score_ext_tc_kernel<false, 2, 32, 2>;
    } else if (bk == 32) {
        if (CL == 1) kern = hook ? score_ext_tc_kernel<true, 1, 32, 1> : score_ext_tc_kernel<false, 1, 32, 1>;
        else if (CL == 2) kern = hook ? score_ext_tc_kernel<true, 2, 32, 1> : score_ext_tc_kernel<false, 2, 32, 1>;
        else kern = hook ? score_ext_tc_kernel<true, 4, 32, 1> : score_ext_tc_kernel<false, 4, 32, 1>;
    } else {
        if (CL == 1) kern = hook ? score_ext_tc_kernel<true, 1, 16, 1> : score_ext_tc_kernel<false, 1, 16, 1>;
        else if (CL == 2) kern = hook ? score_ext_tc_kernel<true, 2, 16, 1> : score_ext_tc_kernel<false, 2, 16, 1>;
        else kern = hook ? score_ext_tc_kernel<true, 4, 16, 1> : score_ext_tc_kernel<false, 4, 16, 1>;
    }
    IFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    // how many clusters are co-resident (one CTA per SM; clusters never span GPCs)
    int max_clusters = sms / CL;
    if (CL > 1) {
        cudaLaunchConfig_t qc = {};
        qc.gridDim = dim3((unsigned)(sms / CL * CL));
        qc.blockDim = dim3(THREADS);
        qc.dynamicSmemBytes = smem_bytes;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = (unsigned)CL;
        qa[0].val.clusterDim.y = 1;
        qa[0].val.clusterDim.z = 1;
        qc.attrs = qa;
        qc.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, kern, &qc) == cudaSuccess && nc > 0) max_clusters = std::min(max_clusters, nc);
        else cudaGetLastError();
    }

    for (int64_t r0 = 0; r0 < n_rows; r0 += chunk) {
        const int64_t rows = std::min<int64_t>(chunk, n_rows - r0);
        const float *Xc = layout == IFB_COL_MAJOR ? X + r0 : X + r0 * ld;
        ext_tc_prepare_rows<<<(unsigned)((rows + 31) / 32), 256, prep_smem, stream>>>(Xc, rows, d, ld, layout, kp, xh, xl, xr,
                                                                                     rscale, rflag);
        IFB_CUDA(cudaGetLastError());
        count_launch();
        CUtensorMap m_xh, m_xl;
        rc = make_tc_tmap(&m_xh, xh, rows, kp, BM, bk);
        if (rc) return rc;
        rc = make_tc_tmap(&m_xl, xl, rows, kp, BM, bk);
        if (rc) return rc;
        Params p;
        p.meta = f->d_tc_meta;
        p.n_blocks = f->tc_blocks;
        p.kp = kp;
        p.k = f->tc_k;
        p.n_rows = rows;
        p.rscale = rscale;
        p.rflag = rflag;
        p.xr = xr;
        p.w = f->d_ext_w;
        p.idx = f->tc_sparse ? f->d_ext_idx : nullptr;
        p.slot_len = f->tc_sparse ? f->d_tc_slot_len : nullptr;
        p.ks = f->max_nnz;
        p.wabs = f->d_ext_wabs;
        p.col_off = f->d_tc_col_off;
        p.max_depth = std::max(f->max_depth, 1);
        p.total_trees = f->num_trees;
        p.avg_path = f->avg_path_norm;
        p.accumulate_only = accumulate_only ? 1 : 0;
        p.eb_scale = eb_scale;
        p.scores = scores ? scores + r0 : nullptr;
        p.path_sum = path_sum ? path_sum + r0 : nullptr;
        p.depth_sum = depth_sum ? depth_sum + r0 : nullptr;
        p.probe = (probe_out && r0 == 0) ? probe_out : nullptr;
        p.stats = want_stats ? stats : nullptr;
        const int64_t n_tiles = (rows + BM - 1) / BM;
        const int64_t want_clusters = (n_tiles + CL - 1) / CL;
        const int grid = (int)std::min<int64_t>(want_clusters, max_clusters) * CL;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = CL > 1 ? 1 : 0;
        IFB_CUDA(cudaLaunchKernelEx(&cfg, kern, m_xh, m_xl, m_wh, m_wl, p));
        IFB_CUDA(cudaGetLastError());
        count_launch();
    }
    if (want_stats) {
        unsigned long long h[16] = {0};
        IFB_CUDA(cudaMemcpyAsync(h, stats, 128, cudaMemcpyDeviceToHost, stream));
        IFB_CUDA(cudaStreamSynchronize(stream));
        fprintf(stderr, "[ifb] tensor-core path: %llu visits decided exactly (rows %lld, trees %d)\n", h[0], (long long)n_rows,
                f->num_trees);
        // cycle accounts summed over the CTAs of the last row chunk (one producer lane, one MMA lane, epilogue warp 2 each)
        auto pct = [](unsigned long long a, unsigned long long b) { return b ? 100.0 * (double)a / (double)b : 0.0; };
        fprintf(stderr, "[ifb]   producer: waits descriptor ring %.1f %%, stage free %.1f %%;  MMA issuer: waits accumulator free "
                        "%.1f %%, operands landed %.1f %%;  epilogue warp: waits descriptor %.1f %%, accumulator ready %.1f %%, "
                        "leaf buffer free %.1f %%, drains %.1f %% (of which in tcgen05.ld + wait %.1f %%), walks %.1f %%\n",
                pct(h[1], h[3]), pct(h[2], h[3]), pct(h[4], h[6]), pct(h[5], h[6]), pct(h[7], h[12]), pct(h[8], h[12]),
                pct(h[9], h[12]), pct(h[10], h[12]), pct(h[13], h[12]), pct(h[11], h[12]));
    }
    return IFB_OK;
}

}  // namespace ifb
