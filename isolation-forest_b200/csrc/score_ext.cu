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
    int32_t k;            // real hyperplane width (<= D)
    int32_t fast_ok;      // weights are finite with |w| <= 2^40 and the fast path is not disabled
    double fast_scale;    // dev knob: multiplies the bound (0 = never fall back, huge = always)
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
            // columns beyond the hyperplane width carry zero weights in the blob; they must read as 0, not as data
            // (an Inf / NaN there would turn 0 * x into NaN, while the reference never reads those columns)
            if (live && c < p.k)
                v = p.layout == IFB_COL_MAJOR ? __ldg(p.X + (int64_t)c * p.ld + row) : __ldg(p.X + row * p.ld + c);
            xr[c] = v;
        }
        // per-row prefactor of the fast-path bound; rows with non-finite or huge features never take the fast path
        double xsq = 0.0;
        bool x_ok = true;
#pragma unroll
        for (int c = 0; c < D; c++) {
            xsq += (double)xr[c] * (double)xr[c];
            x_ok = x_ok && (fabsf(xr[c]) <= 0x1p60f);
        }
        const double ebound = ((double)(D + 5) * 0x1.0p-24 * 1.001) * (sqrt(xsq) * 1.0000001) + (double)D * 0x1.0p-140;
        const bool fast = x_ok && p.fast_ok && (ebound == ebound);
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
            const double *wn = off + ipad;
            const float *w = reinterpret_cast<const float *>(wn + ipad);
            int node = 0;
            int c = child[0];
            while (c >= 0) {
                const int hs = slot[node];
                const float4 *wp = reinterpret_cast<const float4 *>(w + (size_t)hs * WS);
                // Fast decision: f32 FMA dot product (four partial chains).  For ANY evaluation order of an f32 FMA
                // dot product |S_f - sum w_i x_i| <= gamma_k * B  (B = sum|w_i x_i|, gamma_k = k u/(1-k u),
                // u = 2^-24), and the reference value S_ref (f32-rounded products, f64 sequential sum) satisfies
                // |S_ref - sum w_i x_i| <= (u + k 2^-53)(1+u) B; with B <= ||w||_2 ||x||_2 (Cauchy-Schwarz):
                //     |S_f - S_ref| <= (k + 2) * 2^-24 * 1.001 * ||w||_2 * ||x||_2 + k * 2^-148 =: E.
                // If |S_f - offset| > E the strict test has the same outcome as the reference's; otherwise (a near
                // tie, ~1e-5 of the visits, or non-finite / out-of-range data: `fast` is then false) the visit is
                // evaluated with the reference's exact arithmetic below.  Decisions are therefore bit-exact.
                bool decided = false, left = false;
                if (fast) {
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int q = 0; q < D / 4; q++) {
                        const float4 w4 = wp[q];
                        s0 = fmaf(w4.x, xr[4 * q + 0], s0);
                        s1 = fmaf(w4.y, xr[4 * q + 1], s1);
                        s2 = fmaf(w4.z, xr[4 * q + 2], s2);
                        s3 = fmaf(w4.w, xr[4 * q + 3], s3);
                    }
                    const float sf = (s0 + s1) + (s2 + s3);
                    const double dlt = (double)sf - off[hs];
                    const double E = ebound * wn[hs] * p.fast_scale;        // ebound = (k+5) 2^-24 1.001 ||x||_2 + slack (per row)
                    if (fabs(dlt) > E) {
                        decided = true;
                        left = dlt < 0.0;
                    }
                }
                if (!__all_sync(__activemask(), decided)) {
                    if (!decided) {
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
                        left = sum < off[hs];
                    }
                }
                node = c + (left ? 0 : 1);
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
    auto go = [&](auto kern) -> int {
        IFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, R, smem, stream>>>(p);
        return IFB_OK;
    };
    int rc = go(score_ext_dense_kernel<D, R>);
    if (rc) return rc;
    IFB_CUDA(cudaGetLastError());
    count_launch();
    return IFB_OK;
}


// ---- wide kernel: fully extended forests with k = d > 64 (BASELINE config 5: d = 1024) -------------------
// A warp owns G rows of the CTA's row tile (rows live in shared memory, [row][d+4] floats).  Per tree and level
// the warp groups its rows by current node; per group it streams the node's weight row once (coalesced LDG.128,
// the forest is L2 resident) and the 32 lanes split the terms (i = 128 j + 4 lane + q).  Three tiers, each one
// PROVABLY giving the reference's decision (the reference adds p_i = (double)fl32(w_i x_i) in index order):
//   1. f32 FMA partial dots combined by a shuffle tree.  |S_1 - S_ref| <= (k+16) 2^-24 1.001 ||w||_2 ||x||_2 =: E1
//      (Higham's bound for an FMA dot product in ANY order, plus the reference's own product roundings,
//      Cauchy-Schwarz for sum|w_i x_i|).  |S_1 - offset| > E1  =>  same outcome as the reference.
//   2. the exact addends p_i, summed in f64 per lane and by a shuffle tree (a re-association):
//      |S_2 - S_ref| <= 2 gamma_{k-1} sum|p_i| <= 4 k 2^-53 max|x| sum|w| =: E2.   |S_2 - offset| > E2 => same.
//   3. the reference's sequential order.
// Tier 2 is reached by ~1e-3 of the visits at k = 1024, tier 3 essentially never; rows with non-finite or huge
// features skip tier 1 (its bound would not hold) and fall through.
struct ScoreExtWideParams {
    const float *X;
    int64_t n_rows, ld;
    int32_t d, layout;
    const float *w;
    const double *off;
    const double *wabs;
    const WideNode *nodes;
    const int32_t *tree_slot;
    const int64_t *tree_node;
    int32_t num_trees, total_trees;
    int32_t rows_per_tile;   // multiple of the warp count
    float avg_path;
    int32_t accumulate_only;
    int32_t fast_ok;
    double fast_scale;
    double *scores;
    float *path_sum;
    int32_t *depth_sum;
};

template <int G, int NW, int PB>
__global__ void __launch_bounds__(NW * 32) score_ext_wide_kernel(const ScoreExtWideParams p) {
    extern __shared__ __align__(16) float xs_w[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = G * NW;
    const int d = p.d;
    const int xstride = d + 4;  // floats; keeps 16-byte alignment, breaks the power-of-two row pitch
    float *rowmax = xs_w + (size_t)R * xstride;                         // [R] max |x| per row (NaN if non-finite)
    double *rownorm = reinterpret_cast<double *>(rowmax + ((R + 1) & ~1));  // [R] ||x||_2, inflated
    const int64_t n_tiles = (p.n_rows + R - 1) / R;
    const int chunks = d / 128;        // full 128-term chunks (4 terms per lane)
    const int tail0 = chunks * 128;    // remaining terms [tail0, d) handled 1 per lane per step
    // PB = weight chunks in flight per warp
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * R;
        __syncthreads();
        if (p.layout == IFB_COL_MAJOR) {
            for (int64_t e = tid; e < (int64_t)R * d; e += NW * 32) {
                const int c = (int)(e / R), r = (int)(e % R);
                const int64_t row = row0 + r;
                xs_w[(size_t)r * xstride + c] = row < p.n_rows ? __ldg(p.X + (int64_t)c * p.ld + row) : 0.f;
            }
        } else {
            for (int64_t e = tid; e < (int64_t)R * d; e += NW * 32) {
                const int r = (int)(e / d), c = (int)(e % d);
                const int64_t row = row0 + r;
                xs_w[(size_t)r * xstride + c] = row < p.n_rows ? __ldg(p.X + row * p.ld + c) : 0.f;
            }
        }
        __syncthreads();
        for (int g = 0; g < G; g++) {
            const int r = warp * G + g;
            float m = 0.f;
            double sq = 0.0;
            bool bad = false;
            for (int c = lane; c < d; c += 32) {
                const float xv = xs_w[(size_t)r * xstride + c];
                const float v = fabsf(xv);
                bad = bad || !(v <= 3.0e38f);
                m = fmaxf(m, v);
                sq += (double)xv * (double)xv;
            }
            for (int o = 16; o > 0; o >>= 1) {
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                sq += __shfl_xor_sync(0xffffffffu, sq, o);
            }
            bad = __any_sync(0xffffffffu, bad);
            if (lane == 0) {
                rowmax[r] = bad ? __int_as_float(0x7fc00000) : m;
                rownorm[r] = sqrt(sq) * 1.000001;
            }
        }
        __syncwarp();

        float s[G];
        int32_t dsum[G];
        bool tier1[G];   // warp-uniform: this row may use the f32 tier
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int64_t row = row0 + warp * G + g;
            const bool live = row < p.n_rows;
            s[g] = (p.accumulate_only && live) ? p.path_sum[row] : 0.f;
            dsum[g] = (p.accumulate_only && live && p.depth_sum) ? p.depth_sum[row] : 0;
            tier1[g] = p.fast_ok && (rowmax[warp * G + g] <= 0x1p60f);
        }
        for (int t = 0; t < p.num_trees; t++) {
            const int64_t base = p.tree_node[t];
            int32_t node[G], ch[G];      // warp-uniform: current node and ITS weight slot (-1 = leaf)
            const int32_t root_slot = __ldg(p.tree_slot + t);
#pragma unroll
            for (int g = 0; g < G; g++) {
                node[g] = 0;
                ch[g] = root_slot;
            }
            while (true) {
                uint32_t pending = 0;    // rows still at an internal node
#pragma unroll
                for (int g = 0; g < G; g++) pending |= (ch[g] >= 0) ? (1u << g) : 0u;
                if (!pending) break;
                while (pending) {
                    const int g0 = __ffs(pending) - 1;
                    int32_t n0 = 0;
#pragma unroll
                    for (int g = 0; g < G; g++) if (g == g0) n0 = node[g];
                    uint32_t grp = 0;
#pragma unroll
                    for (int g = 0; g < G; g++) if (((pending >> g) & 1u) && node[g] == n0) grp |= 1u << g;
                    pending &= ~grp;
                    int32_t slot = 0;
#pragma unroll
                    for (int g = 0; g < G; g++) if (g == g0) slot = ch[g];
                    // the node record and the weight row are independent loads: one L2 round trip per visit
                    const float *wrow = p.w + (int64_t)slot * d;
                    const int4 *nrec = reinterpret_cast<const int4 *>(p.nodes + (base + n0));
                    const int4 na = __ldg(nrec), nb = __ldg(nrec + 1);
                    const double offv = __hiloint2double(na.y, na.x);
                    const float wnorm_f = __int_as_float(na.z);
                    const int32_t cbase = nb.x, slot_l = nb.y, slot_r = nb.z;
                    uint32_t todo = grp;                         // rows whose decision is still open
                    uint32_t leftmask = 0;

                    // ---- tier 1: f32 FMA ----
                    uint32_t g1 = 0;
#pragma unroll
                    for (int g = 0; g < G; g++) if (((grp >> g) & 1u) && tier1[g]) g1 |= 1u << g;
                    const double e1_node = ((double)(d + 16) * 0x1.0p-24 * 1.001) * (double)wnorm_f * p.fast_scale;
                    if (g1) {
                        float a32[G][4];
#pragma unroll
                        for (int g = 0; g < G; g++) a32[g][0] = a32[g][1] = a32[g][2] = a32[g][3] = 0.f;
                        for (int jb = 0; jb < chunks; jb += PB) {
                            float4 wv[PB];
#pragma unroll
                            for (int u = 0; u < PB; u++)
                                if (jb + u < chunks) wv[u] = __ldg(reinterpret_cast<const float4 *>(wrow + (jb + u) * 128) + lane);
#pragma unroll
                            for (int u = 0; u < PB; u++) {
                                if (jb + u < chunks) {
#pragma unroll
                                    for (int g = 0; g < G; g++) {
                                        if ((g1 >> g) & 1u) {
                                            const float4 x4 = *reinterpret_cast<const float4 *>(
                                                xs_w + (size_t)(warp * G + g) * xstride + (jb + u) * 128 + lane * 4);
                                            a32[g][0] = fmaf(wv[u].x, x4.x, a32[g][0]);
                                            a32[g][1] = fmaf(wv[u].y, x4.y, a32[g][1]);
                                            a32[g][2] = fmaf(wv[u].z, x4.z, a32[g][2]);
                                            a32[g][3] = fmaf(wv[u].w, x4.w, a32[g][3]);
                                        }
                                    }
                                }
                            }
                        }
                        for (int i = tail0 + lane; i < d; i += 32) {
                            const float wv1 = __ldg(wrow + i);
#pragma unroll
                            for (int g = 0; g < G; g++)
                                if ((g1 >> g) & 1u) a32[g][0] = fmaf(wv1, xs_w[(size_t)(warp * G + g) * xstride + i], a32[g][0]);
                        }
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            if ((g1 >> g) & 1u) {
                                float v = (a32[g][0] + a32[g][1]) + (a32[g][2] + a32[g][3]);
                                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                                v = __shfl_sync(0xffffffffu, v, 0);
                                const double E1 = e1_node * rownorm[warp * G + g] + (double)d * 0x1.0p-140;
                                const double dlt = (double)v - offv;
                                if (fabs(dlt) > E1) {
                                    todo &= ~(1u << g);
                                    if (dlt < 0.0) leftmask |= 1u << g;
                                }
                            }
                        }
                    }
                    // ---- tier 2: exact addends, f64, lane-parallel ----
                    if (todo) {
                        double acc[G][4];
#pragma unroll
                        for (int g = 0; g < G; g++) acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0.0;
                        for (int j = 0; j < chunks; j++) {
                            const float4 w4 = __ldg(reinterpret_cast<const float4 *>(wrow + j * 128) + lane);
#pragma unroll
                            for (int g = 0; g < G; g++) {
                                if ((todo >> g) & 1u) {
                                    const float4 x4 = *reinterpret_cast<const float4 *>(
                                        xs_w + (size_t)(warp * G + g) * xstride + j * 128 + lane * 4);
                                    acc[g][0] = __dadd_rn(acc[g][0], (double)__fmul_rn(w4.x, x4.x));
                                    acc[g][1] = __dadd_rn(acc[g][1], (double)__fmul_rn(w4.y, x4.y));
                                    acc[g][2] = __dadd_rn(acc[g][2], (double)__fmul_rn(w4.z, x4.z));
                                    acc[g][3] = __dadd_rn(acc[g][3], (double)__fmul_rn(w4.w, x4.w));
                                }
                            }
                        }
                        for (int i = tail0 + lane; i < d; i += 32) {
                            const float wv1 = __ldg(wrow + i);
#pragma unroll
                            for (int g = 0; g < G; g++)
                                if ((todo >> g) & 1u)
                                    acc[g][0] = __dadd_rn(acc[g][0], (double)__fmul_rn(wv1, xs_w[(size_t)(warp * G + g) * xstride + i]));
                        }
                        const double wabs = __ldg(p.wabs + slot);
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            if ((todo >> g) & 1u) {
                                double v = __dadd_rn(__dadd_rn(acc[g][0], acc[g][1]), __dadd_rn(acc[g][2], acc[g][3]));
                                for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));
                                v = __shfl_sync(0xffffffffu, v, 0);
                                const double A = (double)rowmax[warp * G + g] * wabs * 1.0000002;
                                const double E2 = 4.0 * (double)d * 0x1.0p-53 * A;
                                bool left;
                                if (fabs(v - offv) > E2) {
                                    left = v < offv;
                                } else {
                                    // ---- tier 3: the reference's sequential order, every lane redundantly ----
                                    double sq = 0.0;
                                    const float *xr = xs_w + (size_t)(warp * G + g) * xstride;
                                    for (int i = 0; i < d; i++) sq = __dadd_rn(sq, (double)__fmul_rn(__ldg(wrow + i), xr[i]));
                                    left = sq < offv;
                                }
                                if (left) leftmask |= 1u << g;
                            }
                        }
                    }
#pragma unroll
                    for (int g = 0; g < G; g++) {
                        if ((grp >> g) & 1u) {
                            const bool l = (leftmask >> g) & 1u;
                            node[g] = cbase + (l ? 0 : 1);
                            ch[g] = l ? slot_l : slot_r;
                            dsum[g]++;
                        }
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < G; g++) s[g] = s[g] + __ldg(&p.nodes[base + node[g]].leaf);
        }
        if (lane == 0) {
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int64_t row = row0 + warp * G + g;
                if (row < p.n_rows) {
                    if (!p.accumulate_only) {
                        const float e = __fdiv_rn(s[g], (float)p.total_trees);
                        const float z = __fdiv_rn(-e, p.avg_path);
                        p.scores[row] = exp2((double)z);
                    }
                    if (p.path_sum) p.path_sum[row] = s[g];
                    if (p.depth_sum) p.depth_sum[row] = dsum[g];
                }
            }
        }
    }
}

template <int G, int NW, int PB>
int launch_wide(const ifb_forest *f, const ScoreExtWideParams &p0, cudaStream_t stream) {
    ScoreExtWideParams p = p0;
    p.rows_per_tile = G * NW;
    const size_t smem = ((size_t)G * NW * (p.d + 4) + (size_t)((G * NW + 1) & ~1)) * 4 + (size_t)G * NW * 8;
    const int sms = device_sm_count(f->device);
    const int64_t n_tiles = (p.n_rows + G * NW - 1) / (G * NW);
    const int grid = (int)std::min<int64_t>(n_tiles, sms);
    IFB_CUDA(cudaFuncSetAttribute(score_ext_wide_kernel<G, NW, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    score_ext_wide_kernel<G, NW, PB><<<grid, NW * 32, smem, stream>>>(p);
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
    if (getenv("IFB_EXT_GENERIC") == nullptr) {
        // fully-extended forests: all hyperplanes as one GEMM on the tensor cores (score_ext_tc.cu)
        const int rc = launch_score_extended_tc(f, X, n_rows, d, ld, layout, scores, depth_sum, path_sum, accumulate_only, stream);
        if (rc >= 0) return rc;
    }
    if (f->ext_dense_identity && f->max_nnz <= 64 && d <= 64 && getenv("IFB_EXT_GENERIC") == nullptr) {
        const int brc = ensure_ext_blob(const_cast<ifb_forest *>(f));   // first call of this fallback builds its tables
        if (brc) return brc;
    }
    if (f->ext_blob_D > 0 && d <= 64 && getenv("IFB_EXT_GENERIC") == nullptr) {
        ScoreExtDenseParams q;
        q.X = X; q.n_rows = n_rows; q.ld = ld; q.d = d; q.layout = layout;
        q.blob = f->d_ext_blob; q.blob_off = f->d_ext_blob_off;
        q.num_trees = f->num_trees; q.total_trees = f->num_trees; q.blob_max = f->ext_blob_max;
        q.avg_path = f->avg_path_norm; q.accumulate_only = accumulate_only ? 1 : 0;
        q.scores = scores; q.path_sum = path_sum; q.depth_sum = depth_sum;
        q.k = f->max_nnz; q.fast_ok = (f->ext_w_safe && getenv("IFB_EXT_NOFAST") == nullptr) ? 1 : 0;
        q.fast_scale = getenv("IFB_EXT_FAST_SCALE") ? atof(getenv("IFB_EXT_FAST_SCALE")) : 1.0;
        int rc;
        switch (f->ext_blob_D) {
            case 8: rc = launch_dense<8>(f, q, stream); break;
            case 16: rc = launch_dense<16>(f, q, stream); break;
            case 32: rc = launch_dense<32>(f, q, stream); break;
            default: rc = launch_dense<64>(f, q, stream); break;
        }
        if (rc >= 0) return rc;
    }
    if (f->ext_dense_identity && f->max_nnz == d && d > 64 && d % 4 == 0 && getenv("IFB_EXT_GENERIC") == nullptr) {
        ScoreExtWideParams q;
        q.X = X; q.n_rows = n_rows; q.ld = ld; q.d = d; q.layout = layout;
        q.w = f->d_ext_w; q.off = f->d_ext_off; q.wabs = f->d_ext_wabs;
        q.nodes = reinterpret_cast<const WideNode *>(f->d_ext_wide_nodes); q.tree_slot = f->d_ext_tree_slot;
        q.fast_ok = (f->ext_w_safe && getenv("IFB_EXT_NOFAST") == nullptr) ? 1 : 0;
        q.fast_scale = getenv("IFB_EXT_FAST_SCALE") ? atof(getenv("IFB_EXT_FAST_SCALE")) : 1.0;
        q.tree_node = f->d_ext_tree_node;
        q.num_trees = f->num_trees; q.total_trees = f->num_trees; q.rows_per_tile = 0;
        q.avg_path = f->avg_path_norm; q.accumulate_only = accumulate_only ? 1 : 0;
        q.scores = scores; q.path_sum = path_sum; q.depth_sum = depth_sum;
        // rows per warp: as many as shared memory allows (<= 4), 16 warps per CTA
        const size_t budget = (size_t)device_smem_optin(f->device) - 1024;
        int G = (int)std::min<size_t>(4, budget / ((size_t)16 * ((size_t)d + 8) * 4));
        static const int wide_nw = getenv("IFB_WIDE_NW") ? atoi(getenv("IFB_WIDE_NW")) : 32;
        if (wide_nw == 32 && budget / ((size_t)32 * ((size_t)d + 8) * 4) >= 1) {
            if (budget / ((size_t)32 * ((size_t)d + 8) * 4) >= 2 && getenv("IFB_WIDE_G1") == nullptr) return launch_wide<2, 32, 4>(f, q, stream);
            return launch_wide<1, 32, 4>(f, q, stream);
        }
        if (G >= 1) {
            switch (G) {
                case 4: return launch_wide<4, 16, 8>(f, q, stream);
                case 3: return launch_wide<3, 16, 8>(f, q, stream);
                case 2: return launch_wide<2, 16, 8>(f, q, stream);
                default: return launch_wide<1, 16, 8>(f, q, stream);
            }
        }
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
