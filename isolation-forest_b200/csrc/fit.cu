// Tree building for sm_100a: one launch builds every tree of a forest shard, one CTA per tree.
//
// Replaces IsolationTree.fit / generateIsolationTree (IF/IsolationTree.scala:53-183),
// ExtendedIsolationTree.fit / generateExtendedIsolationTree (IF/extended/ExtendedIsolationTree.scala:67-270)
// and the per-tree setup of trainIsolationTrees (IF/core/SharedTrainLogic.scala:276-317).
//
// Determinism contract (DESIGN.md "fit"): a tree is a pure function of (randomSeed, numPartitions, tree id,
// data).  Thread 0 of the CTA owns the tree's java.util.Random stream (48-bit LCG, nextInt/nextDouble/
// nextGaussian exactly as the JDK specifies, StrictMath.log restated from fdlibm) and consumes it in the
// reference's order: node by node in PRE-ORDER (the recursion order of the Scala code), so the node tables
// come out in the persisted pre-order layout.  All data-parallel work of a node (feature min/max, the
// hyperplane dot products, the partition of the node's row list) is spread over the CTA's threads with
// warp-shuffle reductions and ballot/popc scans.  The tree's sampled rows are staged once, feature-major, in
// shared memory or an L2-resident scratch (SampleView), and nodes with at most one row replay the feature
// draws on thread 0 alone instead of running numFeatures block-wide min/max rounds.  This translation unit is
// compiled with -fmad=false: the JVM never contracts a*b+c, so neither may the f64 split arithmetic here.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "ifb_internal.h"

namespace ifb {
namespace {

constexpr int BT = 256;  // threads per CTA (one CTA builds one tree)

struct FitDev {
    const float *X;
    int64_t N, ld;
    int32_t d, layout;
    int32_t n, num_features, bootstrap;
    int64_t random_seed;
    int32_t num_partitions, ext_level, k;
    int32_t tree_begin, height_limit, cap, cap_internal;
    // outputs (per tree at stride cap / cap_internal)
    int32_t *n_nodes, *n_internal;
    int32_t *left, *right;
    int64_t *num_instances;
    int32_t *feature;
    double *threshold;
    double *offset;
    int32_t *hp_slot;
    int32_t *hp_idx;
    float *hp_w;
    int64_t *rows;  // [trees][n] sampled row ids
    // scratch (per tree)
    int32_t *perm, *perm2;       // [n]
    int64_t *hkeys, *hvals;      // [hcap]
    int32_t hcap;
    int32_t *feat_perm;          // [d]
    int32_t *feat_idx;           // [num_features] sorted feature subset
    int32_t *avail;              // [num_features]
    int32_t *e_idx;              // [k]
    double *e_raw;               // [k]
    float *e_w, *e_mn, *e_mx;    // [k]
    // staging (see "staged sample" below)
    int32_t small_in_smem;       // per-tree scratch lists live in shared memory
    int32_t stage;               // 0: read X through rows[]; 1: sample staged in shared memory; 2: in `sample`
    float *sample;               // [trees][d][n] staged sample when stage == 2
    int32_t warp_nodes;          // standard builder: subtrees of <= 32 rows are built by warp 0 alone (no block barriers)
    int32_t dbg;                 // IFB_FIT_DBG: tree 0 prints the cycles of its phases
};

// ---- java.util.Random ---------------------------------------------------------------------------
struct JRandom {
    unsigned long long seed;
    int have_next;
    double next_gauss;
};
__device__ __forceinline__ void jr_init(JRandom &r, long long seed) {
    r.seed = ((unsigned long long)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
    r.have_next = 0;
    r.next_gauss = 0.0;
}
__device__ __forceinline__ int jr_next(JRandom &r, int bits) {
    r.seed = (r.seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int)(long long)(r.seed >> (48 - bits));
}
__device__ int jr_next_int(JRandom &r, int bound) {
    int rr = jr_next(r, 31);
    const int m = bound - 1;
    if ((bound & m) == 0) return (int)(((long long)bound * (long long)rr) >> 31);
    for (int u = rr;; u = jr_next(r, 31)) {
        rr = u % bound;
        if ((int)((unsigned)u - (unsigned)rr + (unsigned)m) >= 0) break;
    }
    return rr;
}
__device__ __forceinline__ double jr_next_double(JRandom &r) {
    const long long hi = (long long)jr_next(r, 26);
    const long long lo = (long long)jr_next(r, 27);
    return (double)((hi << 27) + lo) * 0x1.0p-53;
}
__device__ long long jr_bounded(JRandom &r, long long m) {
    if (m < 0x7fffffffLL) return jr_next_int(r, (int)m);
    const unsigned long long hi = (unsigned long long)jr_next(r, 31), lo = (unsigned long long)jr_next(r, 31);
    return (long long)(((hi << 31) | lo) % (unsigned long long)m);
}

// StrictMath.log: fdlibm __ieee754_log (public algorithm).  No FMA contraction (-fmad=false).
__device__ double fdlibm_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01,
                 Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    int hx = __double2hiint(x);
    unsigned lx = (unsigned)__double2loint(x);
    int k = 0, i, j;
    double hfsq, f, s, z, R, w, t1, t2, dk;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -INFINITY;
        if (hx < 0) return NAN;
        k -= 54;
        x *= two54;
        hx = __double2hiint(x);
        lx = (unsigned)__double2loint(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    x = __hiloint2double(hx | (i ^ 0x3ff00000), (int)lx);
    k += (i >> 20);
    f = x - 1.0;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) {
            if (k == 0) return 0.0;
            dk = (double)k;
            return dk * ln2_hi + dk * ln2_lo;
        }
        R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        dk = (double)k;
        return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    s = f / (2.0 + f);
    dk = (double)k;
    z = s * s;
    i = hx - 0x6147a;
    w = z * z;
    j = 0x6b851 - hx;
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
        hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

__device__ double jr_next_gaussian(JRandom &r) {
    if (r.have_next) {
        r.have_next = 0;
        return r.next_gauss;
    }
    double v1, v2, s;
    do {
        v1 = 2 * jr_next_double(r) - 1;
        v2 = 2 * jr_next_double(r) - 1;
        s = v1 * v1 + v2 * v2;
    } while (s >= 1 || s == 0);
    const double multiplier = sqrt(-2 * fdlibm_log(s) / s);
    r.next_gauss = v2 * multiplier;
    r.have_next = 1;
    return v1 * multiplier;
}

// ---- warp-parallel nextGaussian --------------------------------------------------------------------------------
// java.util.Random.nextGaussian is the polar method: an ATTEMPT draws two nextDouble (4 LCG steps) and is accepted with
// probability pi/4; an accepted attempt yields two values (the second one is cached).  The stream position after n
// values therefore depends on the data, but attempts are independent given their starting state, and the LCG can be
// advanced by any number of steps in O(1):  seed_{i+m} = A^m seed_i + C (A^m - 1)/(A - 1)  (mod 2^48).  The 32 lanes
// of a warp evaluate 32 consecutive attempts from skip-ahead states, a ballot + prefix count puts the accepted ones in
// stream order, and the state after the last USED attempt becomes the new seed -- exactly the values, the order and
// the final stream position of the sequential loop (wide hyperplanes: 1024 draws per node were 95 % of the builder).
struct LcgJump {
    unsigned long long mul, add;   // seed -> mul * seed + add (mod 2^48) == 4 * lane steps
};
__device__ __forceinline__ LcgJump lcg_jump_steps(int steps) {
    const unsigned long long M = (1ULL << 48) - 1;
    unsigned long long cm = 0x5DEECE66DULL, ca = 0xBULL;   // one step
    unsigned long long am = 1ULL, aa = 0ULL;               // identity
    for (int e = steps; e > 0; e >>= 1) {
        if (e & 1) {                                       // acc = cur o acc
            aa = (cm * aa + ca) & M;
            am = (cm * am) & M;
        }
        ca = (cm * ca + ca) & M;                           // cur = cur o cur
        cm = (cm * cm) & M;
    }
    return LcgJump{am, aa};
}
// Called by ALL 32 lanes of one warp; `r` is meaningful in lane 0 only and is updated there.  out[0..n) may live in
// shared or global memory; a trailing __syncwarp makes it visible to the warp.
__device__ void jr_fill_gaussians_warp(JRandom &r, double *out, int n, int lane) {
    const unsigned long long M = (1ULL << 48) - 1;
    unsigned long long seed = __shfl_sync(0xffffffffu, r.seed, 0);
    int have = __shfl_sync(0xffffffffu, r.have_next, 0);
    double cached = __shfl_sync(0xffffffffu, r.next_gauss, 0);
    int pos = 0;
    if (n > 0 && have) {
        if (lane == 0) out[0] = cached;
        pos = 1;
        have = 0;
    }
    const LcgJump jump = lcg_jump_steps(4 * lane);
    while (pos < n) {
        JRandom t;
        t.seed = (jump.mul * seed + jump.add) & M;         // state before attempt number `lane` of this round
        t.have_next = 0;
        t.next_gauss = 0.0;
        const double v1 = 2 * jr_next_double(t) - 1;
        const double v2 = 2 * jr_next_double(t) - 1;
        const double sq = v1 * v1 + v2 * v2;
        const bool acc = !(sq >= 1 || sq == 0);
        const unsigned m = __ballot_sync(0xffffffffu, acc);
        const int rank = __popc(m & ((1u << lane) - 1u));
        const int need = (n - pos + 1) >> 1;               // attempts still to be accepted
        const int got = __popc(m);
        if (acc && rank < need) {
            const double multiplier = sqrt(-2 * fdlibm_log(sq) / sq);
            const int idx = pos + 2 * rank;
            out[idx] = v1 * multiplier;
            if (idx + 1 < n) out[idx + 1] = v2 * multiplier;
            else cached = v2 * multiplier;                 // the value nextGaussian keeps for its next call
        }
        if (got >= need) {
            const int last = __fns(m, 0, need);            // lane of the need-th accepted attempt
            seed = __shfl_sync(0xffffffffu, t.seed, last);
            have = ((n - pos) & 1) ? 1 : 0;
            cached = __shfl_sync(0xffffffffu, cached, last);
            pos = n;
        } else {
            seed = __shfl_sync(0xffffffffu, t.seed, 31);   // all 32 attempts consumed
            pos += 2 * got;
        }
    }
    if (lane == 0) {
        r.seed = seed;
        r.have_next = have;
        r.next_gauss = cached;
    }
    __syncwarp();
}

// The stream effect of `for (m <- F to 1 by -1) nextInt(m)` (values unused: what getFeatureToSplit does on a node that
// cannot be split), evaluated by a whole warp: lane j looks at draw b + j of each round of 32 from a skip-ahead state.
// nextInt(bound) consumes ONE step unless its rejection test fires (probability < bound / 2^31); if no draw rejects the
// stream simply moved F steps, otherwise lane 0 replays the loop sequentially.  `r` is meaningful in lane 0 only.
// jl = lcg_jump_steps(lane + 1), j32 = lcg_jump_steps(32).
__device__ void jr_skip_nextints_warp(JRandom &r, int F, int lane, const LcgJump &jl, const LcgJump &j32) {
    const unsigned long long M = (1ULL << 48) - 1;
    const unsigned long long seed0 = __shfl_sync(0xffffffffu, r.seed, 0);
    unsigned long long base = seed0, last = seed0;
    bool rej = false;
    for (int b = 0; b < F; b += 32) {
        const int i = b + lane;
        const unsigned long long st = (jl.mul * base + jl.add) & M;   // state after the LCG step of draw i
        if (i < F) {
            const int bound = F - i, m = bound - 1;
            if ((bound & m) != 0) {
                const int u = (int)(long long)(st >> 17);             // next(31)
                const int rr = u % bound;
                rej = rej || ((int)((unsigned)u - (unsigned)rr + (unsigned)m) < 0);
            }
        }
        last = __shfl_sync(0xffffffffu, st, min(F - b, 32) - 1);      // state after the last draw of this round
        base = (j32.mul * base + j32.add) & M;
    }
    if (__any_sync(0xffffffffu, rej)) {
        if (lane == 0) {
            r.seed = seed0;
            for (int m = F; m >= 1; m--) (void)jr_next_int(r, m);
        }
    } else if (lane == 0) {
        r.seed = last;
    }
}

// scala.util.Random.shuffle by a whole warp: the len - 1 draws nextInt(len), nextInt(len - 1), ..., nextInt(2) are
// evaluated 32 at a time from skip-ahead states into `draws` (a draw consumes one LCG step unless its rejection test
// fires: then lane 0 redoes the whole shuffle sequentially, probability < len^2 / 2^32), and lane 0 applies the swaps in
// order.  `r` is meaningful in lane 0 only; jl = lcg_jump_steps(lane + 1), j32 = lcg_jump_steps(32).
__device__ void scala_shuffle_warp(JRandom &r, int32_t *a, int len, int32_t *draws, int lane, const LcgJump &jl,
                                   const LcgJump &j32) {
    const unsigned long long M = (1ULL << 48) - 1;
    const unsigned long long seed0 = __shfl_sync(0xffffffffu, r.seed, 0);
    unsigned long long base = seed0, last = seed0;
    const int F = len - 1;   // number of draws
    bool rej = false;
    for (int b = 0; b < F; b += 32) {
        const int i = b + lane;
        const unsigned long long st = (jl.mul * base + jl.add) & M;
        if (i < F) {
            const int bound = len - i, m = bound - 1;
            const int u = (int)(long long)(st >> 17);   // next(31)
            int kk;
            if ((bound & m) == 0) {
                kk = (int)(((long long)bound * (long long)u) >> 31);
            } else {
                kk = u % bound;
                rej = rej || ((int)((unsigned)u - (unsigned)kk + (unsigned)m) < 0);
            }
            draws[i] = kk;
        }
        last = __shfl_sync(0xffffffffu, st, min(F - b, 32) - 1);
        base = (j32.mul * base + j32.add) & M;
    }
    const bool any_rej = __any_sync(0xffffffffu, rej);
    __syncwarp();
    if (lane == 0) {
        if (any_rej) {
            r.seed = seed0;
            for (int n = len; n >= 2; n--) {
                const int kk = jr_next_int(r, n);
                const int32_t tmp = a[n - 1];
                a[n - 1] = a[kk];
                a[kk] = tmp;
            }
        } else {
            int kk = F > 0 ? draws[0] : 0;
            for (int i = 0; i < F; i++) {
                const int n = len - i;
                const int kk_next = (i + 1 < F) ? draws[i + 1] : 0;   // `draws` and `a` never alias: keep the next index in flight
                const int32_t hi = a[n - 1], lo = a[kk];
                a[n - 1] = lo;
                a[kk] = hi;
                kk = kk_next;
            }
            if (F > 0) r.seed = last;
        }
    }
    __syncwarp();
}

// scala.util.Random.shuffle on an int array: for (n <- len to 2 by -1) swap(n-1, nextInt(n))
__device__ void scala_shuffle(JRandom &r, int32_t *a, int len) {
    for (int n = len; n >= 2; n--) {
        const int kk = jr_next_int(r, n);
        const int32_t tmp = a[n - 1];
        a[n - 1] = a[kk];
        a[kk] = tmp;
    }
}

__device__ __forceinline__ float feat_value(const FitDev &p, long long row, int f) {
    return p.layout == IFB_COL_MAJOR ? __ldg(p.X + (long long)f * p.ld + row) : __ldg(p.X + row * p.ld + f);
}

// Value of feature f of the tree's sample slot pr.  The tree's n sampled rows are staged once, feature-major
// ([f][slot]), in shared memory (or an L2-resident scratch) so that the hundreds of min/max and partition passes of a
// tree never go back to the (possibly 50 GB, TLB-hostile) training matrix.
struct SampleView {
    const FitDev &p;
    const float *S;        // staged sample or nullptr
    const int64_t *rows;   // slot -> training row
    __device__ __forceinline__ float operator()(int pr, int f) const {
        return S ? S[(long long)f * p.n + pr] : feat_value(p, rows[pr], f);
    }
};

struct NodeTask {
    int32_t start, count, height, parent, is_right;
};

struct Shared {
    NodeTask cur;
    int32_t done, id, trial, found, leaf, nl;
    int32_t feature;
    double split;
    float red_mn[BT / 32], red_mx[BT / 32];
    int32_t warp_l[BT / 32], warp_r[BT / 32];
    int32_t base_l, base_r;
    int32_t nnz, slot;
    double offset;
    NodeTask stack[72];
    unsigned long long jmul[33], jadd[33];   // extended builder: LCG jumps of 0, 2, ..., 64 steps (j nextDouble draws)
};

// block-wide min/max of feature f over rows perm[start .. start+count)
__device__ void block_min_max(const SampleView &val, Shared &sh, const int32_t *perm, int start, int count, int f,
                              float &mn_out, float &mx_out) {
    float mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < count; i += BT) {
        const float v = val(perm[start + i], f);
        // Scala's `if (v < mn) mn = v; if (v > mx) mx = v` (NaN never replaces)
        if (v < mn) mn = v;
        if (v > mx) mx = v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        if (a < mn) mn = a;
        if (b > mx) mx = b;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        sh.red_mn[threadIdx.x >> 5] = mn;
        sh.red_mx[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    mn = sh.red_mn[0];
    mx = sh.red_mx[0];
#pragma unroll
    for (int w = 1; w < BT / 32; w++) {
        if (sh.red_mn[w] < mn) mn = sh.red_mn[w];
        if (sh.red_mx[w] > mx) mx = sh.red_mx[w];
    }
    mn_out = mn;
    mx_out = mx;
}

// Partition perm[start..start+count) so that rows with go_left come first; returns the left count in sh.nl.
// Order inside the halves is irrelevant to the algorithm (only min/max and sizes are ever taken).
template <typename Pred>
__device__ void block_partition(Shared &sh, int32_t *perm, int32_t *perm2, int start, int count, Pred go_left) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        sh.base_l = 0;
        sh.base_r = 0;
    }
    __syncthreads();
    for (int c0 = 0; c0 < count; c0 += BT) {
        const int i = c0 + threadIdx.x;
        const bool valid = i < count;
        const int32_t pr = valid ? perm[start + i] : 0;
        const bool l = valid && go_left(pr);
        const bool r = valid && !l;
        const unsigned bl = __ballot_sync(0xffffffffu, l), br = __ballot_sync(0xffffffffu, r);
        if (lane == 0) {
            sh.warp_l[warp] = __popc(bl);
            sh.warp_r[warp] = __popc(br);
        }
        __syncthreads();
        int off_l = sh.base_l, off_r = sh.base_r, tot_l = 0, tot_r = 0;
#pragma unroll
        for (int w = 0; w < BT / 32; w++) {
            if (w < warp) {
                off_l += sh.warp_l[w];
                off_r += sh.warp_r[w];
            }
            tot_l += sh.warp_l[w];
            tot_r += sh.warp_r[w];
        }
        const unsigned below = (1u << lane) - 1u;
        if (l) perm2[start + off_l + __popc(bl & below)] = pr;
        if (r) perm2[start + count - 1 - (off_r + __popc(br & below))] = pr;
        __syncthreads();
        if (threadIdx.x == 0) {
            sh.base_l += tot_l;
            sh.base_r += tot_r;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < count; i += BT) perm[start + i] = perm2[start + i];
    if (threadIdx.x == 0) sh.nl = sh.base_l;
    __syncthreads();
}

extern __shared__ __align__(16) unsigned char fit_arena[];

template <bool EXT>
__global__ void __launch_bounds__(BT, EXT ? 2 : 4) fit_kernel(const FitDev p) {
    __shared__ Shared sh;
    const int tl = blockIdx.x;                 // tree index inside the shard
    const int tid = threadIdx.x;
    const long long tree_id = (long long)p.tree_begin + tl;
    // treeSeed = randomSeed + 2*(P+1) + treeId  (IF/IsolationForest.scala:76-78, SharedTrainLogic.scala:283)
    const long long tree_seed = p.random_seed + 2LL * ((long long)p.num_partitions + 1) + tree_id;

    int64_t *rows = p.rows + (int64_t)tl * p.n;
    int32_t *perm, *perm2, *feat_perm, *feat_idx, *avail, *e_idx = nullptr;
    int64_t *hkeys, *hvals;
    double *e_raw = nullptr;
    float *e_w = nullptr, *e_mn = nullptr, *e_mx = nullptr;
    float *stage_buf = nullptr;
    if (p.small_in_smem) {   // carve the arena: 8-byte arrays first (same order as fit_smem_layout on the host)
        unsigned char *a = fit_arena;
        hkeys = (int64_t *)a; a += (size_t)p.hcap * 8;
        hvals = (int64_t *)a; a += (size_t)p.hcap * 8;
        e_raw = (double *)a; a += (size_t)(EXT ? p.k : 0) * 8;
        perm = (int32_t *)a; a += (size_t)p.n * 4;
        perm2 = (int32_t *)a; a += (size_t)p.n * 4;
        feat_perm = (int32_t *)a; a += (size_t)p.d * 4;
        feat_idx = (int32_t *)a; a += (size_t)p.num_features * 4;
        avail = (int32_t *)a; a += (size_t)p.num_features * 4;
        if (EXT) {
            e_idx = (int32_t *)a; a += (size_t)p.k * 4;
            e_w = (float *)a; a += (size_t)p.k * 4;
            e_mn = (float *)a; a += (size_t)p.k * 4;
            e_mx = (float *)a; a += (size_t)p.k * 4;
        }
        a = (unsigned char *)(((uintptr_t)a + 15) & ~(uintptr_t)15);
        if (p.stage == 1) stage_buf = (float *)a;
    } else {
        perm = p.perm + (int64_t)tl * p.n;
        perm2 = p.perm2 + (int64_t)tl * p.n;
        hkeys = p.hkeys + (int64_t)tl * p.hcap;
        hvals = p.hvals + (int64_t)tl * p.hcap;
        feat_perm = p.feat_perm + (int64_t)tl * p.d;
        feat_idx = p.feat_idx + (int64_t)tl * p.num_features;
        avail = p.avail + (int64_t)tl * p.num_features;
        if (EXT) {
            e_idx = p.e_idx + (int64_t)tl * p.k;
            e_raw = p.e_raw + (int64_t)tl * p.k;
            e_w = p.e_w + (int64_t)tl * p.k;
            e_mn = p.e_mn + (int64_t)tl * p.k;
            e_mx = p.e_mx + (int64_t)tl * p.k;
        }
    }
    if (p.stage == 2) stage_buf = p.sample + (int64_t)tl * p.d * p.n;
    int32_t *o_left = p.left + (int64_t)tl * p.cap, *o_right = p.right + (int64_t)tl * p.cap;
    int64_t *o_ninst = p.num_instances + (int64_t)tl * p.cap;

    // ---- per-tree sampling (engine-defined contract, DESIGN.md "fit: sampling") -------------------
    for (int i = tid; i < p.hcap; i += BT) hkeys[i] = -1;
    for (int i = tid; i < p.d; i += BT) feat_perm[i] = i;
    for (int i = tid; i < p.n; i += BT) perm[i] = i;
    __syncthreads();
    JRandom rnd;
    if (tid == 0) {
        jr_init(rnd, tree_seed);
        if (p.bootstrap) {
            for (int i = 0; i < p.n; i++) rows[i] = jr_bounded(rnd, p.N);
        } else {
            const int mask = p.hcap - 1;
            auto slot_of = [&](long long key) {
                int s = (int)(((unsigned long long)key * 0x9E3779B97F4A7C15ULL) >> 40) & mask;
                while (hkeys[s] != -1 && hkeys[s] != key) s = (s + 1) & mask;
                return s;
            };
            for (int i = 0; i < p.n; i++) {
                const long long j = (long long)i + jr_bounded(rnd, p.N - i);
                int sj = slot_of(j);
                const long long aj = hkeys[sj] == j ? hvals[sj] : j;
                const int si = slot_of((long long)i);
                const long long ai = hkeys[si] == (long long)i ? hvals[si] : (long long)i;
                rows[i] = aj;
                sj = slot_of(j);
                hkeys[sj] = j;
                hvals[sj] = ai;
            }
        }
        // featureIndices = shuffle(0 until d).take(numFeatures).sorted   (SharedTrainLogic.scala:300-304)
        scala_shuffle(rnd, feat_perm, p.d);
        for (int a = 0; a < p.num_features; a++) feat_idx[a] = feat_perm[a];
        for (int a = 1; a < p.num_features; a++) {
            const int32_t kv = feat_idx[a];
            int b = a - 1;
            while (b >= 0 && feat_idx[b] > kv) {
                feat_idx[b + 1] = feat_idx[b];
                b--;
            }
            feat_idx[b + 1] = kv;
        }
        // the builder restarts from new Random(treeSeed)  (IF/IsolationTree.scala:63)
        jr_init(rnd, tree_seed);
        sh.stack[0] = NodeTask{0, p.n, 0, -1, 0};
    }
    __syncthreads();
    // ---- stage the tree's sample, feature-major: stage_buf[f][slot] = X(rows[slot], f) ---------------
    if (stage_buf) {
        const long long total = (long long)p.d * p.n;
        if (p.layout == IFB_COL_MAJOR) {
            for (long long e = tid; e < total; e += BT) {
                const int f = (int)(e / p.n), i = (int)(e - (long long)f * p.n);
                stage_buf[e] = __ldg(p.X + (long long)f * p.ld + rows[i]);
            }
        } else {
            for (long long e = tid; e < total; e += BT) {
                const int i = (int)(e / p.d), f = (int)(e - (long long)i * p.d);
                stage_buf[(long long)f * p.n + i] = __ldg(p.X + rows[i] * p.ld + f);
            }
        }
        __syncthreads();
    }
    const SampleView val{p, stage_buf, rows};
    if (EXT && tid < 33) {
        const LcgJump jj = lcg_jump_steps(2 * tid);
        sh.jmul[tid] = jj.mul;
        sh.jadd[tid] = jj.add;
    }
    if (EXT) __syncthreads();

    int sp = 1;        // thread 0 only
    int nnodes = 0;    // thread 0 only
    int ninternal = 0; // thread 0 only

    while (true) {
        if (tid == 0) {
            if (sp == 0) {
                sh.done = 1;
            } else if (nnodes >= p.cap) {
                // Degenerate splits (a column holding both -inf and +inf, or split == min) send every row right and
                // leave a 0-row left leaf level after level; never write past this tree's slice of the tables: stop
                // and report one node too many, which the host turns into an error.
                sh.done = 1;
                nnodes = p.cap + 1;
            } else {
                sh.done = 0;
                sh.cur = sh.stack[--sp];
                sh.id = nnodes++;
                if (sh.cur.is_right) o_right[sh.cur.parent] = sh.id;
            }
        }
        __syncthreads();
        if (sh.done) break;
        const NodeTask cur = sh.cur;
        const int id = sh.id;

        if (!EXT && p.warp_nodes && cur.count <= 32) {
            // ---- small subtrees: warp 0 alone ----------------------------------------------------------------------
            // A node of <= 32 rows fits one row per lane: min/max are shuffles, the partition is a ballot, nothing needs
            // a block barrier.  Warp 0 keeps popping tasks while the top of the stack is small (the same stack, so the
            // pre-order numbering and the order in which the tree's java.util.Random is consumed do not change); the
            // other warps meet it at ONE barrier afterwards.  ~85 % of the nodes of a 256-row tree take this path.
            if (tid < 32) {
                const int lane = tid;
                const LcgJump jump_lane = lcg_jump_steps(lane + 1), jump_32 = lcg_jump_steps(32);
                NodeTask nd = cur;
                int nid = id;
                int32_t *o_feat = p.feature + (int64_t)tl * p.cap;
                double *o_thr = p.threshold + (int64_t)tl * p.cap;
                while (true) {
                    const bool valid = lane < nd.count;
                    const int pr = valid ? perm[nd.start + lane] : 0;
                    // same shortcut as the block path: <= 1 row and no NaN => every feature is tried and rejected
                    bool fast_leaf = nd.count <= 1;
                    if (fast_leaf && nd.count == 1) {
                        const int pr0 = perm[nd.start];
                        int has_nan = 0;
                        for (int i = lane; i < p.num_features; i += 32) {
                            const float v = val(pr0, feat_idx[i]);
                            has_nan |= (v != v) ? 1 : 0;
                        }
                        fast_leaf = !__any_sync(0xffffffffu, has_nan != 0);
                    }
                    int found = 0, feature = -1;
                    double split = 0.0;
                    if (fast_leaf) {
                        jr_skip_nextints_warp(rnd, p.num_features, lane, jump_lane, jump_32);
                    } else {
                        for (int i = lane; i < p.num_features; i += 32) avail[i] = feat_idx[i];
                        __syncwarp();
                        int n_avail = p.num_features;
                        while (!found && n_avail > 0) {
                            int pick = 0;
                            if (lane == 0) pick = jr_next_int(rnd, n_avail);
                            pick = __shfl_sync(0xffffffffu, pick, 0);
                            const int trial = avail[pick];
                            __syncwarp();
                            for (int base = pick; base + 1 < n_avail; base += 32) {   // ListBuffer.remove(pick)
                                const int q = base + lane;
                                const int32_t nxt = (q + 1 < n_avail) ? avail[q + 1] : 0;
                                __syncwarp();
                                if (q + 1 < n_avail) avail[q] = nxt;
                                __syncwarp();
                            }
                            n_avail--;
                            float mn = INFINITY, mx = -INFINITY;
                            if (valid) {
                                const float v = val(pr, trial);
                                if (v < mn) mn = v;    // NaN never replaces (Scala's `if (v < mn) mn = v`)
                                if (v > mx) mx = v;
                            }
                            for (int o = 16; o > 0; o >>= 1) {
                                const float a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
                                if (a < mn) mn = a;
                                if (b > mx) mx = b;
                            }
                            const double dmn = (double)mn, dmx = (double)mx;
                            if (nd.count > 0 && dmn != dmx) {
                                found = 1;
                                feature = trial;
                                if (lane == 0) split = (dmx - dmn) * jr_next_double(rnd) + dmn;  // :145-146
                                split = __shfl_sync(0xffffffffu, split, 0);
                            }
                        }
                    }
                    const bool leaf = fast_leaf || !found || nd.height >= p.height_limit || nd.count <= 1;
                    if (lane == 0) {
                        if (leaf) {
                            o_left[nid] = -1;
                            o_right[nid] = -1;
                            o_feat[nid] = -1;
                            o_thr[nid] = 0.0;
                            o_ninst[nid] = nd.count;
                        } else {
                            o_left[nid] = nid + 1;
                            o_feat[nid] = feature;
                            o_thr[nid] = split;
                            o_ninst[nid] = -1;
                        }
                    }
                    if (!leaf) {
                        const bool l = valid && (double)val(pr, feature) < split;
                        const bool r = valid && !l;
                        const unsigned bl = __ballot_sync(0xffffffffu, l), br = __ballot_sync(0xffffffffu, r);
                        const unsigned below = (1u << lane) - 1u;
                        const int nl = __popc(bl);
                        __syncwarp();   // every lane holds its row: the slots may be overwritten
                        if (l) perm[nd.start + __popc(bl & below)] = pr;
                        if (r) perm[nd.start + nd.count - 1 - __popc(br & below)] = pr;
                        __syncwarp();
                        if (lane == 0) {
                            sh.stack[sp++] = NodeTask{nd.start + nl, nd.count - nl, nd.height + 1, nid, 1};
                            sh.stack[sp++] = NodeTask{nd.start, nl, nd.height + 1, nid, 0};
                        }
                    }
                    // next task, as long as it is small too (otherwise the block takes over again)
                    // (broadcast by shuffles, not through sh.cur: the other warps may still be reading the task the block
                    // popped)
                    int go = 0;
                    NodeTask nx = nd;
                    if (lane == 0) {
                        go = (sp > 0 && sh.stack[sp - 1].count <= 32 && nnodes < p.cap) ? 1 : 0;
                        if (go) {
                            nx = sh.stack[--sp];
                            nid = nnodes++;
                            if (nx.is_right) o_right[nx.parent] = nid;
                        }
                    }
                    go = __shfl_sync(0xffffffffu, go, 0);
                    if (!go) break;
                    nid = __shfl_sync(0xffffffffu, nid, 0);
                    nd.start = __shfl_sync(0xffffffffu, nx.start, 0);
                    nd.count = __shfl_sync(0xffffffffu, nx.count, 0);
                    nd.height = __shfl_sync(0xffffffffu, nx.height, 0);
                    nd.parent = __shfl_sync(0xffffffffu, nx.parent, 0);
                    nd.is_right = __shfl_sync(0xffffffffu, nx.is_right, 0);
                }
            }
            __syncthreads();
            continue;
        }
        if (!EXT) {
            // getFeatureToSplit runs before the stop test and consumes draws (IF/IsolationTree.scala:124-156)
            // A node with at most one row can never find a feature with min != max (unless the row holds a NaN:
            // then min/max stay at +/-inf and the general loop below decides): every feature of the subset is
            // tried and rejected, i.e. the only effect is nextInt(m) for m = numFeatures .. 1 on the tree's stream.
            // Thread 0 replays exactly those draws without the per-trial block reductions.
            bool fast_leaf = cur.count <= 1;
            if (fast_leaf && cur.count == 1) {
                const int pr = perm[cur.start];
                int has_nan = 0;
                for (int i = tid; i < p.num_features; i += BT) {
                    const float v = val(pr, feat_idx[i]);
                    has_nan |= (v != v) ? 1 : 0;
                }
                fast_leaf = __syncthreads_or(has_nan) == 0;
            }
            if (fast_leaf) {
                if (tid == 0) {
                    for (int m = p.num_features; m >= 1; m--) (void)jr_next_int(rnd, m);
                    o_left[id] = -1;
                    o_right[id] = -1;
                    (p.feature + (int64_t)tl * p.cap)[id] = -1;
                    (p.threshold + (int64_t)tl * p.cap)[id] = 0.0;
                    o_ninst[id] = cur.count;
                }
                __syncthreads();   // every thread has consumed sh.cur / sh.id before thread 0 pops the next task
                continue;
            }
            for (int i = tid; i < p.num_features; i += BT) avail[i] = feat_idx[i];
            if (tid == 0) sh.found = 0;
            __syncthreads();
            int n_avail = p.num_features;  // replicated in every thread
            while (true) {
                if (tid == 0) {
                    if (sh.found || n_avail == 0) {
                        sh.trial = -1;
                    } else {
                        const int pick = jr_next_int(rnd, n_avail);
                        sh.trial = avail[pick];
                        for (int q = pick; q + 1 < n_avail; q++) avail[q] = avail[q + 1];  // ListBuffer.remove
                    }
                }
                __syncthreads();
                const int trial = sh.trial;
                if (trial < 0) break;
                n_avail--;
                float mn = 0.f, mx = 0.f;
                if (cur.count > 0) block_min_max(val, sh, perm, cur.start, cur.count, trial, mn, mx);
                if (tid == 0 && cur.count > 0) {
                    const double dmn = (double)mn, dmx = (double)mx;
                    if (dmn != dmx) {
                        sh.found = 1;
                        sh.feature = trial;
                        sh.split = (dmx - dmn) * jr_next_double(rnd) + dmn;  // :145-146
                    }
                }
                __syncthreads();
            }
            if (tid == 0) {
                sh.leaf = (!sh.found || cur.height >= p.height_limit || cur.count <= 1) ? 1 : 0;
                int32_t *o_feat = p.feature + (int64_t)tl * p.cap;
                double *o_thr = p.threshold + (int64_t)tl * p.cap;
                if (sh.leaf) {
                    o_left[id] = -1;
                    o_right[id] = -1;
                    o_feat[id] = -1;
                    o_thr[id] = 0.0;
                    o_ninst[id] = cur.count;
                } else {
                    o_left[id] = id + 1;
                    o_feat[id] = sh.feature;
                    o_thr[id] = sh.split;
                    o_ninst[id] = -1;
                }
            }
            __syncthreads();
            if (sh.leaf) continue;
            const int f = sh.feature;
            const double split = sh.split;
            // left = x < split (f32 widened, strict); everything else goes right.  (The reference filters
            // right with x >= split, dropping NaN rows; NaN training features are outside the contract.)
            block_partition(sh, perm, perm2, cur.start, cur.count,
                            [&](int32_t pr) { return (double)val(pr, f) < split; });
            if (tid == 0) {
                const int nl = sh.nl;
                sh.stack[sp++] = NodeTask{cur.start + nl, cur.count - nl, cur.height + 1, id, 1};
                sh.stack[sp++] = NodeTask{cur.start, nl, cur.height + 1, id, 0};
            }
            __syncthreads();
        } else {
            // ---- extended: IF/extended/ExtendedIsolationTree.scala:139-260 -----------------------
            double *o_off = p.offset + (int64_t)tl * p.cap;
            int32_t *o_slot = p.hp_slot + (int64_t)tl * p.cap;
            const int dim = p.num_features;
            const int nnz = p.k;  // min(extensionLevel + 1, dim)
            if (tid == 0) sh.leaf = (cur.height >= p.height_limit || cur.count <= 1) ? 1 : 0;  // :152-153
            __syncthreads();
            long long tq0 = p.dbg ? clock64() : 0, tq1 = 0, tq2 = 0, tq3 = 0, tq4 = 0, tq5 = 0;
            if (!sh.leaf) {
                for (int i = tid; i < dim; i += BT) feat_perm[i] = i;
                __syncthreads();
                if (tid < 32) {
                    // :160 shuffle (its draws 32 at a time from skip-ahead states, the swaps in order by lane 0), :168
                    // chosen coordinates, then the :169 Gaussians by the whole warp (same values, order and stream
                    // position as the sequential loop: in the reference the draws of :169 follow the shuffle and
                    // interleave with nothing else)
                    const LcgJump jump_lane = lcg_jump_steps(tid + 1), jump_32 = lcg_jump_steps(32);
                    scala_shuffle_warp(rnd, feat_perm, dim, avail, tid, jump_lane, jump_32);
                    for (int i = tid; i < nnz; i += 32) e_idx[i] = feat_idx[feat_perm[i]];
                    __syncwarp();
                    if (p.dbg) tq1 = clock64();
                    jr_fill_gaussians_warp(rnd, e_raw, nnz, tid);
                    if (tid == 0) {
                        double sq = 0.0;
                        for (int i = 0; i < nnz; i++) sq += e_raw[i] * e_raw[i];          // :174-179 (sequential f64 sum)
                        const double norm = sqrt(sq);
                        sh.offset = norm;                                                 // parked for the division below
                        if (norm == 0) sh.leaf = 1;                                       // :183-184
                    }
                }
                __syncthreads();
                if (!sh.leaf) {
                    const double norm = sh.offset;
                    for (int i = tid; i < nnz; i += BT) e_w[i] = (float)(e_raw[i] / norm);    // :190-195
                }
                __syncthreads();
            }
            if (sh.leaf) {
                if (tid == 0) {
                    o_left[id] = -1;
                    o_right[id] = -1;
                    o_off[id] = 0.0;
                    o_slot[id] = -1;
                    o_ninst[id] = cur.count;
                }
                __syncthreads();
                continue;
            }
            if (p.dbg) tq2 = clock64();
            // per-coordinate min/max over the node's rows: one warp per coordinate, lanes along rows; four coordinates
            // in flight per warp (the staged sample may live in an L2 scratch: one dependent load chain per coordinate
            // left the warp waiting on latency)
            if (cur.count <= 32) {
                // small node: one THREAD per coordinate (no shuffles; every load of a thread is independent)
                for (int kk = tid; kk < nnz; kk += BT) {
                    const int j = e_idx[kk];
                    float mn = INFINITY, mx = -INFINITY;
                    int i = 0;
                    for (; i + 8 <= cur.count; i += 8) {
                        float v[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) v[u] = val(perm[cur.start + i + u], j);
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            if (v[u] < mn) mn = v[u];
                            if (v[u] > mx) mx = v[u];
                        }
                    }
                    for (; i < cur.count; i++) {
                        const float v = val(perm[cur.start + i], j);
                        if (v < mn) mn = v;
                        if (v > mx) mx = v;
                    }
                    e_mn[kk] = mn;
                    e_mx[kk] = mx;
                }
            } else {
                const int lane = tid & 31, warp = tid >> 5;
                constexpr int NW = BT / 32;
                for (int k0 = warp; k0 < nnz; k0 += 4 * NW) {
                    float mn[4], mx[4];
                    int j[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        mn[u] = INFINITY;
                        mx[u] = -INFINITY;
                        j[u] = (k0 + u * NW < nnz) ? e_idx[k0 + u * NW] : e_idx[k0];
                    }
                    for (int i = lane; i < cur.count; i += 32) {
                        const int pr = perm[cur.start + i];
                        float v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) v[u] = val(pr, j[u]);
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (v[u] < mn[u]) mn[u] = v[u];
                            if (v[u] > mx[u]) mx[u] = v[u];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        for (int o = 16; o > 0; o >>= 1) {
                            const float a = __shfl_xor_sync(0xffffffffu, mn[u], o), b = __shfl_xor_sync(0xffffffffu, mx[u], o);
                            if (a < mn[u]) mn[u] = a;
                            if (b > mx[u]) mx[u] = b;
                        }
                        if (lane == 0 && k0 + u * NW < nnz) {
                            e_mn[k0 + u * NW] = mn[u];
                            e_mx[k0 + u * NW] = mx[u];
                        }
                    }
                }
            }
            __syncthreads();
            if (p.dbg) tq3 = clock64();
            if (tid < 32) {
                // :201-217: p_j = mn if mn == mx else mn + nextDouble() * (mx - mn), in coordinate order.  Which coordinates
                // draw is known (e_mn / e_mx), so the draws are taken 32 coordinates at a time from skip-ahead states
                // (nextDouble = two LCG steps, no rejection); the f64 sum of w_j * p_j stays sequential on lane 0.
                const unsigned long long M48 = (1ULL << 48) - 1;
                unsigned long long base = __shfl_sync(0xffffffffu, rnd.seed, 0);
                for (int b = 0; b < nnz; b += 32) {
                    const int kk = b + tid;
                    const bool in = kk < nnz;
                    const double mn = in ? (double)e_mn[kk] : 0.0, mx = in ? (double)e_mx[kk] : 0.0;
                    const bool draws = in && !(mn == mx);
                    const unsigned dm = __ballot_sync(0xffffffffu, draws);
                    const int before = __popc(dm & ((1u << tid) - 1u));
                    double iv = mn;
                    if (draws) {
                        JRandom t;
                        t.seed = (sh.jmul[before] * base + sh.jadd[before]) & M48;
                        t.have_next = 0;
                        t.next_gauss = 0.0;
                        iv = mn + jr_next_double(t) * (mx - mn);
                    }
                    if (in) e_raw[kk] = iv;   // the raw Gaussians are no longer needed
                    const int nd = __popc(dm);
                    base = (sh.jmul[nd] * base + sh.jadd[nd]) & M48;
                }
                __syncwarp();
                if (tid == 0) {
                    rnd.seed = base;
                    double off = 0.0;
                    int kk = 0;
                    for (; kk + 8 <= nnz; kk += 8) {   // operands first, then the f64 sum in coordinate order
                        double pw[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) pw[u] = (double)e_w[kk + u] * e_raw[kk + u];
#pragma unroll
                        for (int u = 0; u < 8; u++) off += pw[u];
                    }
                    for (; kk < nnz; kk++) off += (double)e_w[kk] * e_raw[kk];
                    sh.offset = off;
                    sh.slot = ninternal++;
                    o_left[id] = id + 1;
                    o_off[id] = off;
                    o_slot[id] = sh.slot;
                    o_ninst[id] = -1;
                }
            }
            __syncthreads();
            if (p.dbg) tq4 = clock64();
            // canonical order (:220-226): rank of each chosen index among the chosen indices.  e_idx[i] =
            // feat_idx[feat_perm[i]] with feat_idx ascending, so when every coordinate is chosen the rank is feat_perm[i].
            // The ordered copy also goes to shared memory (e_mn / e_mx are free now) for the partition's dot products.
            int32_t *h_idx = p.hp_idx + ((int64_t)tl * p.cap_internal + sh.slot) * p.k;
            float *h_w = p.hp_w + ((int64_t)tl * p.cap_internal + sh.slot) * p.k;
            float *s_w = e_mn;
            int32_t *s_idx = reinterpret_cast<int32_t *>(e_mx);
            for (int i = tid; i < nnz; i += BT) {
                const int32_t me = e_idx[i];
                int rank = 0;
                if (nnz == dim) {
                    rank = feat_perm[i];
                } else {
                    for (int q = 0; q < nnz; q++) rank += (e_idx[q] < me) ? 1 : 0;
                }
                const float wv = e_w[i];
                h_idx[rank] = me;
                h_w[rank] = wv;
                s_idx[rank] = me;
                s_w[rank] = wv;
            }
            __syncthreads();
            const double off = sh.offset;
            // :230-232  left iff dot(x) < offset with the reference's arithmetic (f32 products, f64 sum in index order).
            // One WARP per row: every lane sums the exact addends of its 1/32 of the terms, a butterfly adds the partial
            // sums -- a re-association whose distance from the sequential sum is <= 4 k 2^-53 max|x| sum|w| (the same
            // bound, with the same proof, as tier 2 of the scoring kernels) -- and only rows closer to
            // the offset than that, or with non-finite features, redo the sum sequentially.  All the loads of a row are
            // in flight at once; a thread per row walked 1024 dependent terms with the staged sample in L2.
            uint8_t *go_left_flag = reinterpret_cast<uint8_t *>(e_raw);    // [n] by sample slot; e_raw (8 k bytes) is free by now
            // (wide hyperplanes and small nodes only: with few terms, or all 256 threads busy, a thread per row is faster)
            const bool warp_rows = nnz >= 256 && p.n <= 8 * p.k && cur.count <= 64;
            if (warp_rows) {
                const int lane = tid & 31, warp = tid >> 5;
                double wabs = 0.0;
                for (int q = lane; q < nnz; q += 32) wabs += fabs((double)s_w[q]);
                for (int o = 16; o > 0; o >>= 1) wabs += __shfl_xor_sync(0xffffffffu, wabs, o);
                wabs *= 1.0000001;   // lane-parallel sum instead of a sequential one: <= k 2^-53 relative
                for (int i = warp; i < cur.count; i += BT / 32) {
                    const int pr = perm[cur.start + i];
                    double acc = 0.0;
                    float mxa = 0.f;
                    bool bad = false;
                    for (int q0 = lane; q0 < nnz; q0 += 32 * 8) {
                        float v[8], w8[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const int q = q0 + 32 * u;
                            v[u] = q < nnz ? val(pr, s_idx[q]) : 0.f;
                            w8[u] = q < nnz ? s_w[q] : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const float a = fabsf(v[u]);
                            bad = bad || !(a <= 3.0e38f);
                            mxa = fmaxf(mxa, a);
                            acc = acc + (double)__fmul_rn(w8[u], v[u]);
                        }
                    }
                    for (int o = 16; o > 0; o >>= 1) {
                        acc = acc + __shfl_xor_sync(0xffffffffu, acc, o);
                        mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, o));
                    }
                    bad = __any_sync(0xffffffffu, bad);
                    const double E2 = 4.0 * (double)nnz * 0x1.0p-53 * ((double)mxa * wabs * 1.0000002);
                    bool left;
                    if (!bad && fabs(acc - off) > E2) {
                        left = acc < off;
                    } else {   // every lane redundantly: the reference's own order
                        double sq = 0.0;
                        for (int q = 0; q < nnz; q++) sq = sq + (double)__fmul_rn(s_w[q], val(pr, s_idx[q]));
                        left = sq < off;
                    }
                    if (lane == 0) go_left_flag[pr] = left ? 1 : 0;
                }
                __syncthreads();
            }
            block_partition(sh, perm, perm2, cur.start, cur.count, [&](int32_t pr) {
                if (warp_rows) return go_left_flag[pr] != 0;
                double sum = 0.0;
                int q = 0;
                for (; q + 32 <= nnz; q += 32) {    // 32 loads in flight (the staged sample may sit in L2), the f64 sum in
                    float v[32];                    // index order
#pragma unroll
                    for (int u = 0; u < 32; u++) v[u] = val(pr, s_idx[q + u]);
#pragma unroll
                    for (int u = 0; u < 32; u++) sum = sum + (double)__fmul_rn(s_w[q + u], v[u]);
                }
                for (; q + 8 <= nnz; q += 8) {
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) v[u] = val(pr, s_idx[q + u]);
#pragma unroll
                    for (int u = 0; u < 8; u++) sum = sum + (double)__fmul_rn(s_w[q + u], v[u]);
                }
                for (; q < nnz; q++) sum = sum + (double)__fmul_rn(s_w[q], val(pr, s_idx[q]));
                return sum < off;
            });
            if (p.dbg && tl == 0 && tid == 0) {
                tq5 = clock64();
                printf("[ifb fit dbg] node %d count %d: shuffle %lld gauss+norm %lld minmax %lld offsets %lld rank+partition %lld cycles\n",
                       id, cur.count, tq1 - tq0, tq2 - tq1, tq3 - tq2, tq4 - tq3, tq5 - tq4);
            }
            if (tid == 0) {
                const int nl = sh.nl;
                sh.stack[sp++] = NodeTask{cur.start + nl, cur.count - nl, cur.height + 1, id, 1};
                sh.stack[sp++] = NodeTask{cur.start, nl, cur.height + 1, id, 0};
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        p.n_nodes[tl] = nnodes;
        if (EXT) p.n_internal[tl] = ninternal;
    }
}

struct DevBuf {
    void *p = nullptr;
    cudaStream_t s;
    explicit DevBuf(cudaStream_t st) : s(st) {}
    ~DevBuf() {
        if (p) cudaFreeAsync(p, s);
    }
    int alloc(size_t bytes) {
        cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 16, s);
        if (e != cudaSuccess) {
            set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? IFB_ENOMEM : IFB_ECUDA;
        }
        return IFB_OK;
    }
    template <typename T>
    T *as() { return reinterpret_cast<T *>(p); }
};

// Pinned host staging block shared by the fit calls of this process (copy + compaction phase is serialised).
struct PinnedStage {
    std::mutex mu;
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {   // caller holds mu
        if (bytes <= cap) return IFB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("pinned staging allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
            return IFB_ENOMEM;
        }
        cap = want;
        return IFB_OK;
    }
};
PinnedStage g_stage;

// heightLimit = ceil(log10(n)/log10(2))  (IF/IsolationTree.scala:60-61)
int height_limit_of(int32_t n) { return (int)std::ceil(std::log10((double)n) / std::log10(2.0)); }

}  // namespace
}  // namespace ifb

using namespace ifb;

extern "C" int ifb_fit_device(int32_t device, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                              const ifb_fit_params *prm, ifb_forest **out, void *stream_) {
    IFB_REQUIRE(out, "out is null");
    *out = nullptr;
    IFB_REQUIRE(prm && X, "null argument");
    IFB_REQUIRE(layout == IFB_COL_MAJOR || layout == IFB_ROW_MAJOR, "unknown layout %d", layout);
    IFB_REQUIRE(d >= 1 && n_rows >= 1, "empty training matrix");
    IFB_REQUIRE(layout == IFB_COL_MAJOR ? ld >= n_rows : ld >= d, "leading dimension %lld too small", (long long)ld);
    IFB_REQUIRE(prm->num_estimators > 0, "parameter numEstimators must be > 0, got %d", prm->num_estimators);
    // messages of validateAndResolveParams (IF/core/SharedTrainLogic.scala:43-75)
    IFB_REQUIRE(prm->num_features > 0, "parameter maxFeatures specifying the use of %d features, but >0 features are required.",
                prm->num_features);
    IFB_REQUIRE(prm->num_features <= d, "parameter maxFeatures specifying the use of %d features, but only %d features are available.",
                prm->num_features, d);
    IFB_REQUIRE(prm->num_samples >= 2, "parameter maxSamples specifying the use of %d samples, but >=2 samples are required.",
                prm->num_samples);
    IFB_REQUIRE((int64_t)prm->num_samples <= n_rows,
                "parameter maxSamples specifying the use of %d samples, but only %lld samples are in the input dataset.",
                prm->num_samples, (long long)n_rows);
    IFB_REQUIRE(prm->num_samples <= (1 << 20), "numSamples %d exceeds the device builder's limit of 1048576", prm->num_samples);
    const bool ext = prm->extension_level >= 0;
    // IF/extended/ExtendedIsolationForest.scala:57-68
    IFB_REQUIRE(!ext || prm->extension_level <= prm->num_features - 1,
                "parameter extensionLevel given invalid value %d, but must be in [0, %d] for a subspace of %d features.",
                prm->extension_level, prm->num_features - 1, prm->num_features);
    int32_t tb = prm->tree_begin, te = prm->tree_end;
    if (tb == 0 && te == 0) te = prm->num_estimators;
    IFB_REQUIRE(0 <= tb && tb <= te && te <= prm->num_estimators, "tree shard [%d,%d) outside [0,%d)", tb, te,
                prm->num_estimators);
    const int ntrees = te - tb;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available; this engine has no CPU fallback");
        return IFB_ENOGPU;
    }
    IFB_REQUIRE(device >= 0 && device < ndev, "device %d out of range", device);
    DeviceGuard dg(device);
    cudaStream_t stream = (cudaStream_t)stream_;

    // IFB_FIT_TIMING=1 (diagnostic): wall-clock phases of this call on stderr
    const bool timing = std::getenv("IFB_FIT_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(now() - t).count();
    };
    const auto t_start = now();
    const int n = prm->num_samples;
    const int hl = height_limit_of(n);
    IFB_REQUIRE(hl <= 32, "height limit %d too large", hl);
    const int k = ext ? std::min(prm->extension_level + 1, prm->num_features) : 1;
    // node capacity: standard trees have <= 2n-1 nodes; extended trees may keep empty leaves, bound by the
    // complete tree of height hl, but never more than 2*internal+1 with internal <= min(2^hl - 1, ...)
    // (a standard tree with degenerate splits can carry up to two extra nodes per level: 0-row leaves, which
    // ifb_forest_create_standard then rejects with the reference's ExternalNode requirement, IF/Nodes.scala:27-31)
    int64_t cap = ext ? std::min<int64_t>((1LL << (hl + 1)) - 1, 4LL * n + 1) : 2LL * n - 1 + 2LL * hl + 2;
    int64_t cap_internal = ext ? std::min<int64_t>((1LL << hl) - 1, 2LL * n) : 0;
    if (ext) cap = std::min<int64_t>(cap, 2 * cap_internal + 1);
    int hcap = 1;
    while (hcap < 4 * n) hcap <<= 1;

    FitDev p;
    std::memset(&p, 0, sizeof p);
    p.X = X; p.N = n_rows; p.ld = ld; p.d = d; p.layout = layout;
    p.n = n; p.num_features = prm->num_features; p.bootstrap = prm->bootstrap ? 1 : 0;
    p.random_seed = prm->random_seed; p.num_partitions = prm->num_partitions; p.ext_level = prm->extension_level;
    p.k = k; p.tree_begin = tb; p.height_limit = hl; p.cap = (int32_t)cap; p.cap_internal = (int32_t)cap_internal;
    p.hcap = hcap;

    // every output / scratch array of the launch is a 256-byte aligned slice of ONE stream-ordered allocation (two
    // dozen pool allocations per call were a tenth of a small fit)
    DevBuf b_arena(stream);
    int rc;
    const size_t T = (size_t)std::max(ntrees, 1);
    struct Slice {
        void **dst;
        size_t bytes, at;
    };
    std::vector<Slice> slices;
    size_t arena_bytes = 0;
    auto take = [&](void **dst, size_t bytes) {
        slices.push_back(Slice{dst, bytes, arena_bytes});
        arena_bytes += (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
    };
    take((void **)&p.n_nodes, T * 4);
    take((void **)&p.n_internal, T * 4);
    take((void **)&p.left, T * cap * 4);
    take((void **)&p.right, T * cap * 4);
    take((void **)&p.num_instances, T * cap * 8);
    take((void **)&p.rows, T * n * 8);
    take((void **)&p.perm, T * n * 4);
    take((void **)&p.perm2, T * n * 4);
    take((void **)&p.hkeys, T * hcap * 8);
    take((void **)&p.hvals, T * hcap * 8);
    take((void **)&p.feat_perm, T * d * 4);
    take((void **)&p.feat_idx, T * prm->num_features * 4);
    take((void **)&p.avail, T * prm->num_features * 4);
    p.feature = nullptr; p.threshold = nullptr; p.offset = nullptr; p.hp_slot = nullptr; p.hp_idx = nullptr; p.hp_w = nullptr;
    p.e_idx = nullptr; p.e_raw = nullptr; p.e_w = nullptr; p.e_mn = nullptr; p.e_mx = nullptr;
    if (ext) {
        take((void **)&p.offset, T * cap * 8);
        take((void **)&p.hp_slot, T * cap * 4);
        take((void **)&p.hp_idx, T * cap_internal * k * 4);
        take((void **)&p.hp_w, T * cap_internal * k * 4);
        take((void **)&p.e_idx, T * k * 4);
        take((void **)&p.e_raw, T * k * 8);
        take((void **)&p.e_w, T * k * 4);
        take((void **)&p.e_mn, T * k * 4);
        take((void **)&p.e_mx, T * k * 4);
    } else {
        take((void **)&p.feature, T * cap * 4);
        take((void **)&p.threshold, T * cap * 8);
    }
    if ((rc = b_arena.alloc(arena_bytes))) return rc;
    for (const Slice &sl : slices) *sl.dst = b_arena.as<unsigned char>() + sl.at;

    // shared-memory arena (must mirror the carve-up at the top of fit_kernel)
    const size_t small_bytes = (size_t)hcap * 16 + (ext ? (size_t)k * 8 : 0) +
                               4 * ((size_t)2 * n + d + 2 * (size_t)prm->num_features + (ext ? 4 * (size_t)k : 0));
    const size_t small_aligned = (small_bytes + 15) & ~(size_t)15;
    const size_t sample_bytes = (size_t)d * n * 4;
    constexpr size_t kSmallMax = 64 << 10, kArenaMax = 200 << 10, kScratchMax = (size_t)2 << 30;
    const bool no_stage = std::getenv("IFB_FIT_NO_STAGE") != nullptr;   // test hook: the unstaged path
    p.small_in_smem = small_bytes <= kSmallMax && !no_stage;
    p.stage = 0;
    p.warp_nodes = std::getenv("IFB_FIT_NO_WARP_NODES") == nullptr ? 1 : 0;   // A/B hook: 0 = every node through the block path
    p.dbg = std::getenv("IFB_FIT_DBG") != nullptr ? 1 : 0;      // tree 0 prints the cycles of its phases (device printf)
    DevBuf b_sample(stream);
    if (!no_stage) {
        if (p.small_in_smem && small_aligned + sample_bytes <= kArenaMax) p.stage = 1;
        else if (T * sample_bytes <= kScratchMax) {
            p.stage = 2;
            if ((rc = b_sample.alloc(T * sample_bytes))) return rc;
            p.sample = b_sample.as<float>();
        }
    }
    // The builder is barrier-latency bound (profiles/r01_fit_kernel_ncu.md: 25 stall cycles per issue on
    // __syncthreads, 5 % issue utilisation), so what matters is how many trees are in flight.  A shared-memory sample
    // of 100+ KB leaves one CTA per SM; when that would take more than one wave, stage in the L2 scratch instead (20 KB
    // of shared memory per CTA, every tree resident at once): 512 trees, d = 128: 5.7 -> 2.4 ms.
    if (p.stage == 1 && T * sample_bytes <= kScratchMax) {
        const size_t per_cta = small_aligned + sample_bytes + 3072;          // + static Shared + the 1 KB reserve
        const size_t ctas_per_sm = std::max<size_t>(1, ((size_t)228 << 10) / per_cta);
        const bool one_wave = (size_t)ntrees <= ctas_per_sm * (size_t)device_sm_count(device);
        const char *force = std::getenv("IFB_FIT_STAGE");                    // test hook: 1 / 2 pins the choice
        const bool to_scratch = force ? std::atoi(force) == 2 : !one_wave;
        if (to_scratch) {
            p.stage = 2;
            if ((rc = b_sample.alloc(T * sample_bytes))) return rc;
            p.sample = b_sample.as<float>();
        }
    }
    const size_t dyn = p.small_in_smem ? small_aligned + (p.stage == 1 ? sample_bytes : 0) : 0;
    if (ntrees > 0) {
        if (ext) {
            IFB_CUDA(cudaFuncSetAttribute(fit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            fit_kernel<true><<<ntrees, BT, dyn, stream>>>(p);
        } else {
            IFB_CUDA(cudaFuncSetAttribute(fit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            fit_kernel<false><<<ntrees, BT, dyn, stream>>>(p);
        }
        IFB_CUDA(cudaGetLastError());
        count_launch();
    }
    if (timing) {
        const double t_launch = ms_since(t_start);
        cudaStreamSynchronize(stream);
        std::fprintf(stderr, "[ifb fit] alloc+launch %.3f ms, kernel done at %.3f ms (stage=%d small_in_smem=%d dyn=%zu)\n",
                     t_launch, ms_since(t_start), p.stage, p.small_in_smem, dyn);
    }
    // node tables back to the host (one pinned staging block, grown on demand and reused across calls: pageable
    // D2H of the T x cap tables was a third of a 512-tree fit), then compacted into the persisted layout
    std::unique_lock<std::mutex> stage_lock(g_stage.mu);
    const size_t per_node = 4 + 4 + 8 + (ext ? 8 + 4 : 4 + 8);
    // hyperplane rows of all trees ride in the same batch when they are small enough to stage whole
    const size_t hp_elems = ext ? T * (size_t)cap_internal * k : 0;
    // dense hyperplanes (k == d: indices are 0..k-1 by construction) never leave the device: the forest
    // gathers its scoring tables from the builder's weight array directly (forest.cu::create_extended_from_device)
    const bool hp_on_device = ext && k == d && std::getenv("IFB_FIT_HOST_HP") == nullptr;
    const bool hp_staged = ext && !hp_on_device && hp_elems * 8 <= ((size_t)64 << 20);
    const size_t stage_bytes = T * 8 + T * (size_t)cap * per_node + (hp_staged ? hp_elems * 8 + 16 : 0) + 64;
    if ((rc = g_stage.reserve(stage_bytes))) return rc;
    unsigned char *sp = (unsigned char *)g_stage.p;
    auto carve = [&](size_t bytes) { void *r = sp; sp += (bytes + 7) & ~(size_t)7; return r; };
    int64_t *ninst = (int64_t *)carve(T * cap * 8);
    double *thr = ext ? nullptr : (double *)carve(T * cap * 8);
    double *off = ext ? (double *)carve(T * cap * 8) : nullptr;
    int32_t *left = (int32_t *)carve(T * cap * 4), *right = (int32_t *)carve(T * cap * 4);
    int32_t *feature = ext ? nullptr : (int32_t *)carve(T * cap * 4);
    int32_t *slot = ext ? (int32_t *)carve(T * cap * 4) : nullptr;
    int32_t *n_nodes = (int32_t *)carve(T * 4), *n_int = (int32_t *)carve(T * 4);
    int32_t *hidx_all = hp_staged ? (int32_t *)carve(hp_elems * 4) : nullptr;
    float *hw_all = hp_staged ? (float *)carve(hp_elems * 4) : nullptr;
    if (hp_staged && ntrees > 0) {
        IFB_CUDA(cudaMemcpyAsync(hidx_all, p.hp_idx, hp_elems * 4, cudaMemcpyDeviceToHost, stream));
        IFB_CUDA(cudaMemcpyAsync(hw_all, p.hp_w, hp_elems * 4, cudaMemcpyDeviceToHost, stream));
    }
    IFB_CUDA(cudaMemcpyAsync(n_nodes, p.n_nodes, T * 4, cudaMemcpyDeviceToHost, stream));
    IFB_CUDA(cudaMemcpyAsync(left, p.left, T * cap * 4, cudaMemcpyDeviceToHost, stream));
    IFB_CUDA(cudaMemcpyAsync(right, p.right, T * cap * 4, cudaMemcpyDeviceToHost, stream));
    IFB_CUDA(cudaMemcpyAsync(ninst, p.num_instances, T * cap * 8, cudaMemcpyDeviceToHost, stream));
    if (ext) {
        IFB_CUDA(cudaMemcpyAsync(n_int, p.n_internal, T * 4, cudaMemcpyDeviceToHost, stream));
        IFB_CUDA(cudaMemcpyAsync(off, p.offset, T * cap * 8, cudaMemcpyDeviceToHost, stream));
        IFB_CUDA(cudaMemcpyAsync(slot, p.hp_slot, T * cap * 4, cudaMemcpyDeviceToHost, stream));
    } else {
        IFB_CUDA(cudaMemcpyAsync(feature, p.feature, T * cap * 4, cudaMemcpyDeviceToHost, stream));
        IFB_CUDA(cudaMemcpyAsync(thr, p.threshold, T * cap * 8, cudaMemcpyDeviceToHost, stream));
    }
    IFB_CUDA(cudaStreamSynchronize(stream));
    std::vector<int32_t> node_off(ntrees + 1, 0);
    for (int t = 0; t < ntrees; t++) {
        if (n_nodes[t] < 1 || n_nodes[t] > cap) {
            set_error("internal: tree %d produced %d nodes (capacity %lld)", tb + t, n_nodes[t], (long long)cap);
            return IFB_ESTATE;
        }
        node_off[t + 1] = node_off[t] + n_nodes[t];
    }
    const int64_t total = node_off[ntrees];
    std::vector<int32_t> c_left(total), c_right(total), c_feat;
    std::vector<int64_t> c_ninst(total);
    std::vector<double> c_thr, c_off;
    std::vector<int64_t> hp_off;
    std::vector<int32_t> hp_idx;
    std::vector<float> hp_w;
    if (ext) {
        c_off.resize(total);
        hp_off.assign(total + 1, 0);
        size_t internal_total = 0;
        for (int t = 0; t < ntrees; t++) internal_total += (size_t)std::max(n_int[t], 0);
        hp_idx.reserve(internal_total * k);
        hp_w.reserve(internal_total * k);
    } else {
        c_feat.resize(total);
        c_thr.resize(total);
    }
    std::vector<int32_t> hidx_t;
    std::vector<float> hw_t;
    std::vector<int64_t> src_row(hp_on_device ? (size_t)total : 0, -1);
    for (int t = 0; t < ntrees; t++) {
        const int64_t src = (int64_t)t * cap, dst = node_off[t];
        const int nn = n_nodes[t];
        std::memcpy(&c_left[dst], &left[src], nn * 4);
        std::memcpy(&c_right[dst], &right[src], nn * 4);
        std::memcpy(&c_ninst[dst], &ninst[src], nn * 8);
        if (!ext) {
            std::memcpy(&c_feat[dst], &feature[src], nn * 4);
            std::memcpy(&c_thr[dst], &thr[src], nn * 8);
        } else {
            std::memcpy(&c_off[dst], &off[src], nn * 8);
            if (hp_on_device) {
                for (int i = 0; i < nn; i++)
                    if (left[src + i] != -1) src_row[(size_t)(dst + i)] = (int64_t)t * cap_internal + slot[src + i];
                continue;
            }
            const int ni = n_int[t];
            const int32_t *hidx_p;
            const float *hw_p;
            if (hp_staged) {
                hidx_p = hidx_all + (size_t)t * cap_internal * k;
                hw_p = hw_all + (size_t)t * cap_internal * k;
            } else {
                hidx_t.resize((size_t)ni * k);
                hw_t.resize((size_t)ni * k);
                if (ni > 0) {
                    IFB_CUDA(cudaMemcpy(hidx_t.data(), p.hp_idx + (int64_t)t * cap_internal * k, (size_t)ni * k * 4,
                                        cudaMemcpyDeviceToHost));
                    IFB_CUDA(cudaMemcpy(hw_t.data(), p.hp_w + (int64_t)t * cap_internal * k, (size_t)ni * k * 4,
                                        cudaMemcpyDeviceToHost));
                }
                hidx_p = hidx_t.data();
                hw_p = hw_t.data();
            }
            for (int i = 0; i < nn; i++) {
                const int s = slot[src + i];
                if (left[src + i] != -1) {
                    hp_idx.insert(hp_idx.end(), hidx_p + (size_t)s * k, hidx_p + (size_t)(s + 1) * k);
                    hp_w.insert(hp_w.end(), hw_p + (size_t)s * k, hw_p + (size_t)(s + 1) * k);
                }
                hp_off[dst + i + 1] = (int64_t)hp_idx.size();
            }
        }
    }
    stage_lock.unlock();
    if (timing) std::fprintf(stderr, "[ifb fit] tables on the host and compacted at %.3f ms\n", ms_since(t_start));
    if (hp_on_device) {
        DeviceHyperplanes dh;
        dh.w = p.hp_w;
        dh.src_row = src_row.data();
        rc = create_extended_from_device(device, ntrees, node_off.data(), c_left.data(), c_right.data(), c_ninst.data(),
                                         c_off.data(), k, dh, n, d, out);
        if (timing) std::fprintf(stderr, "[ifb fit] forest handle (device-resident hyperplanes) created at %.3f ms\n", ms_since(t_start));
        return rc;
    }
    rc = ext ? ifb_forest_create_extended(device, ntrees, node_off.data(), c_left.data(), c_right.data(), c_ninst.data(),
                                          c_off.data(), hp_off.data(), hp_idx.data(), hp_w.data(), n, d, out)
             : ifb_forest_create_standard(device, ntrees, node_off.data(), c_left.data(), c_right.data(), c_feat.data(),
                                          c_thr.data(), c_ninst.data(), n, d, out);
    if (timing) std::fprintf(stderr, "[ifb fit] forest handle created at %.3f ms\n", ms_since(t_start));
    return rc;
}

extern "C" int ifb_fit_host(int32_t device, const float *X, int64_t n_rows, int32_t d, int64_t ld, int32_t layout,
                            const ifb_fit_params *prm, ifb_forest **out) {
    IFB_REQUIRE(out && X && prm, "null argument");
    IFB_REQUIRE(d >= 1 && n_rows >= 1, "empty training matrix");
    IFB_REQUIRE(layout == IFB_COL_MAJOR ? ld >= n_rows : ld >= d, "leading dimension %lld too small", (long long)ld);
    DeviceGuard dg(device);
    // The builder touches only numEstimators * numSamples rows, but which ones is decided on the device, so
    // the matrix is staged whole (it is needed on the device for the threshold pass of fit anyway).  Only the
    // addressed extent of the caller's (possibly strided) buffer is read: (d-1)*ld + n_rows resp. (n_rows-1)*ld + d
    // elements; the device copy is compact.
    const bool cm = layout == IFB_COL_MAJOR;
    const int64_t ld_dev = cm ? ((n_rows + 3) & ~3LL) : d;
    const size_t width = (size_t)(cm ? n_rows : d) * 4, height = (size_t)(cm ? d : n_rows);
    float *dX = nullptr;
    IFB_CUDA(cudaMalloc((void **)&dX, (size_t)ld_dev * height * 4));
    cudaError_t e = cudaMemcpy2D(dX, (size_t)ld_dev * 4, X, (size_t)ld * 4, width, height, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(dX);
        set_error("host->device copy failed: %s", cudaGetErrorString(e));
        return IFB_ECUDA;
    }
    ld = ld_dev;
    int rc = ifb_fit_device(device, dX, n_rows, d, ld, layout, prm, out, nullptr);
    cudaFree(dX);
    return rc;
}
