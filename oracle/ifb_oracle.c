/*
 * ifb_oracle.c -- CPU ORACLE for the isolation-forest hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the reference algorithm (linkedin/isolation-forest @ 10b5f0a).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (libifb200.so) never links, loads or calls anything in oracle/.
 *
 * Parity pin: checked in tests/test_oracle_golden.py against (a) the reference's f32 known answers
 * (IFT/core/UtilsTest.scala:12-16, IFT/IsolationTreeTest.scala:27-42,
 * IFT/extended/ExtendedIsolationTreeTest.scala:32-82), (b) the 11,183 reference-computed scores in
 * isolation-forest-onnx/test/resources/savedIsolationForestModel/mammographyModel/
 * mammographyOutlierScores.csv, (c) the stored exact-quantile thresholds of both saved models under
 * isolation-forest/src/test/resources.  Fit has NO stream-level pin in the reference (its tests are
 * statistical only); the fit restatement below follows the reference source line by
 * line including java.util.Random, and is pinned only through the reference's statistical bands.
 *
 * Shorthand in citations:  IF/ = isolation-forest/src/main/scala/com/linkedin/relevance/isolationforest/
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction anywhere, the JVM never
 * fuses a*b+c).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IFBO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ */
/* a1: Utils.avgPathLength  -- IF/core/Utils.scala:74-92                                            */
/* ------------------------------------------------------------------------------------------------ */
IFBO_API float ifbo_avg_path_length(int64_t num_instances) {
    if (num_instances <= 1) return 0.0f;
    const float euler = 0.5772156649f;              /* Utils.scala:74 */
    float nf = (float)num_instances;                /* numInstances.toFloat */
    float lg = (float)log((double)(nf - 1.0f));     /* math.log(Float->Double).toFloat */
    float a = 2.0f * (lg + euler);
    float b = (2.0f * (nf - 1.0f)) / nf;
    return a - b;
}

/* ------------------------------------------------------------------------------------------------ */
/* Forest tables: pre-order node rows exactly as persisted by the reference                          */
/* (IF/IsolationForestModelReadWrite.scala:60-67, IF/extended/...ReadWrite.scala:59-67):             */
/* leaves carry left=right=-1; internal nodes carry num_instances=-1.                                */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t num_trees;
    const int32_t *node_off;       /* [T+1] first node row of each tree */
    const int32_t *left, *right;   /* node ids local to the tree, -1 at leaves */
    const int64_t *num_instances;  /* leaf size, -1 at internal nodes */
    /* standard */
    const int32_t *feature;        /* splitAttribute */
    const double *threshold;       /* splitValue */
    /* extended (CSR hyperplanes) */
    const double *offset;
    const int64_t *hp_off;         /* [nodes+1] */
    const int32_t *hp_idx;
    const float *hp_w;
} ifbo_forest;

/* a3: IsolationTree.pathLength -- IF/IsolationTree.scala:196-230.  x is one row, stride xs. */
static inline float walk_standard(const ifbo_forest *f, int t, const float *x, int64_t xs, int32_t *depth_out) {
    int32_t base = f->node_off[t];
    int32_t node = 0;
    float cur = 0.0f;                      /* currentPathLength: Float, += 1 per level */
    int32_t depth = 0;
    while (f->left[base + node] != -1) {
        int32_t a = f->feature[base + node];
        double sv = f->threshold[base + node];
        /* Float widened to Double, strict '<' (IsolationTree.scala:222); NaN -> right */
        if ((double)x[(int64_t)a * xs] < sv) node = f->left[base + node];
        else node = f->right[base + node];
        cur = cur + 1.0f;
        depth++;
    }
    *depth_out = depth;
    return cur + ifbo_avg_path_length(f->num_instances[base + node]);   /* :218 */
}

/* a5: SplitHyperplane.dot -- IF/extended/ExtendedUtils.scala:36-44: Float*Float -> Float, += Double */
static inline double hp_dot(const ifbo_forest *f, int64_t gnode, const float *x, int64_t xs) {
    double sum = 0.0;
    for (int64_t i = f->hp_off[gnode]; i < f->hp_off[gnode + 1]; i++) {
        float p = f->hp_w[i] * x[(int64_t)f->hp_idx[i] * xs];
        sum += (double)p;
    }
    return sum;
}

/* a7: ExtendedIsolationTree.pathLength -- IF/extended/ExtendedIsolationTree.scala:283-320 */
static inline float walk_extended(const ifbo_forest *f, int t, const float *x, int64_t xs, int32_t *depth_out) {
    int32_t base = f->node_off[t];
    int32_t node = 0;
    float cur = 0.0f;
    int32_t depth = 0;
    while (f->left[base + node] != -1) {
        double dp = hp_dot(f, (int64_t)base + node, x, xs);
        if (dp < f->offset[base + node]) node = f->left[base + node];     /* :310 strict '<' */
        else node = f->right[base + node];
        cur = cur + 1.0f;
        depth++;
    }
    *depth_out = depth;
    return cur + ifbo_avg_path_length(f->num_instances[base + node]);
}

typedef struct {
    const ifbo_forest *f;
    int extended;
    const float *X;
    int64_t row_stride, col_stride;   /* element strides: x[r][c] = X[r*row_stride + c*col_stride] */
    int64_t r0, r1;
    int32_t num_samples;
    double *scores;
    int32_t *depth_sum;
    float *path_sum;
} score_job;

/* a4/a8: transform UDF body -- IF/IsolationForestModel.scala:128-139,
 *                              IF/extended/ExtendedIsolationForestModel.scala:110-120 */
static void *score_range(void *arg) {
    score_job *j = (score_job *)arg;
    const ifbo_forest *f = j->f;
    const float avg_path = ifbo_avg_path_length(j->num_samples);
    const int T = f->num_trees;
    for (int64_t r = j->r0; r < j->r1; r++) {
        const float *x = j->X + r * j->row_stride;
        float s = 0.0f;            /* Array[Float].sum: foldLeft from 0f */
        int32_t ds = 0;
        for (int t = 0; t < T; t++) {
            int32_t dep;
            float pl = j->extended ? walk_extended(f, t, x, j->col_stride, &dep)
                                   : walk_standard(f, t, x, j->col_stride, &dep);
            s = s + pl;
            ds += dep;
        }
        float e = s / (float)T;                       /* Float / Int */
        float z = (-e) / avg_path;                    /* Float / Float */
        j->scores[r] = pow(2.0, (double)z);           /* Math.pow(2, Float->Double) */
        if (j->depth_sum) j->depth_sum[r] = ds;
        if (j->path_sum) j->path_sum[r] = s;
    }
    return NULL;
}

IFBO_API int ifbo_score(const ifbo_forest *f, int extended, const float *X, int64_t n_rows,
                        int64_t row_stride, int64_t col_stride, int32_t num_samples, int n_threads,
                        double *scores, int32_t *depth_sum, float *path_sum) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    score_job jobs[256];
    int64_t chunk = (n_rows + n_threads - 1) / n_threads;
    int used = 0;
    for (int i = 0; i < n_threads; i++) {
        int64_t r0 = i * chunk, r1 = r0 + chunk;
        if (r0 >= n_rows) break;
        if (r1 > n_rows) r1 = n_rows;
        jobs[i] = (score_job){f, extended, X, row_stride, col_stride, r0, r1, num_samples, scores, depth_sum, path_sum};
        used++;
    }
    if (used == 1) { score_range(&jobs[0]); return 0; }
    for (int i = 0; i < used; i++) pthread_create(&th[i], NULL, score_range, &jobs[i]);
    for (int i = 0; i < used; i++) pthread_join(th[i], NULL);
    return 0;
}

/* Single-tree path length (KAT entry point): returns pathLength(x) of tree t. */
IFBO_API float ifbo_path_length(const ifbo_forest *f, int extended, int t, const float *x) {
    int32_t dep;
    return extended ? walk_extended(f, t, x, 1, &dep) : walk_standard(f, t, x, 1, &dep);
}

/* ------------------------------------------------------------------------------------------------ */
/* java.util.Random (the generator behind scala.util.Random) -- algorithm as specified in the       */
/* java.util.Random Javadoc: 48-bit LCG, next(bits), nextInt(bound), nextDouble, nextGaussian.      */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t seed;
    int have_next;
    double next_gauss;
} jrandom;

static void jr_init(jrandom *r, int64_t seed) {
    r->seed = ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
    r->have_next = 0;
    r->next_gauss = 0.0;
}
static inline int32_t jr_next(jrandom *r, int bits) {
    r->seed = (r->seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int32_t)(int64_t)(r->seed >> (48 - bits));   /* (int)(seed >>> (48-bits)) */
}
static int32_t jr_next_int(jrandom *r, int32_t bound) {
    int32_t rr = jr_next(r, 31);
    int32_t m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)rr) >> 31);
    for (int32_t u = rr;; u = jr_next(r, 31)) {
        rr = u % bound;
        /* u - rr + m < 0 in wrapping int arithmetic */
        if ((int32_t)((uint32_t)u - (uint32_t)rr + (uint32_t)m) >= 0) break;
    }
    return rr;
}
static double jr_next_double(jrandom *r) {
    int64_t hi = (int64_t)jr_next(r, 26);
    int64_t lo = (int64_t)jr_next(r, 27);
    return (double)((hi << 27) + lo) * 0x1.0p-53;
}

/* StrictMath.log == fdlibm __ieee754_log (public algorithm, fdlibm 5.3 e_log.c), restated so that the
 * Gaussian draws do not depend on the host libm. */
static double fdlibm_log(double x) {
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01,
                        Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                        Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01,
                        Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    uint64_t bits;
    memcpy(&bits, &x, 8);
    int32_t hx = (int32_t)(bits >> 32);
    uint32_t lx = (uint32_t)bits;
    int32_t k = 0, i, j;
    double hfsq, f, s, z, R, w, t1, t2, dk;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -INFINITY;
        if (hx < 0) return NAN;
        k -= 54;
        x *= two54;
        memcpy(&bits, &x, 8);
        hx = (int32_t)(bits >> 32);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    bits = (bits & 0xffffffffULL) | ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32);
    memcpy(&x, &bits, 8);
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
IFBO_API double ifbo_fdlibm_log(double x) { return fdlibm_log(x); }

static double jr_next_gaussian(jrandom *r) {
    if (r->have_next) {
        r->have_next = 0;
        return r->next_gauss;
    }
    double v1, v2, s;
    do {
        v1 = 2 * jr_next_double(r) - 1;
        v2 = 2 * jr_next_double(r) - 1;
        s = v1 * v1 + v2 * v2;
    } while (s >= 1 || s == 0);
    double multiplier = sqrt(-2 * fdlibm_log(s) / s);
    r->next_gauss = v2 * multiplier;
    r->have_next = 1;
    return v1 * multiplier;
}

/* KAT hooks for the generator (tests compare against values published in the JDK documentation /
 * widely known sequences, e.g. new Random(42).nextInt() == -1170105035). */
IFBO_API void ifbo_jrandom_kat(int64_t seed, int32_t *ints3, double *dbl1, double *gauss2, int32_t *bounded3) {
    jrandom r;
    jr_init(&r, seed);
    for (int i = 0; i < 3; i++) ints3[i] = jr_next(&r, 32);
    jr_init(&r, seed);
    *dbl1 = jr_next_double(&r);
    jr_init(&r, seed);
    gauss2[0] = jr_next_gaussian(&r);
    gauss2[1] = jr_next_gaussian(&r);
    jr_init(&r, seed);
    for (int i = 0; i < 3; i++) bounded3[i] = jr_next_int(&r, 10);
}

/* scala.util.Random.shuffle: for (n <- len to 2 by -1) { k = nextInt(n); swap(n-1, k) } */
static void scala_shuffle(jrandom *r, int32_t *a, int32_t len) {
    for (int32_t n = len; n >= 2; n--) {
        int32_t k = jr_next_int(r, n);
        int32_t tmp = a[n - 1];
        a[n - 1] = a[k];
        a[k] = tmp;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Tree output buffers (pre-order).                                                                  */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t cap, n;                 /* node capacity / count */
    int32_t *left, *right, *feature;
    double *threshold;
    int64_t *num_instances;
    /* extended */
    int32_t k;                      /* non-zeros per hyperplane */
    double *offset;
    int32_t *hp_idx;                /* [cap*k] */
    float *hp_w;                    /* [cap*k] */
    int32_t *hp_len;                /* [cap] 0 at leaves */
} tree_out;

typedef struct {
    const float *data;   /* n x d row-major sample matrix of this tree */
    int32_t d;
    jrandom rnd;
    const int32_t *feat; /* featureIndices */
    int32_t n_feat;
    int32_t height_limit;
    int32_t ext_level;
    tree_out *out;
} build_ctx;

static int32_t emit_leaf_std(tree_out *o, int64_t n) {
    int32_t id = o->n++;
    o->left[id] = -1; o->right[id] = -1; o->feature[id] = -1; o->threshold[id] = 0.0;
    o->num_instances[id] = n;
    return id;
}

/* a9: generateIsolationTreeInternal -- IF/IsolationTree.scala:107-180 (pre-order emission). */
static int32_t build_std(build_ctx *c, int32_t *rows, int32_t n, int32_t height) {
    /* getFeatureToSplit (:124-150) runs BEFORE the stop test (:152-156) and consumes draws */
    int32_t avail[c->n_feat > 0 ? c->n_feat : 1];
    int32_t n_avail = c->n_feat;
    memcpy(avail, c->feat, sizeof(int32_t) * (size_t)c->n_feat);
    int32_t feature_index = -1;
    double split_value = 0.0;
    while (feature_index == -1 && n_avail > 0) {
        int32_t pick = jr_next_int(&c->rnd, n_avail);
        int32_t trial = avail[pick];
        memmove(&avail[pick], &avail[pick + 1], sizeof(int32_t) * (size_t)(n_avail - pick - 1));
        n_avail--;
        /* featureValues.min / .max over the node (Float), then .toDouble.  Scala's `min` over an EMPTY
         * array throws; the reference cannot reach that (size-0 nodes never occur in standard IF:
         * both filter sides are non-empty whenever min != max). */
        if (n == 0) continue;
        float mn = c->data[(int64_t)rows[0] * c->d + trial], mx = mn;
        for (int32_t i = 1; i < n; i++) {
            float v = c->data[(int64_t)rows[i] * c->d + trial];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        double dmn = (double)mn, dmx = (double)mx;
        if (dmn != dmx) {
            feature_index = trial;
            split_value = (dmx - dmn) * jr_next_double(&c->rnd) + dmn;   /* :145-146 */
        }
    }
    if (feature_index == -1 || height >= c->height_limit || n <= 1) return emit_leaf_std(c->out, n);
    tree_out *o = c->out;
    int32_t id = o->n++;
    int32_t *l = (int32_t *)malloc(sizeof(int32_t) * (size_t)n * 2);
    int32_t *r = l + n;
    int32_t nl = 0, nr = 0;
    for (int32_t i = 0; i < n; i++) {
        double v = (double)c->data[(int64_t)rows[i] * c->d + feature_index];
        if (v < split_value) l[nl++] = rows[i];          /* :158 */
        if (v >= split_value) r[nr++] = rows[i];         /* :159 (NaN rows vanish, as in the reference) */
    }
    o->feature[id] = feature_index;
    o->threshold[id] = split_value;
    o->num_instances[id] = -1;
    o->left[id] = build_std(c, l, nl, height + 1);
    o->right[id] = build_std(c, r, nr, height + 1);
    free(l);
    return id;
}

static int32_t emit_leaf_ext(tree_out *o, int64_t n) {
    int32_t id = o->n++;
    o->left[id] = -1; o->right[id] = -1; o->offset[id] = 0.0; o->hp_len[id] = 0;
    o->num_instances[id] = n;
    return id;
}

/* a10: generateExtendedIsolationTreeInternal -- IF/extended/ExtendedIsolationTree.scala:139-260 */
static int32_t build_ext(build_ctx *c, int32_t *rows, int32_t n, int32_t height) {
    tree_out *o = c->out;
    if (height >= c->height_limit || n <= 1) return emit_leaf_ext(o, n);      /* :152-153 */
    int32_t dim = c->n_feat;
    int32_t nnz = c->ext_level + 1 < dim ? c->ext_level + 1 : dim;            /* :157 */
    int32_t *perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)dim);
    for (int32_t i = 0; i < dim; i++) perm[i] = i;
    scala_shuffle(&c->rnd, perm, dim);                                         /* :160 */
    int32_t sparse_idx[nnz];
    double raw[nnz];
    for (int32_t i = 0; i < nnz; i++) {
        sparse_idx[i] = c->feat[perm[i]];                                      /* :168 */
        raw[i] = jr_next_gaussian(&c->rnd);                                    /* :169 */
    }
    free(perm);
    double sq = 0.0;
    for (int32_t i = 0; i < nnz; i++) sq += raw[i] * raw[i];                   /* :174-179 */
    double norm = sqrt(sq);
    if (norm == 0) return emit_leaf_ext(o, n);                                 /* :183-184 */
    float w[nnz];
    for (int32_t i = 0; i < nnz; i++) w[i] = (float)(raw[i] / norm);           /* :190-195 */
    double split_offset = 0.0;
    for (int32_t k = 0; k < nnz; k++) {                                        /* :201-217 */
        int32_t j = sparse_idx[k];
        double mn = INFINITY, mx = -INFINITY;
        for (int32_t r = 0; r < n; r++) {
            double v = (double)c->data[(int64_t)rows[r] * c->d + j];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        double iv = (mn == mx) ? mn : mn + jr_next_double(&c->rnd) * (mx - mn);
        split_offset += (double)w[k] * iv;        /* Float * Double -> Double */
    }
    /* canonical sort by index (:220-226); indices are distinct so a stable insertion sort suffices */
    for (int32_t a = 1; a < nnz; a++) {
        int32_t ki = sparse_idx[a];
        float kw = w[a];
        int32_t b = a - 1;
        while (b >= 0 && sparse_idx[b] > ki) { sparse_idx[b + 1] = sparse_idx[b]; w[b + 1] = w[b]; b--; }
        sparse_idx[b + 1] = ki;
        w[b + 1] = kw;
    }
    int32_t id = o->n++;
    o->hp_len[id] = nnz;
    memcpy(&o->hp_idx[(int64_t)id * o->k], sparse_idx, sizeof(int32_t) * (size_t)nnz);
    memcpy(&o->hp_w[(int64_t)id * o->k], w, sizeof(float) * (size_t)nnz);
    o->offset[id] = split_offset;
    o->num_instances[id] = -1;
    int32_t *l = (int32_t *)malloc(sizeof(int32_t) * (size_t)n * 2);
    int32_t *r = l + n;
    int32_t nl = 0, nr = 0;
    for (int32_t i = 0; i < n; i++) {                                          /* :230-232 */
        const float *x = c->data + (int64_t)rows[i] * c->d;
        double sum = 0.0;
        for (int32_t q = 0; q < nnz; q++) {
            float p = w[q] * x[sparse_idx[q]];
            sum += (double)p;
        }
        if (sum < split_offset) l[nl++] = rows[i];
        else r[nr++] = rows[i];
    }
    o->left[id] = build_ext(c, l, nl, height + 1);
    o->right[id] = build_ext(c, r, nr, height + 1);
    free(l);
    return id;
}

/* heightLimit = ceil(log10(n)/log10(2)) -- IF/IsolationTree.scala:60-61 */
IFBO_API int32_t ifbo_height_limit(int32_t n) { return (int32_t)ceil(log10((double)n) / log10(2.0)); }

/* Fit ONE tree on an explicit sample matrix (n x d row-major), mirroring
 * IsolationTree.fit(data, randomSeed, featureIndices) (IF/IsolationTree.scala:53-66) or
 * ExtendedIsolationTree.fit(..., extensionLevel) (IF/extended/ExtendedIsolationTree.scala:67-92).
 * ext_level < 0 selects the standard builder.  Output arrays must hold 2n-1 nodes (standard) or
 * 2^(heightLimit+1)-1 nodes (extended: empty children are legal).  Returns node count. */
IFBO_API int32_t ifbo_fit_tree(const float *data, int32_t n, int32_t d, int64_t seed, const int32_t *feat,
                               int32_t n_feat, int32_t ext_level, int32_t cap, int32_t *left, int32_t *right,
                               int32_t *feature, double *threshold, int64_t *num_instances, double *offset,
                               int32_t *hp_len, int32_t *hp_idx, float *hp_w, int32_t k) {
    tree_out o = {cap, 0, left, right, feature, threshold, num_instances, k, offset, hp_idx, hp_w, hp_len};
    build_ctx c;
    c.data = data; c.d = d; c.feat = feat; c.n_feat = n_feat; c.ext_level = ext_level; c.out = &o;
    c.height_limit = ifbo_height_limit(n);
    jr_init(&c.rnd, seed);
    int32_t *rows = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int32_t i = 0; i < n; i++) rows[i] = i;
    if (ext_level < 0) build_std(&c, rows, n, 0);
    else build_ext(&c, rows, n, 0);
    free(rows);
    return o.n;
}

/* ------------------------------------------------------------------------------------------------ */
/* Per-tree sampling (ENGINE-DEFINED, see DESIGN.md "fit: sampling contract").                        */
/* The reference bags with per-(row,tree) Bernoulli/Poisson draws at rate (n+7*sqrt(n))/N, hash-      */
/* partitions by tree id, shuffles each partition and keeps the first n rows                          */
/* (IF/core/SharedTrainLogic.scala:99-153,287; IF/core/BaggedPoint.scala:114-217): a uniform random   */
/* n-subset in random order (without replacement), or n iid draws (bootstrap).  Its stream depends on */
/* Spark partitioning and commons-math3 and is unpinned.  The engine draws the same     */
/* distribution directly:  rnd = java.util.Random(treeSeed);                                          */
/*   !bootstrap: partial Fisher-Yates over the virtual identity permutation of [0,N):                 */
/*               for i in 0..n-1: j = i + bounded(N-i); swap(a[i],a[j]); sample[i] = a[i]             */
/*    bootstrap: sample[i] = bounded(N)                                                               */
/*   bounded(m) = nextInt(m) for m < 2^31, else floorMod-free ((next(31)<<31 | next(31)) % m)         */
/* then (same rnd, as SharedTrainLogic.scala:300-304) featureIndices =                                */
/*   shuffle(0 until d).take(numFeatures).sorted; the tree builder then restarts from                 */
/*   new Random(treeSeed) exactly as IsolationTree.fit does (IF/IsolationTree.scala:63).              */
/* treeSeed = randomSeed + 2*(P+1) + treeId  (IF/IsolationForest.scala:76-78,                         */
/* SharedTrainLogic.scala:283; P = input partition count, an engine parameter, default 1).            */
/* ------------------------------------------------------------------------------------------------ */
static int64_t jr_bounded(jrandom *r, int64_t m) {
    if (m < 0x7fffffffLL) return jr_next_int(r, (int32_t)m);
    uint64_t hi = (uint64_t)jr_next(r, 31), lo = (uint64_t)jr_next(r, 31);
    return (int64_t)(((hi << 31) | lo) % (uint64_t)m);
}

IFBO_API void ifbo_sample_tree(int64_t tree_seed, int64_t N, int32_t n, int32_t d, int32_t num_features,
                               int bootstrap, int64_t *rows_out, int32_t *feat_out) {
    jrandom r;
    jr_init(&r, tree_seed);
    if (bootstrap) {
        for (int32_t i = 0; i < n; i++) rows_out[i] = jr_bounded(&r, N);
    } else {
        /* open-addressing map of displaced positions: key -> value, <= 2n entries */
        int32_t cap = 1;
        while (cap < 4 * n) cap <<= 1;
        int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap * 2);
        int64_t *vals = keys + cap;
        for (int32_t i = 0; i < cap; i++) keys[i] = -1;
#define MAP_SLOT(key, slot)                                                     \
    do {                                                                        \
        slot = (int32_t)(((uint64_t)(key) * 0x9E3779B97F4A7C15ULL) >> 40) & (cap - 1); \
        while (keys[slot] != -1 && keys[slot] != (key)) slot = (slot + 1) & (cap - 1); \
    } while (0)
        for (int32_t i = 0; i < n; i++) {
            int64_t j = (int64_t)i + jr_bounded(&r, N - i);
            int32_t sj, si;
            MAP_SLOT(j, sj);
            int64_t aj = keys[sj] == j ? vals[sj] : j;
            MAP_SLOT((int64_t)i, si);
            int64_t ai = keys[si] == (int64_t)i ? vals[si] : (int64_t)i;
            rows_out[i] = aj;                 /* a[i] after the swap */
            MAP_SLOT(j, sj);
            keys[sj] = j;
            vals[sj] = ai;                    /* a[j] = old a[i] */
        }
#undef MAP_SLOT
        free(keys);
    }
    int32_t *perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)d);
    for (int32_t i = 0; i < d; i++) perm[i] = i;
    scala_shuffle(&r, perm, d);
    /* take(numFeatures).sorted */
    for (int32_t a = 0; a < num_features; a++) feat_out[a] = perm[a];
    for (int32_t a = 1; a < num_features; a++) {
        int32_t kv = feat_out[a], b = a - 1;
        while (b >= 0 && feat_out[b] > kv) { feat_out[b + 1] = feat_out[b]; b--; }
        feat_out[b + 1] = kv;
    }
    free(perm);
}

/* Fit a whole forest on X (element strides as in ifbo_score).  Node tables are written tree after
 * tree at stride `cap` nodes per tree; n_nodes[t] receives each tree's node count.  ext_level < 0 =>
 * standard.  k = hyperplane width (min(ext_level+1, num_features)) for the extended layout. */
IFBO_API int ifbo_fit_forest(const float *X, int64_t N, int32_t d, int64_t row_stride, int64_t col_stride,
                             int32_t num_trees, int32_t n, int32_t num_features, int bootstrap,
                             int64_t random_seed, int32_t num_partitions, int32_t ext_level, int32_t cap,
                             int32_t *n_nodes, int32_t *left, int32_t *right, int32_t *feature,
                             double *threshold, int64_t *num_instances, double *offset, int32_t *hp_len,
                             int32_t *hp_idx, float *hp_w, int32_t k, int64_t *sample_rows_out) {
    float *data = (float *)malloc(sizeof(float) * (size_t)n * (size_t)d);
    int64_t *rows = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int32_t *feat = (int32_t *)malloc(sizeof(int32_t) * (size_t)d);
    for (int32_t t = 0; t < num_trees; t++) {
        int64_t seed = random_seed + 2 * ((int64_t)num_partitions + 1) + t;
        ifbo_sample_tree(seed, N, n, d, num_features, bootstrap, rows, feat);
        if (sample_rows_out) memcpy(sample_rows_out + (int64_t)t * n, rows, sizeof(int64_t) * (size_t)n);
        for (int32_t i = 0; i < n; i++)
            for (int32_t c = 0; c < d; c++) data[(int64_t)i * d + c] = X[rows[i] * row_stride + (int64_t)c * col_stride];
        int64_t o = (int64_t)t * cap;
        n_nodes[t] = ifbo_fit_tree(data, n, d, seed, feat, num_features, ext_level, cap, left + o, right + o,
                                   feature ? feature + o : NULL, threshold ? threshold + o : NULL,
                                   num_instances + o, offset ? offset + o : NULL, hp_len ? hp_len + o : NULL,
                                   hp_idx ? hp_idx + o * k : NULL, hp_w ? hp_w + o * k : NULL, k);
    }
    free(data); free(rows); free(feat);
    return 0;
}
