/*
 * orpcd_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C float64 restatement of the Open3D 0.18.0 algorithms that the
 * reference (dpolimeni/Multi-scale-pointcloud-registration, package or_pcd)
 * calls on its registration hot path.  Open3D is an un-vendored third-party
 * dependency of the reference (pinned open3d==0.18.0 at setup.py:26) and is not
 * installable here, so this file restates its PUBLISHED algorithms; SURVEY.md
 * Appendix A is the specification it follows.
 *
 *   PARITY UNPINNED (the arithmetic of the five Open3D entry points below):
 *   the reference ships no tests and no golden vectors for this path, and
 *   Open3D itself cannot be run here.  This file is pinned only against
 *   independent libraries (scipy cKDTree, numpy eigh/inv/sqrtm) and known
 *   answers.  What IS pinned to the reference itself is everything around
 *   these entry points (oracle.py): golden vectors of the reference's own
 *   classes -- RadiusScaler, RandomDownsampler, FarthestDownsampler, the
 *   Aligner (control flow, RNG consumption, transfrom, refine_registration),
 *   GeneralizedICP, FastGlobalOptimizer, SOR and VoxelDownsampler, the last
 *   five run unmodified over an open3d stub backed by this oracle
 *   (tests/golden/make_golden*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (multi-scale-pointcloud-registration_b200/) never does.
 *
 * Reference call sites restated here:
 *   K1  sor.py:69-71                      -> oracle_sor()
 *   K2  generalizedICP.py:59-70           -> oracle_gicp()
 *   K3  fastGlobalOptimizer.py:118-127    -> oracle_normals(hybrid)
 *   K4  fastGlobalOptimizer.py:130-143    -> oracle_fpfh()
 *   K5  fastGlobalOptimizer.py:158-174    -> oracle_fgr()
 *
 * Conventions fixed by BASELINE.json north_star: neighbour ties break towards
 * the lowest index; reductions run in index order (Open3D's OpenMP merge order
 * is non-deterministic).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* KD-tree (exact, any dimension), nanoflann-equivalent results             */
/* (KDTreeFlann, leaf_max_size 15, L2 metric, sorted output) */
/* ------------------------------------------------------------------------ */

typedef struct {
    int dim;
    int n;
    const double *pts; /* n x dim, row-major (not owned) */
    int *perm;         /* point indices, grouped by leaf */
    int n_nodes;
    int cap_nodes;
    int *left;         /* child node or -1 for leaf */
    int *right;
    int *lo;           /* leaf: [lo,hi) in perm */
    int *hi;
    int *sdim;
    double *sval_lo;   /* max of left subtree on sdim  */
    double *sval_hi;   /* min of right subtree on sdim */
} KDTree;

#define KD_LEAF 15

/* squared L2 distance in nanoflann's L2_Adaptor::evalMetric order:
 * groups of four summed left to right then added to the accumulator, then a
 * one-element tail loop.  dim 3 => ((dx*dx)+dy*dy)+dz*dz.                  */
static inline double sqdist(const double *a, const double *b, int dim)
{
    double result = 0.0;
    int i = 0;
    for (; i + 3 < dim; i += 4) {
        const double d0 = a[i] - b[i];
        const double d1 = a[i + 1] - b[i + 1];
        const double d2 = a[i + 2] - b[i + 2];
        const double d3 = a[i + 3] - b[i + 3];
        result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    for (; i < dim; ++i) {
        const double d0 = a[i] - b[i];
        result += d0 * d0;
    }
    return result;
}

static int kd_new_node(KDTree *t)
{
    if (t->n_nodes == t->cap_nodes) {
        int c = t->cap_nodes * 2;
        t->left = (int *)realloc(t->left, c * sizeof(int));
        t->right = (int *)realloc(t->right, c * sizeof(int));
        t->lo = (int *)realloc(t->lo, c * sizeof(int));
        t->hi = (int *)realloc(t->hi, c * sizeof(int));
        t->sdim = (int *)realloc(t->sdim, c * sizeof(int));
        t->sval_lo = (double *)realloc(t->sval_lo, c * sizeof(double));
        t->sval_hi = (double *)realloc(t->sval_hi, c * sizeof(double));
        t->cap_nodes = c;
    }
    return t->n_nodes++;
}

static void kd_nth(KDTree *t, int lo, int hi, int nth, int d)
{
    /* quickselect on perm[lo,hi) by coordinate d (ties by index) */
    const double *p = t->pts;
    const int dim = t->dim;
    while (hi - lo > 1) {
        int mid = lo + (hi - lo) / 2;
        int pi = t->perm[mid];
        double pv = p[(size_t)pi * dim + d];
        int i = lo, j = hi - 1;
        while (i <= j) {
            while (1) {
                int a = t->perm[i];
                double av = p[(size_t)a * dim + d];
                if (av < pv || (av == pv && a < pi)) ++i; else break;
            }
            while (1) {
                int a = t->perm[j];
                double av = p[(size_t)a * dim + d];
                if (av > pv || (av == pv && a > pi)) --j; else break;
            }
            if (i <= j) {
                int tmp = t->perm[i]; t->perm[i] = t->perm[j]; t->perm[j] = tmp;
                ++i; --j;
            }
        }
        if (nth <= j) hi = j + 1;
        else if (nth >= i) lo = i;
        else return;
    }
}

static int kd_build_rec(KDTree *t, int lo, int hi)
{
    int node = kd_new_node(t);
    t->lo[node] = lo;
    t->hi[node] = hi;
    t->left[node] = t->right[node] = -1;
    if (hi - lo <= KD_LEAF) return node;
    const int dim = t->dim;
    int best = 0;
    double best_span = -1.0;
    for (int d = 0; d < dim; ++d) {
        double mn = INFINITY, mx = -INFINITY;
        for (int i = lo; i < hi; ++i) {
            double v = t->pts[(size_t)t->perm[i] * dim + d];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        if (mx - mn > best_span) { best_span = mx - mn; best = d; }
    }
    if (!(best_span > 0.0)) return node; /* all points identical: keep as leaf */
    int mid = lo + (hi - lo) / 2;
    kd_nth(t, lo, hi, mid, best);
    double lmax = -INFINITY, rmin = INFINITY;
    for (int i = lo; i < mid; ++i) {
        double v = t->pts[(size_t)t->perm[i] * dim + best];
        if (v > lmax) lmax = v;
    }
    for (int i = mid; i < hi; ++i) {
        double v = t->pts[(size_t)t->perm[i] * dim + best];
        if (v < rmin) rmin = v;
    }
    t->sdim[node] = best;
    t->sval_lo[node] = lmax;
    t->sval_hi[node] = rmin;
    int l = kd_build_rec(t, lo, mid);
    int r = kd_build_rec(t, mid, hi);
    t->left[node] = l;
    t->right[node] = r;
    return node;
}

static KDTree *kd_build(const double *pts, int n, int dim)
{
    KDTree *t = (KDTree *)calloc(1, sizeof(KDTree));
    t->dim = dim; t->n = n; t->pts = pts;
    t->perm = (int *)malloc((n > 0 ? n : 1) * sizeof(int));
    for (int i = 0; i < n; ++i) t->perm[i] = i;
    t->cap_nodes = 64;
    t->left = (int *)malloc(64 * sizeof(int));
    t->right = (int *)malloc(64 * sizeof(int));
    t->lo = (int *)malloc(64 * sizeof(int));
    t->hi = (int *)malloc(64 * sizeof(int));
    t->sdim = (int *)malloc(64 * sizeof(int));
    t->sval_lo = (double *)malloc(64 * sizeof(double));
    t->sval_hi = (double *)malloc(64 * sizeof(double));
    if (n > 0) kd_build_rec(t, 0, n);
    return t;
}

static void kd_free(KDTree *t)
{
    if (!t) return;
    free(t->perm); free(t->left); free(t->right); free(t->lo); free(t->hi);
    free(t->sdim); free(t->sval_lo); free(t->sval_hi); free(t);
}

/* result set ordered lexicographically by (d2, index): traversal-independent */
typedef struct { int k; int cnt; int *idx; double *d2; } KSet;

static inline int kset_better(double d, int i, double d_o, int i_o)
{
    return d < d_o || (d == d_o && i < i_o);
}

static inline void kset_add(KSet *s, double d, int i)
{
    if (s->cnt == s->k && !kset_better(d, i, s->d2[s->k - 1], s->idx[s->k - 1])) return;
    int pos = (s->cnt < s->k) ? s->cnt++ : s->k - 1;
    while (pos > 0 && kset_better(d, i, s->d2[pos - 1], s->idx[pos - 1])) {
        s->d2[pos] = s->d2[pos - 1];
        s->idx[pos] = s->idx[pos - 1];
        --pos;
    }
    s->d2[pos] = d;
    s->idx[pos] = i;
}

static void kd_search_rec(const KDTree *t, int node, const double *q, KSet *s)
{
    if (t->left[node] < 0) {
        for (int i = t->lo[node]; i < t->hi[node]; ++i) {
            int pi = t->perm[i];
            kset_add(s, sqdist(q, t->pts + (size_t)pi * t->dim, t->dim), pi);
        }
        return;
    }
    const int d = t->sdim[node];
    const double v = q[d];
    /* sval_lo <= sval_hi (split by sorted order); descend into the nearer box */
    int first, second;
    double gap;
    if (v <= 0.5 * (t->sval_lo[node] + t->sval_hi[node])) {
        first = t->left[node]; second = t->right[node]; gap = t->sval_hi[node] - v;
    } else {
        first = t->right[node]; second = t->left[node]; gap = v - t->sval_lo[node];
    }
    kd_search_rec(t, first, q, s);
    /* visit the far side unless it provably cannot hold a (d2,idx)-better
     * point: one squared coordinate gap is a lower bound of the rounded d2
     * (monotone rounding), so prune only on a strict '>'.                  */
    if (s->cnt < s->k) {
        kd_search_rec(t, second, q, s);
    } else {
        double lb = gap > 0.0 ? gap * gap : 0.0;
        if (!(lb > s->d2[s->k - 1])) kd_search_rec(t, second, q, s);
    }
}

static int kd_knn(const KDTree *t, const double *q, int k, int *idx, double *d2)
{
    KSet s; s.k = k; s.cnt = 0; s.idx = idx; s.d2 = d2;
    if (t->n > 0 && k > 0) kd_search_rec(t, 0, q, &s);
    return s.cnt;
}

/* Brute-force twin of kd_knn (used to validate the tree and by small tests). */
static int brute_knn(const double *pts, int n, int dim, const double *q, int k, int *idx, double *d2)
{
    KSet s; s.k = k; s.cnt = 0; s.idx = idx; s.d2 = d2;
    for (int i = 0; i < n; ++i) kset_add(&s, sqdist(q, pts + (size_t)i * dim, dim), i);
    return s.cnt;
}

/* KDTreeFlann::SearchKNN for a batch of queries.  out_idx/out_d2: m x k,
 * padded with -1 / +inf when fewer than k points exist.  method 0 = kd-tree,
 * 1 = brute force.  Returns 0.                                             */
ORACLE_API int oracle_knn(const double *pts, int n, int dim, const double *queries, int m,
                          int k, int method, int *out_idx, double *out_d2)
{
    KDTree *t = method == 0 ? kd_build(pts, n, dim) : NULL;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < m; ++i) {
        int *oi = out_idx + (size_t)i * k;
        double *od = out_d2 + (size_t)i * k;
        int c = method == 0 ? kd_knn(t, queries + (size_t)i * dim, k, oi, od)
                            : brute_knn(pts, n, dim, queries + (size_t)i * dim, k, oi, od);
        for (; c < k; ++c) { oi[c] = -1; od[c] = INFINITY; }
    }
    kd_free(t);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* K1: PointCloud::RemoveStatisticalOutliers (sor.py:69-71)   */
/* ------------------------------------------------------------------------ */
ORACLE_API int oracle_sor(const double *pts, int n, int nb_neighbors, double std_ratio,
                          unsigned char *mask, double *avg_out, double *stats /* mean,std,thr */)
{
    if (n <= 0 || nb_neighbors < 1 || std_ratio <= 0) return -1;
    KDTree *t = kd_build(pts, n, 3);
    double *avg = (double *)malloc(n * sizeof(double));
#pragma omp parallel
    {
        int *idx = (int *)malloc(nb_neighbors * sizeof(int));
        double *d2 = (double *)malloc(nb_neighbors * sizeof(double));
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < n; ++i) {
            int c = kd_knn(t, pts + (size_t)i * 3, nb_neighbors, idx, d2);
            double mean = -1.0;
            if (c > 0) {
                double s = 0.0;
                for (int j = 0; j < c; ++j) s += sqrt(d2[j]);
                mean = s / (double)c;
            }
            avg[i] = mean;
        }
        free(idx); free(d2);
    }
    size_t valid = 0;
    double sum = 0.0;
    for (int i = 0; i < n; ++i) if (avg[i] > 0) { ++valid; sum += avg[i]; }
    if (valid == 0) {
        for (int i = 0; i < n; ++i) mask[i] = 0;
        if (avg_out) memcpy(avg_out, avg, n * sizeof(double));
        free(avg); kd_free(t);
        return 0;
    }
    double cloud_mean = sum / (double)valid;
    double sq = 0.0;
    for (int i = 0; i < n; ++i) if (avg[i] > 0) { double d = avg[i] - cloud_mean; sq += d * d; }
    double std_dev = sqrt(sq / (double)(valid - 1));
    double thr = cloud_mean + std_ratio * std_dev;
    for (int i = 0; i < n; ++i) mask[i] = (avg[i] > 0 && avg[i] < thr) ? 1 : 0;
    if (avg_out) memcpy(avg_out, avg, n * sizeof(double));
    if (stats) { stats[0] = cloud_mean; stats[1] = std_dev; stats[2] = thr; }
    free(avg); kd_free(t);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Deterministic acos / cos                                                  */
/* ------------------------------------------------------------------------ */
/* Open3D calls libm's acos / cos inside FastEigen3x3 and acos inside FPFH's
 * ComputePairFeatures, where the results feed DISCONTINUOUS decisions (which
 * eigenvector branch, which point of a pair is the "source").  libm results
 * are not specified bit for bit (glibc and the CUDA math library differ in the
 * last ulp), and range-grid scans contain structural near-ties (neighbouring
 * points with the same neighbour set have normals equal to ~1e-16), so the two
 * sides of a parity test would take different branches on a few points in
 * 5000.  The oracle and the CUDA path therefore share these implementations,
 * which use only + - * / sqrt and bit masks (Sun fdlibm's e_acos.c / k_cos.c /
 * k_sin.c algorithms restated, < 1 ulp), compiled without FMA contraction on
 * both sides: bit-identical results.                                        */
static inline double dbl_clear_low(double x)
{
    uint64_t u; memcpy(&u, &x, 8); u &= 0xFFFFFFFF00000000ull; memcpy(&x, &u, 8); return x;
}
static inline uint32_t dbl_high(double x) { uint64_t u; memcpy(&u, &x, 8); return (uint32_t)(u >> 32); }
static inline double dbl_from_high(uint32_t hi) { uint64_t u = (uint64_t)hi << 32; double x; memcpy(&x, &u, 8); return x; }

static double det_acos(double x)
{
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17,
                 pi = 3.14159265358979311600e+00;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    const double ax = fabs(x);
    if (ax >= 1.0) {
        if (ax == 1.0) return x > 0 ? 0.0 : pi + 2.0 * pio2_lo;
        return (x - x) / (x - x);
    }
    if (ax < 0.5) {
        if (ax < 6.938893903907228e-18) return pio2_hi + pio2_lo;   /* 2^-57 */
        const double z = x * x;
        const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const double r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    } else if (x < 0) {
        const double z = (1.0 + x) * 0.5;
        const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const double s = sqrt(z);
        const double r = p / q;
        const double w = r * s - pio2_lo;
        return pi - 2.0 * (s + w);
    } else {
        const double z = (1.0 - x) * 0.5;
        const double s = sqrt(z);
        const double df = dbl_clear_low(s);
        const double c = (z - df * df) / (s + df);
        const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const double r = p / q;
        const double w = r * s + c;
        return 2.0 * (df + w);
    }
}

/* cos on |x| <= pi/4 (+ tail y) */
static double det_kcos(double x, double y)
{
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double ax = fabs(x);
    if (ax < 7.450580596923828e-09) return 1.0;   /* 2^-27 */
    const double z = x * x;
    const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    if (ax < 0.3) return 1.0 - (0.5 * z - (z * r - x * y));
    const double qx = ax > 0.78125 ? 0.28125 : dbl_from_high(dbl_high(ax) - 0x00200000u);
    const double hz = 0.5 * z - qx;
    const double a = 1.0 - qx;
    return a - (hz - (z * r - x * y));
}

/* sin on |x| <= pi/4 (+ tail y) */
static double det_ksin(double x, double y)
{
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = x * x;
    const double v = z * x;
    const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

/* cos(x) for 0 <= x <= pi (the only range FastEigen3x3 needs) */
static double det_cos(double x)
{
    const double pio4 = 7.85398163397448278999e-01, three_pio4 = 2.35619449019234483700e+00;
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
    const double pi_hi = 3.14159265358979311600e+00, pi_lo = 1.22464679914735317723e-16;
    if (x <= pio4) return det_kcos(x, 0.0);
    if (x < three_pio4) {
        const double t = x - pio2_hi;                 /* exact (Sterbenz) */
        const double y = t - pio2_lo;
        const double yt = (t - y) - pio2_lo;
        return -det_ksin(y, yt);
    }
    const double t = x - pi_hi;                       /* exact (Sterbenz) */
    const double y = t - pi_lo;
    const double yt = (t - y) - pi_lo;
    return -det_kcos(y, yt);
}

/* atan (Sun fdlibm s_atan.c restated: argument reduction to |x| < 7/16 around 0.5, 1, 1.5, inf;
 * odd polynomial split into two Horner chains), < 1 ulp */
static double det_atan(double x)
{
    const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01,
                              1.57079632679489655800e+00};
    const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17,
                              6.12323399573676603587e-17};
    const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01, aT2 = 1.42857142725034663711e-01,
                 aT3 = -1.11111104054623557880e-01, aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
                 aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02, aT8 = 4.97687799461593236017e-02,
                 aT9 = -3.65315727442169155270e-02, aT10 = 1.62858201153657823623e-02;
    const uint32_t hx = dbl_high(x), ix = hx & 0x7fffffffu;
    int id;
    if (ix >= 0x44100000u) {                        /* |x| >= 2^66 (or NaN) */
        if (x != x) return x + x;
        return (hx >> 31) ? -atanhi[3] - atanlo[3] : atanhi[3] + atanlo[3];
    }
    if (ix < 0x3fdc0000u) {                         /* |x| < 0.4375 */
        if (ix < 0x3e200000u) return x;             /* |x| < 2^-29 */
        id = -1;
    } else {
        x = fabs(x);
        if (ix < 0x3ff30000u) {                     /* |x| < 1.1875 */
            if (ix < 0x3fe60000u) { id = 0; x = (2.0 * x - 1.0) / (2.0 + x); }
            else { id = 1; x = (x - 1.0) / (x + 1.0); }
        } else {
            if (ix < 0x40038000u) { id = 2; x = (x - 1.5) / (1.0 + 1.5 * x); }
            else { id = 3; x = -1.0 / x; }
        }
    }
    const double z = x * x;
    const double w = z * z;
    const double s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const double s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) return x - x * (s1 + s2);
    const double r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return (hx >> 31) ? -r : r;
}

/* atan2 (fdlibm e_atan2.c restated; finite arguments, signed zeros as in C).  Open3D's
 * ComputePairFeatures bins floor(11 (atan2(..) + pi) / (2 pi)): a discontinuous consumer. */
static double det_atan2(double y, double x)
{
    const double pi_o_2 = 1.5707963267948965580e+00, pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16;
    if (x != x || y != y) return x + y;
    if (x == 1.0) return det_atan(y);
    const uint32_t hx = dbl_high(x), hy = dbl_high(y);
    const int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);       /* 2 * sign(x) + sign(y) */
    if (y == 0.0) {
        switch (m) {
        case 0: case 1: return y;
        case 2: return pi;
        default: return -pi;
        }
    }
    if (x == 0.0) return (hy >> 31) ? -pi_o_2 : pi_o_2;
    const int ix = (int)(hx & 0x7fffffffu), iy = (int)(hy & 0x7fffffffu);
    if (ix >= 0x7ff00000 || iy >= 0x7ff00000) {                          /* infinities: libm semantics not needed here */
        if (iy >= 0x7ff00000 && ix < 0x7ff00000) return (hy >> 31) ? -pi_o_2 : pi_o_2;
        if (ix >= 0x7ff00000 && iy < 0x7ff00000) {
            switch (m) { case 0: return 0.0; case 1: return -0.0; case 2: return pi; default: return -pi; }
        }
        switch (m) { case 0: return 0.5 * pi_o_2; case 1: return -0.5 * pi_o_2; case 2: return 1.5 * pi_o_2; default: return -1.5 * pi_o_2; }
    }
    const int k = (iy - ix) >> 20;
    double z;
    if (k > 60) z = pi_o_2 + 0.5 * pi_lo;
    else if ((hx >> 31) && k < -60) z = 0.0;
    else z = det_atan(fabs(y / x));
    switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

ORACLE_API double oracle_det_acos(double x) { return det_acos(x); }
ORACLE_API double oracle_det_cos(double x) { return det_cos(x); }
ORACLE_API double oracle_det_atan2(double y, double x) { return det_atan2(y, x); }

/* ------------------------------------------------------------------------ */
/* Covariance + FastEigen3x3 normal (EstimateNormals.cpp)     */
/* ------------------------------------------------------------------------ */

/* utility::ComputeCovariance: 9 cumulants in neighbour order, one-pass form. */
static void covariance_from_indices(const double *pts, const int *idx, int cnt, double C[9])
{
    double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < cnt; ++j) {
        const double *p = pts + (size_t)idx[j] * 3;
        c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
        c[3] += p[0] * p[0]; c[4] += p[0] * p[1]; c[5] += p[0] * p[2];
        c[6] += p[1] * p[1]; c[7] += p[1] * p[2]; c[8] += p[2] * p[2];
    }
    const double n = (double)cnt;
    for (int j = 0; j < 9; ++j) c[j] /= n;
    C[0] = c[3] - c[0] * c[0];
    C[4] = c[6] - c[1] * c[1];
    C[8] = c[8] - c[2] * c[2];
    C[1] = C[3] = c[4] - c[0] * c[1];
    C[2] = C[6] = c[5] - c[0] * c[2];
    C[5] = C[7] = c[7] - c[1] * c[2];
}

static inline void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot3(const double a[3], const double b[3])
{
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline double norm3(const double *a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

/* eigenvector of the symmetric A (row-major 3x3) for eigenvalue ev: the
 * largest of the three row cross products of A - ev*I, normalised.         */
static void eigvec_from_rows(const double A[9], double ev, double out[3])
{
    double r0[3] = {A[0] - ev, A[1], A[2]};
    double r1[3] = {A[1], A[4] - ev, A[5]};
    double r2[3] = {A[2], A[5], A[8] - ev};
    double c01[3], c02[3], c12[3];
    cross3(r0, r1, c01); cross3(r0, r2, c02); cross3(r1, r2, c12);
    double d0 = dot3(c01, c01), d1 = dot3(c02, c02), d2 = dot3(c12, c12);
    double dmax = d0; int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    const double *c = imax == 0 ? c01 : (imax == 1 ? c02 : c12);
    double s = sqrt(imax == 0 ? d0 : (imax == 1 ? d1 : d2));
    out[0] = c[0] / s; out[1] = c[1] / s; out[2] = c[2] / s;
}

/* second eigenvector: solve the 2x2 problem in the plane orthogonal to e0. */
static void eigvec_in_plane(const double A[9], const double e0[3], double ev1, double out[3])
{
    double U[3], V[3];
    if (fabs(e0[0]) > fabs(e0[1])) {
        double inv = 1 / sqrt(e0[0] * e0[0] + e0[2] * e0[2]);
        U[0] = -e0[2] * inv; U[1] = 0; U[2] = e0[0] * inv;
    } else {
        double inv = 1 / sqrt(e0[1] * e0[1] + e0[2] * e0[2]);
        U[0] = 0; U[1] = e0[2] * inv; U[2] = -e0[1] * inv;
    }
    cross3(e0, U, V);
    double AU[3] = {A[0] * U[0] + A[1] * U[1] + A[2] * U[2],
                    A[1] * U[0] + A[4] * U[1] + A[5] * U[2],
                    A[2] * U[0] + A[5] * U[1] + A[8] * U[2]};
    double AV[3] = {A[0] * V[0] + A[1] * V[1] + A[2] * V[2],
                    A[1] * V[0] + A[4] * V[1] + A[5] * V[2],
                    A[2] * V[0] + A[5] * V[1] + A[8] * V[2]};
    double m00 = U[0] * AU[0] + U[1] * AU[1] + U[2] * AU[2] - ev1;
    double m01 = U[0] * AV[0] + U[1] * AV[1] + U[2] * AV[2];
    double m11 = V[0] * AV[0] + V[1] * AV[1] + V[2] * AV[2] - ev1;
    double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    if (a00 >= a11) {
        double mx = a00 > a01 ? a00 : a01;
        if (mx > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1 / sqrt(1 + m01 * m01); m01 *= m00; }
            else            { m00 /= m01; m01 = 1 / sqrt(1 + m00 * m00); m00 *= m01; }
            for (int i = 0; i < 3; ++i) out[i] = m01 * U[i] - m00 * V[i];
        } else {
            for (int i = 0; i < 3; ++i) out[i] = U[i];
        }
    } else {
        double mx = a11 > a01 ? a11 : a01;
        if (mx > 0) {
            if (a11 >= a01) { m01 /= m11; m11 = 1 / sqrt(1 + m01 * m01); m01 *= m11; }
            else            { m11 /= m01; m01 = 1 / sqrt(1 + m11 * m11); m11 *= m01; }
            for (int i = 0; i < 3; ++i) out[i] = m11 * U[i] - m01 * V[i];
        } else {
            for (int i = 0; i < 3; ++i) out[i] = U[i];
        }
    }
}

/* FastEigen3x3: eigenvector of the smallest eigenvalue, closed form. */
static void smallest_eigvec(const double C[9], double n[3])
{
    double A[9];
    double mc = C[0];
    for (int i = 1; i < 9; ++i) if (C[i] > mc) mc = C[i];
    if (mc == 0) { n[0] = n[1] = n[2] = 0; return; }
    for (int i = 0; i < 9; ++i) A[i] = C[i] / mc;
    double norm = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (norm > 0) {
        double q = (A[0] + A[4] + A[8]) / 3;
        double b00 = A[0] - q, b11 = A[4] - q, b22 = A[8] - q;
        double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
        double c00 = b11 * b22 - A[5] * A[5];
        double c01 = A[1] * b22 - A[5] * A[2];
        double c02 = A[1] * A[5] - b11 * A[2];
        double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
        double half_det = det * 0.5;
        half_det = fmin(fmax(half_det, -1.0), 1.0);
        double angle = det_acos(half_det) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        double beta2 = det_cos(angle) * 2;
        double beta0 = det_cos(angle + two_thirds_pi) * 2;
        double beta1 = -(beta0 + beta2);
        double ev0 = q + p * beta0, ev1 = q + p * beta1, ev2 = q + p * beta2;
        double e0[3], e1[3], e2[3];
        if (half_det >= 0) {
            eigvec_from_rows(A, ev2, e2);
            if (ev2 < ev0 && ev2 < ev1) { memcpy(n, e2, sizeof(e2)); return; }
            eigvec_in_plane(A, e2, ev1, e1);
            if (ev1 < ev0 && ev1 < ev2) { memcpy(n, e1, sizeof(e1)); return; }
            cross3(e1, e2, n);
            return;
        } else {
            eigvec_from_rows(A, ev0, e0);
            if (ev0 < ev1 && ev0 < ev2) { memcpy(n, e0, sizeof(e0)); return; }
            eigvec_in_plane(A, e0, ev1, e1);
            if (ev1 < ev0 && ev1 < ev2) { memcpy(n, e1, sizeof(e1)); return; }
            cross3(e0, e1, n);
            return;
        }
    } else {
        /* diagonal matrix: compared after scaling back by max_coeff */
        const double a0 = A[0] * mc, a1 = A[4] * mc, a2 = A[8] * mc;
        if (a0 < a1 && a0 < a2) { n[0] = 1; n[1] = 0; n[2] = 0; }
        else if (a1 < a0 && a1 < a2) { n[0] = 0; n[1] = 1; n[2] = 0; }
        else { n[0] = 0; n[1] = 0; n[2] = 1; }
    }
}

/* GetRotationFromE1ToX + C = Rx diag(eps,1,1) Rx^T  (GeneralizedICP.cpp). */
static void gicp_cov_from_normal(const double nrm[3], double eps, double C[9])
{
    double R[9];
    const double c = nrm[0];                     /* e1 . x            */
    const double v[3] = {0.0, -nrm[2], nrm[1]};  /* e1 x x            */
    if (c < -0.99) {
        for (int i = 0; i < 9; ++i) R[i] = 0;
        R[0] = R[4] = R[8] = 1;
    } else {
        /* sv = skew(v); R = I + sv + sv*sv * (1/(1+c)) */
        double sv[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
        double sv2[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                sv2[i * 3 + j] = sv[i * 3 + 0] * sv[0 * 3 + j] + sv[i * 3 + 1] * sv[1 * 3 + j] +
                                 sv[i * 3 + 2] * sv[2 * 3 + j];
        const double f = 1 / (1 + c);
        for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + sv[i] + sv2[i] * f;
    }
    /* (R * D) * R^T with D = diag(eps,1,1) */
    double RD[9];
    for (int i = 0; i < 3; ++i) { RD[i * 3] = R[i * 3] * eps; RD[i * 3 + 1] = R[i * 3 + 1]; RD[i * 3 + 2] = R[i * 3 + 2]; }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = RD[i * 3] * R[j * 3] + RD[i * 3 + 1] * R[j * 3 + 1] + RD[i * 3 + 2] * R[j * 3 + 2];
}

/* EstimateNormals on a cloud without prior normals.
 *   knn: max_nn;  radius <= 0 => KDTreeSearchParamKNN(knn)
 *                 radius  > 0 => KDTreeSearchParamHybrid(radius, knn)
 * Outputs (any may be NULL): covs n x 9 (per-point covariance), normals n x 3,
 * gicp_covs n x 9 (only meaningful with gicp_eps > 0), nn_idx n x knn,
 * nn_cnt n.                                                                */
ORACLE_API int oracle_normals(const double *pts, int n, int knn, double radius, double gicp_eps,
                              double *covs, double *normals, double *gicp_covs,
                              int *nn_idx, int *nn_cnt)
{
    KDTree *t = kd_build(pts, n, 3);
    const double r2 = radius * radius;
#pragma omp parallel
    {
        int *idx = (int *)malloc(knn * sizeof(int));
        double *d2 = (double *)malloc(knn * sizeof(double));
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < n; ++i) {
            int c = kd_knn(t, pts + (size_t)i * 3, knn, idx, d2);
            if (radius > 0) { int cc = 0; while (cc < c && d2[cc] < r2) ++cc; c = cc; }
            double C[9];
            if (c >= 3) covariance_from_indices(pts, idx, c, C);
            else { for (int j = 0; j < 9; ++j) C[j] = 0; C[0] = C[4] = C[8] = 1; }
            double nr[3];
            smallest_eigvec(C, nr);
            if (sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]) == 0.0) { nr[0] = 0; nr[1] = 0; nr[2] = 1; }
            if (covs) memcpy(covs + (size_t)i * 9, C, sizeof(C));
            if (normals) memcpy(normals + (size_t)i * 3, nr, sizeof(nr));
            if (gicp_covs) gicp_cov_from_normal(nr, gicp_eps, gicp_covs + (size_t)i * 9);
            if (nn_idx) { for (int j = 0; j < knn; ++j) nn_idx[(size_t)i * knn + j] = j < c ? idx[j] : -1; }
            if (nn_cnt) nn_cnt[i] = c;
        }
        free(idx); free(d2);
    }
    kd_free(t);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* small dense helpers                                                       */
/* ------------------------------------------------------------------------ */

/* Symmetric 3x3 eigendecomposition by cyclic Jacobi (only + - * / sqrt, so a
 * CUDA twin can reproduce it bit for bit).  A row-major; V columns = vectors. */
static void jacobi3(const double Ain[9], double w[3], double V[9])
{
    double a00 = Ain[0], a01 = Ain[1], a02 = Ain[2], a11 = Ain[4], a12 = Ain[5], a22 = Ain[8];
    double v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = a01 * a01 + a02 * a02 + a12 * a12;
        if (off == 0.0) break;
        /* rotate (0,1) */
        if (a01 != 0.0) {
            double theta = (a11 - a00) / (2.0 * a01);
            double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
            double n00 = a00 - tt * a01, n11 = a11 + tt * a01;
            double n02 = c * a02 - s * a12, n12 = s * a02 + c * a12;
            a00 = n00; a11 = n11; a01 = 0.0; a02 = n02; a12 = n12;
            for (int r = 0; r < 3; ++r) {
                double x = v[r * 3 + 0], y = v[r * 3 + 1];
                v[r * 3 + 0] = c * x - s * y; v[r * 3 + 1] = s * x + c * y;
            }
        }
        /* rotate (0,2) */
        if (a02 != 0.0) {
            double theta = (a22 - a00) / (2.0 * a02);
            double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
            double n00 = a00 - tt * a02, n22 = a22 + tt * a02;
            double n01 = c * a01 - s * a12, n12 = s * a01 + c * a12;
            a00 = n00; a22 = n22; a02 = 0.0; a01 = n01; a12 = n12;
            for (int r = 0; r < 3; ++r) {
                double x = v[r * 3 + 0], z = v[r * 3 + 2];
                v[r * 3 + 0] = c * x - s * z; v[r * 3 + 2] = s * x + c * z;
            }
        }
        /* rotate (1,2) */
        if (a12 != 0.0) {
            double theta = (a22 - a11) / (2.0 * a12);
            double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
            double n11 = a11 - tt * a12, n22 = a22 + tt * a12;
            double n01 = c * a01 - s * a02, n02 = s * a01 + c * a02;
            a11 = n11; a22 = n22; a12 = 0.0; a01 = n01; a02 = n02;
            for (int r = 0; r < 3; ++r) {
                double y = v[r * 3 + 1], z = v[r * 3 + 2];
                v[r * 3 + 1] = c * y - s * z; v[r * 3 + 2] = s * y + c * z;
            }
        }
    }
    w[0] = a00; w[1] = a11; w[2] = a22;
    memcpy(V, v, sizeof(v));
}

/* W = (M^-1)^(1/2) for symmetric positive definite M (Eigen .inverse().sqrt()
 * restated through the symmetric eigendecomposition).        */
static void inv_sqrt_sym3(const double M[9], double W[9])
{
    double w[3], V[9];
    jacobi3(M, w, V);
    double s0 = 1.0 / sqrt(w[0]), s1 = 1.0 / sqrt(w[1]), s2 = 1.0 / sqrt(w[2]);
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            double x = V[i * 3 + 0] * s0 * V[j * 3 + 0] + V[i * 3 + 1] * s1 * V[j * 3 + 1] +
                       V[i * 3 + 2] * s2 * V[j * 3 + 2];
            W[i * 3 + j] = x; W[j * 3 + i] = x;
        }
}
ORACLE_API void oracle_inv_sqrt_sym3(const double *M, double *W) { inv_sqrt_sym3(M, W); }

/* x = A.ldlt().solve(b), 6x6, symmetric diagonal pivoting as in Eigen's LDLT. */
static void ldlt_solve6(const double Ain[36], const double bin[6], double x[6])
{
    double A[36]; int perm[6];
    memcpy(A, Ain, sizeof(A));
    for (int i = 0; i < 6; ++i) perm[i] = i;
    for (int k = 0; k < 6; ++k) {
        int piv = k; double best = fabs(A[k * 6 + k]);
        for (int i = k + 1; i < 6; ++i) if (fabs(A[i * 6 + i]) > best) { best = fabs(A[i * 6 + i]); piv = i; }
        if (piv != k) {
            for (int j = 0; j < 6; ++j) { double t = A[k * 6 + j]; A[k * 6 + j] = A[piv * 6 + j]; A[piv * 6 + j] = t; }
            for (int j = 0; j < 6; ++j) { double t = A[j * 6 + k]; A[j * 6 + k] = A[j * 6 + piv]; A[j * 6 + piv] = t; }
            int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        double d = A[k * 6 + k];
        if (d == 0.0) continue;
        double col[6];
        for (int i = k + 1; i < 6; ++i) col[i] = A[i * 6 + k];
        for (int i = k + 1; i < 6; ++i) {
            double l = col[i] / d;
            for (int j = k + 1; j <= i; ++j) A[i * 6 + j] -= l * col[j];
            A[i * 6 + k] = l;
        }
        for (int i = k + 1; i < 6; ++i)
            for (int j = i + 1; j < 6; ++j) A[i * 6 + j] = A[j * 6 + i];
    }
    double y[6];
    for (int i = 0; i < 6; ++i) y[i] = bin[perm[i]];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) y[i] -= A[i * 6 + j] * y[j];
    for (int i = 0; i < 6; ++i) y[i] = A[i * 6 + i] != 0.0 ? y[i] / A[i * 6 + i] : 0.0;
    for (int i = 5; i >= 0; --i) for (int j = i + 1; j < 6; ++j) y[i] -= A[j * 6 + i] * y[j];
    for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
}
ORACLE_API void oracle_ldlt_solve6(const double *A, const double *b, double *x) { ldlt_solve6(A, b, x); }

/* TransformVector6dToMatrix4d: [Rz(g) Ry(b) Rx(a) | t], row-major 4x4. */
static void vec6_to_mat4(const double x[6], double T[16])
{
    double ca = cos(x[0]), sa = sin(x[0]);
    double cb = cos(x[1]), sb = sin(x[1]);
    double cg = cos(x[2]), sg = sin(x[2]);
    T[0] = cg * cb; T[1] = cg * sb * sa - sg * ca; T[2] = cg * sb * ca + sg * sa; T[3] = x[3];
    T[4] = sg * cb; T[5] = sg * sb * sa + cg * ca; T[6] = sg * sb * ca - cg * sa; T[7] = x[4];
    T[8] = -sb;     T[9] = cb * sa;                T[10] = cb * ca;               T[11] = x[5];
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

static void mat4_mul(const double A[16], const double B[16], double C[16])
{
    double o[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            o[i * 4 + j] = A[i * 4] * B[j] + A[i * 4 + 1] * B[4 + j] + A[i * 4 + 2] * B[8 + j] + A[i * 4 + 3] * B[12 + j];
    memcpy(C, o, sizeof(o));
}

/* PointCloud::Transform on points (+ optional covariances n x 9). */
static void transform_cloud(double *pts, double *covs, int n, const double U[16])
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        double *p = pts + (size_t)i * 3;
        double x = p[0], y = p[1], z = p[2];
        double nx = U[0] * x + U[1] * y + U[2] * z + U[3];
        double ny = U[4] * x + U[5] * y + U[6] * z + U[7];
        double nz = U[8] * x + U[9] * y + U[10] * z + U[11];
        p[0] = nx; p[1] = ny; p[2] = nz;
        if (covs) {
            double *C = covs + (size_t)i * 9, RC[9], O[9];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    RC[r * 3 + c] = U[r * 4] * C[c] + U[r * 4 + 1] * C[3 + c] + U[r * 4 + 2] * C[6 + c];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    O[r * 3 + c] = RC[r * 3] * U[c * 4] + RC[r * 3 + 1] * U[c * 4 + 1] + RC[r * 3 + 2] * U[c * 4 + 2];
            memcpy(C, O, sizeof(O));
        }
    }
}

/* GetRegistrationResultAndCorrespondences (Registration.cpp):
 * SearchHybrid(p, max_dist, 1) for every source point, index-ordered sums.  */
static void evaluate_corr(const double *src, int ns, const KDTree *tt, double max_dist,
                          int *corr, double *corr_d2, double *fitness, double *rmse, int *n_corr)
{
    const double r2 = max_dist * max_dist;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < ns; ++i) {
        int j; double d;
        int c = kd_knn(tt, src + (size_t)i * 3, 1, &j, &d);
        if (c > 0 && d < r2) { corr[i] = j; corr_d2[i] = d; } else { corr[i] = -1; corr_d2[i] = 0; }
    }
    double e2 = 0; int cnt = 0;
    for (int i = 0; i < ns; ++i) if (corr[i] >= 0) { e2 += corr_d2[i]; ++cnt; }
    *n_corr = cnt;
    if (cnt == 0) { *fitness = 0; *rmse = 0; }
    else { *fitness = (double)cnt / (double)ns; *rmse = sqrt(e2 / (double)cnt); }
}

/* registration::EvaluateRegistration for a given 4x4 (row-major, column
 * convention x' = R x + t).                                                 */
ORACLE_API int oracle_evaluate(const double *source, int ns, const double *target, int nt,
                               double max_dist, const double *T, double *fitness, double *rmse,
                               int *n_corr, int *corr_out)
{
    double *s = (double *)malloc((size_t)ns * 3 * sizeof(double));
    memcpy(s, source, (size_t)ns * 3 * sizeof(double));
    transform_cloud(s, NULL, ns, T);
    KDTree *tt = kd_build(target, nt, 3);
    int *corr = (int *)malloc(ns * sizeof(int));
    double *cd = (double *)malloc(ns * sizeof(double));
    evaluate_corr(s, ns, tt, max_dist, corr, cd, fitness, rmse, n_corr);
    if (corr_out) memcpy(corr_out, corr, ns * sizeof(int));
    free(corr); free(cd); free(s); kd_free(tt);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* K2: registration_generalized_icp                      */
/* ------------------------------------------------------------------------ */

/* One ComputeTransformation: index-ordered J^T J / J^T r, LDLT, 6-vector.   */
static void gicp_step(const double *src, const double *scov, const double *tgt, const double *tcov,
                      const int *corr, int ns, double *work /* ns*27 */, double x[6], double U[16])
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < ns; ++i) {
        double *o = work + (size_t)i * 27;
        int j = corr[i];
        if (j < 0) continue;
        const double *vs = src + (size_t)i * 3, *vt = tgt + (size_t)j * 3;
        const double *Cs = scov + (size_t)i * 9, *Ct = tcov + (size_t)j * 9;
        double d[3] = {vs[0] - vt[0], vs[1] - vt[1], vs[2] - vt[2]};
        double M[9], W[9];
        for (int k = 0; k < 9; ++k) M[k] = Ct[k] + Cs[k];
        inv_sqrt_sym3(M, W);
        /* J = W * [-skew(vs) | I] */
        double S[9] = {0, vs[2], -vs[1], -vs[2], 0, vs[0], vs[1], -vs[0], 0}; /* = -skew(vs) */
        double J[18];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c)
                J[r * 6 + c] = W[r * 3] * S[c] + W[r * 3 + 1] * S[3 + c] + W[r * 3 + 2] * S[6 + c];
            J[r * 6 + 3] = W[r * 3]; J[r * 6 + 4] = W[r * 3 + 1]; J[r * 6 + 5] = W[r * 3 + 2];
        }
        double rr[3];
        for (int r = 0; r < 3; ++r) rr[r] = W[r * 3] * d[0] + W[r * 3 + 1] * d[1] + W[r * 3 + 2] * d[2];
        int k = 0;
        for (int a = 0; a < 6; ++a)
            for (int b = a; b < 6; ++b)
                o[k++] = J[a] * J[b] + J[6 + a] * J[6 + b] + J[12 + a] * J[12 + b];
        for (int a = 0; a < 6; ++a) o[k++] = J[a] * rr[0] + J[6 + a] * rr[1] + J[12 + a] * rr[2];
    }
    double acc[27];
    for (int k = 0; k < 27; ++k) acc[k] = 0;
    int cnt = 0;
    for (int i = 0; i < ns; ++i) {
        if (corr[i] < 0) continue;
        const double *o = work + (size_t)i * 27;
        for (int k = 0; k < 27; ++k) acc[k] += o[k];
        ++cnt;
    }
    if (cnt == 0) {
        for (int k = 0; k < 6; ++k) x[k] = 0;
        for (int k = 0; k < 16; ++k) U[k] = (k % 5 == 0) ? 1.0 : 0.0;
        return;
    }
    double JTJ[36], nb[6];
    int k = 0;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) { JTJ[a * 6 + b] = acc[k]; JTJ[b * 6 + a] = acc[k]; ++k; }
    for (int a = 0; a < 6; ++a) nb[a] = -acc[21 + a];
    ldlt_solve6(JTJ, nb, x);
    vec6_to_mat4(x, U);
}

/* Full registration_generalized_icp(source, target, max_dist, I, GICP(eps=1e-3),
 * ICPConvergenceCriteria(max_iteration)).  T_out: row-major 4x4, Open3D column
 * convention.  trace (optional): per Evaluate (fitness, rmse), (max_iter+1) x 2;
 * xtrace (optional): per iteration 6-vector, max_iter x 6.
 * src_cov_in / tgt_cov_in (optional): precomputed GICP covariances (n x 9).   */
ORACLE_API int oracle_gicp(const double *source, int ns, const double *target, int nt,
                           double max_dist, int max_iter, double eps,
                           const double *src_cov_in, const double *tgt_cov_in,
                           double *T_out, double *fitness_out, double *rmse_out, int *iters_out,
                           int *n_corr_out, double *trace, double *xtrace)
{
    double *src = (double *)malloc((size_t)ns * 3 * sizeof(double));
    double *scov = (double *)malloc((size_t)ns * 9 * sizeof(double));
    double *tcov = (double *)malloc((size_t)nt * 9 * sizeof(double));
    memcpy(src, source, (size_t)ns * 3 * sizeof(double));
    if (src_cov_in) memcpy(scov, src_cov_in, (size_t)ns * 9 * sizeof(double));
    else oracle_normals(source, ns, 20, 0.0, eps, NULL, NULL, scov, NULL, NULL);
    if (tgt_cov_in) memcpy(tcov, tgt_cov_in, (size_t)nt * 9 * sizeof(double));
    else oracle_normals(target, nt, 20, 0.0, eps, NULL, NULL, tcov, NULL, NULL);
    KDTree *tt = kd_build(target, nt, 3);
    int *corr = (int *)malloc(ns * sizeof(int));
    double *cd = (double *)malloc(ns * sizeof(double));
    double *work = (double *)malloc((size_t)ns * 27 * sizeof(double));
    double T[16];
    for (int k = 0; k < 16; ++k) T[k] = (k % 5 == 0) ? 1.0 : 0.0;
    double fit, rmse; int nc;
    evaluate_corr(src, ns, tt, max_dist, corr, cd, &fit, &rmse, &nc);
    if (trace) { trace[0] = fit; trace[1] = rmse; }
    int it = 0;
    for (int i = 0; i < max_iter; ++i) {
        double x[6], U[16];
        gicp_step(src, scov, target, tcov, corr, ns, work, x, U);
        if (xtrace) memcpy(xtrace + (size_t)i * 6, x, sizeof(x));
        mat4_mul(U, T, T);
        transform_cloud(src, scov, ns, U);
        double ofit = fit, ormse = rmse;
        evaluate_corr(src, ns, tt, max_dist, corr, cd, &fit, &rmse, &nc);
        it = i + 1;
        if (trace) { trace[(size_t)it * 2] = fit; trace[(size_t)it * 2 + 1] = rmse; }
        if (fabs(ofit - fit) < 1e-6 && fabs(ormse - rmse) < 1e-6) break;
    }
    memcpy(T_out, T, sizeof(T));
    *fitness_out = fit; *rmse_out = rmse; *iters_out = it;
    if (n_corr_out) *n_corr_out = nc;
    free(src); free(scov); free(tcov); free(corr); free(cd); free(work); kd_free(tt);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* K6: registration_icp, point-to-point / point-to-plane                     */
/*     (Registration.cpp RegistrationICP, TransformationEstimation.cpp;      */
/*      Aligner/Aligner.py:319-364)                                          */
/* ------------------------------------------------------------------------ */

/* Eigen::umeyama(src, dst, with_scaling = false) over the correspondences, restated: means,
 * demeaned cross-covariance sigma = 1/n sum (t - mt)(s - ms)^T, optimal proper rotation from the
 * two leading singular pairs, R = [u1 u2 u1xu2] [v1 v2 v1xv2]^T (equals U diag(1,1,det(U)det(V)) V^T
 * of the full SVD), translation mt - R ms.                                            */
static void umeyama_step(const double *src, const double *tgt, const int *corr, int ns, double U[16])
{
    double ms[3] = {0, 0, 0}, mt[3] = {0, 0, 0};
    int n = 0;
    for (int i = 0; i < ns; ++i) {
        const int j = corr[i];
        if (j < 0) continue;
        for (int k = 0; k < 3; ++k) { ms[k] += src[(size_t)i * 3 + k]; mt[k] += tgt[(size_t)j * 3 + k]; }
        ++n;
    }
    for (int k = 0; k < 16; ++k) U[k] = (k % 5 == 0) ? 1.0 : 0.0;
    if (n == 0) return;
    for (int k = 0; k < 3; ++k) { ms[k] /= n; mt[k] /= n; }
    double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < ns; ++i) {
        const int j = corr[i];
        if (j < 0) continue;
        double a[3], b[3];
        for (int k = 0; k < 3; ++k) { a[k] = tgt[(size_t)j * 3 + k] - mt[k]; b[k] = src[(size_t)i * 3 + k] - ms[k]; }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) S[r * 3 + c] += a[r] * b[c];
    }
    for (int k = 0; k < 9; ++k) S[k] /= n;
    /* eigen-decomposition of S^T S: right singular vectors */
    double B[9], w[3], V[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) B[r * 3 + c] = S[0 * 3 + r] * S[0 * 3 + c] + S[1 * 3 + r] * S[1 * 3 + c] + S[2 * 3 + r] * S[2 * 3 + c];
    jacobi3(B, w, V);
    int o[3] = {0, 1, 2};   /* order by decreasing eigenvalue */
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (w[o[b]] > w[o[a]]) { int t = o[a]; o[a] = o[b]; o[b] = t; }
    double v1[3] = {V[0 * 3 + o[0]], V[1 * 3 + o[0]], V[2 * 3 + o[0]]};
    double v2[3] = {V[0 * 3 + o[1]], V[1 * 3 + o[1]], V[2 * 3 + o[1]]};
    double v3[3]; cross3(v1, v2, v3);
    double u1[3], u2[3], u3[3];
    for (int r = 0; r < 3; ++r) { u1[r] = S[r * 3] * v1[0] + S[r * 3 + 1] * v1[1] + S[r * 3 + 2] * v1[2];
                                  u2[r] = S[r * 3] * v2[0] + S[r * 3 + 1] * v2[1] + S[r * 3 + 2] * v2[2]; }
    double n1 = norm3(u1);
    if (!(n1 > 0)) return;                 /* sigma = 0: no information, identity */
    for (int r = 0; r < 3; ++r) u1[r] /= n1;
    double d12 = dot3(u1, u2);
    for (int r = 0; r < 3; ++r) u2[r] -= d12 * u1[r];
    double n2 = norm3(u2);
    if (!(n2 > 1e-300)) {                  /* rank 1: any unit vector orthogonal to u1 */
        double e[3] = {0, 0, 0};
        e[fabs(u1[0]) <= fabs(u1[1]) ? (fabs(u1[0]) <= fabs(u1[2]) ? 0 : 2) : (fabs(u1[1]) <= fabs(u1[2]) ? 1 : 2)] = 1.0;
        cross3(u1, e, u2); n2 = norm3(u2);
    }
    for (int r = 0; r < 3; ++r) u2[r] /= n2;
    cross3(u1, u2, u3);
    double R[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[r * 3 + c] = u1[r] * v1[c] + u2[r] * v2[c] + u3[r] * v3[c];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) U[r * 4 + c] = R[r * 3 + c];
        U[r * 4 + 3] = mt[r] - (R[r * 3] * ms[0] + R[r * 3 + 1] * ms[1] + R[r * 3 + 2] * ms[2]);
    }
}

/* TransformationEstimationPointToPlane::ComputeTransformation: r = (vs - vt).nt, J = [vs x nt, nt]. */
static void point_to_plane_step(const double *src, const double *tgt, const double *tn, const int *corr, int ns, double U[16])
{
    double JTJ[36], JTr[6];
    for (int k = 0; k < 36; ++k) JTJ[k] = 0;
    for (int k = 0; k < 6; ++k) JTr[k] = 0;
    int n = 0;
    for (int i = 0; i < ns; ++i) {
        const int j = corr[i];
        if (j < 0) continue;
        const double *vs = src + (size_t)i * 3, *vt = tgt + (size_t)j * 3, *nt = tn + (size_t)j * 3;
        const double d[3] = {vs[0] - vt[0], vs[1] - vt[1], vs[2] - vt[2]};
        const double r = dot3(d, nt);
        double J[6];
        cross3(vs, nt, J);
        J[3] = nt[0]; J[4] = nt[1]; J[5] = nt[2];
        for (int a = 0; a < 6; ++a) {
            for (int b = 0; b < 6; ++b) JTJ[a * 6 + b] += J[a] * J[b];
            JTr[a] += J[a] * r;
        }
        ++n;
    }
    for (int k = 0; k < 16; ++k) U[k] = (k % 5 == 0) ? 1.0 : 0.0;
    if (n == 0) return;
    double nb[6], x[6];
    for (int a = 0; a < 6; ++a) nb[a] = -JTr[a];
    ldlt_solve6(JTJ, nb, x);
    vec6_to_mat4(x, U);
}

/* registration_icp(source, target, max_dist, init, estimation, ICPConvergenceCriteria(max_iteration)).
 * mode 0 = TransformationEstimationPointToPoint, 1 = PointToPlane (tgt_normals required).
 * init_T: row-major 4x4 in Open3D's column convention, or NULL for identity.            */
ORACLE_API int oracle_icp(const double *source, int ns, const double *target, int nt, double max_dist,
                          int max_iter, const double *init_T, int mode, const double *tgt_normals,
                          double *T_out, double *fitness_out, double *rmse_out, int *iters_out, int *n_corr_out)
{
    if (mode == 1 && !tgt_normals) return -1;
    double *src = (double *)malloc((size_t)ns * 3 * sizeof(double));
    memcpy(src, source, (size_t)ns * 3 * sizeof(double));
    double T[16];
    for (int k = 0; k < 16; ++k) T[k] = init_T ? init_T[k] : ((k % 5 == 0) ? 1.0 : 0.0);
    if (init_T) transform_cloud(src, NULL, ns, T);
    KDTree *tt = kd_build(target, nt, 3);
    int *corr = (int *)malloc(ns * sizeof(int));
    double *cd = (double *)malloc(ns * sizeof(double));
    double fit, rmse; int nc;
    evaluate_corr(src, ns, tt, max_dist, corr, cd, &fit, &rmse, &nc);
    int it = 0;
    for (int i = 0; i < max_iter; ++i) {
        double U[16];
        if (mode == 0) umeyama_step(src, target, corr, ns, U);
        else point_to_plane_step(src, target, tgt_normals, corr, ns, U);
        mat4_mul(U, T, T);
        transform_cloud(src, NULL, ns, U);
        const double ofit = fit, ormse = rmse;
        evaluate_corr(src, ns, tt, max_dist, corr, cd, &fit, &rmse, &nc);
        it = i + 1;
        if (fabs(ofit - fit) < 1e-6 && fabs(ormse - rmse) < 1e-6) break;
    }
    memcpy(T_out, T, sizeof(T));
    *fitness_out = fit; *rmse_out = rmse; *iters_out = it;
    if (n_corr_out) *n_corr_out = nc;
    free(src); free(corr); free(cd); kd_free(tt);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* K4: ComputeFPFHFeature (Feature.cpp;                       */
/*     fastGlobalOptimizer.py:130-143)                                       */
/* ------------------------------------------------------------------------ */

/* ComputePairFeatures(p1, n1, p2, n2) -> (f0 atan2 angle, f1, f2, f3 distance) */
static void pair_features(const double *p1, const double *n1, const double *p2, const double *n2, double f[4])
{
    double dp[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    f[0] = f[1] = f[2] = f[3] = 0.0;
    const double len = sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]);
    if (len == 0.0) return;
    double a[3] = {n1[0], n1[1], n1[2]}, b[3] = {n2[0], n2[1], n2[2]};
    const double angle1 = dot3(a, dp) / len;
    const double angle2 = dot3(b, dp) / len;
    double f2;
    if (det_acos(fabs(angle1)) > det_acos(fabs(angle2))) {
        a[0] = n2[0]; a[1] = n2[1]; a[2] = n2[2];
        b[0] = n1[0]; b[1] = n1[1]; b[2] = n1[2];
        dp[0] *= -1.0; dp[1] *= -1.0; dp[2] *= -1.0;
        f2 = -angle2;
    } else {
        f2 = angle1;
    }
    double v[3], w[3];
    cross3(dp, a, v);
    const double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (vn == 0.0) return;
    v[0] /= vn; v[1] /= vn; v[2] /= vn;
    cross3(a, v, w);
    f[3] = len;
    f[2] = f2;
    f[1] = dot3(v, b);
    f[0] = det_atan2(dot3(w, b), dot3(a, b));
}

static inline int hist_bin(double x)
{
    int h = (int)floor(x);
    if (h < 0) h = 0;
    if (h >= 11) h = 10;
    return h;
}

/* features: n x 33 (point-major; Open3D's 33 x N column-major Feature::data_ has the same
 * memory order).  radius <= 0 => KDTreeSearchParamKNN(knn), else Hybrid(radius, knn).
 * spfh_out (optional) n x 33.                                                          */
ORACLE_API int oracle_fpfh(const double *pts, const double *normals, int n, int knn, double radius,
                           double *fpfh_out, double *spfh_out)
{
    const double PI = 3.14159265358979323846;
    KDTree *t = kd_build(pts, n, 3);
    const double r2 = radius * radius;
    double *spfh = (double *)calloc((size_t)n * 33, sizeof(double));
    int *nn_idx = (int *)malloc((size_t)n * knn * sizeof(int));
    double *nn_d2 = (double *)malloc((size_t)n * knn * sizeof(double));
    int *nn_cnt = (int *)malloc((size_t)n * sizeof(int));
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        int *idx = nn_idx + (size_t)i * knn;
        double *d2 = nn_d2 + (size_t)i * knn;
        int c = kd_knn(t, pts + (size_t)i * 3, knn, idx, d2);
        if (radius > 0) { int cc = 0; while (cc < c && d2[cc] < r2) ++cc; c = cc; }
        nn_cnt[i] = c;
        if (c > 1) {
            double *h = spfh + (size_t)i * 33;
            const double incr = 100.0 / (double)(c - 1);
            for (int k = 1; k < c; ++k) {
                double f[4];
                pair_features(pts + (size_t)i * 3, normals + (size_t)i * 3, pts + (size_t)idx[k] * 3,
                              normals + (size_t)idx[k] * 3, f);
                h[hist_bin(11 * (f[0] + PI) / (2.0 * PI))] += incr;
                h[11 + hist_bin(11 * (f[1] + 1.0) * 0.5)] += incr;
                h[22 + hist_bin(11 * (f[2] + 1.0) * 0.5)] += incr;
            }
        }
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        double *o = fpfh_out + (size_t)i * 33;
        for (int j = 0; j < 33; ++j) o[j] = 0.0;
        const int c = nn_cnt[i];
        if (c > 1) {
            const int *idx = nn_idx + (size_t)i * knn;
            const double *d2 = nn_d2 + (size_t)i * knn;
            double sum[3] = {0.0, 0.0, 0.0};
            for (int k = 1; k < c; ++k) {
                const double dist = d2[k];
                if (dist == 0.0) continue;
                const double *sp = spfh + (size_t)idx[k] * 33;
                for (int j = 0; j < 33; ++j) {
                    const double val = sp[j] / dist;
                    sum[j / 11] += val;
                    o[j] += val;
                }
            }
            for (int j = 0; j < 3; ++j) if (sum[j] != 0.0) sum[j] = 100.0 / sum[j];
            const double *self = spfh + (size_t)i * 33;
            for (int j = 0; j < 33; ++j) { o[j] *= sum[j / 11]; o[j] += self[j]; }
        }
    }
    if (spfh_out) memcpy(spfh_out, spfh, (size_t)n * 33 * sizeof(double));
    free(spfh); free(nn_idx); free(nn_d2); free(nn_cnt); kd_free(t);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* K5: FastGlobalRegistrationBasedOnFeatureMatching                          */
/*     (FastGlobalRegistration.cpp;                           */
/*      fastGlobalOptimizer.py:158-174)                                      */
/* ------------------------------------------------------------------------ */

/* Philox4x32-10 (Salmon et al. 2011).  The reference leaves Open3D's std::mt19937 unseeded
 * (SURVEY.md B.5), so the tuple test has no reproducible stream to match; this project
 * fixes one: trial t of a registration keyed `key` draws
 *   (r0, r1, r2) = philox(counter = (t, 0, 0, 0), key)[0..2],  rand_k = (r_k * ncorr) >> 32. */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        if (r > 0) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
}

ORACLE_API void oracle_philox(uint64_t counter, uint64_t key, uint32_t *out4)
{
    uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    philox4x32_10(c, (uint32_t)key, (uint32_t)(key >> 32));
    memcpy(out4, c, sizeof(c));
}

static inline double dist3(const double *a, const double *b)
{
    const double d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    return norm3(d);
}

/* source (ns x 3), target (nt x 3); feat_s (nfs x 33), feat_t (nft x 33): the feature
 * sets handed to Open3D (in the reference both are the SOURCE's features, so nft == ns;
 * matched pairs whose target index is >= nt -- an out-of-bounds read in Open3D -- are
 * dropped).
 * T_out: source -> target, row-major 4x4, column convention.
 * info_out (optional) int[4]: {#cross-checked pairs, #pairs after the tuple test,
 *                              #tuple trials run, swapped}.
 * corres_out (optional): up to 3*max_tuple_count (src, tgt) pairs after the tuple test. */
ORACLE_API int oracle_fgr(const double *source, int ns, const double *target, int nt,
                          const double *feat_s, int nfs, const double *feat_t, int nft,
                          double division_factor, double tuple_scale, double max_corr_dist,
                          int iteration_number, int decrease_mu, int max_tuple_count, int tuple_test,
                          uint64_t tuple_key, double *T_out, double *fitness_out, double *rmse_out,
                          int *n_corr_out, int *info_out, int *corres_out)
{
    /* ---- NormalizePointCloud (use_absolute_scale = false) */
    double *pc[2];
    int np[2] = {ns, nt};
    double mean[2][3];
    double scale = 0.0;
    pc[0] = (double *)malloc((size_t)(ns > 0 ? ns : 1) * 3 * sizeof(double));
    pc[1] = (double *)malloc((size_t)(nt > 0 ? nt : 1) * 3 * sizeof(double));
    memcpy(pc[0], source, (size_t)ns * 3 * sizeof(double));
    memcpy(pc[1], target, (size_t)nt * 3 * sizeof(double));
    for (int i = 0; i < 2; ++i) {
        double m[3] = {0, 0, 0};
        for (int ii = 0; ii < np[i]; ++ii) { m[0] += pc[i][ii * 3]; m[1] += pc[i][ii * 3 + 1]; m[2] += pc[i][ii * 3 + 2]; }
        m[0] /= np[i]; m[1] /= np[i]; m[2] /= np[i];
        memcpy(mean[i], m, sizeof(m));
        for (int ii = 0; ii < np[i]; ++ii) { pc[i][ii * 3] -= m[0]; pc[i][ii * 3 + 1] -= m[1]; pc[i][ii * 3 + 2] -= m[2]; }
        for (int ii = 0; ii < np[i]; ++ii) { const double t = norm3(pc[i] + (size_t)ii * 3); if (t > scale) scale = t; }
    }
    const double scale_global = scale, scale_start = 1.0;
    for (int i = 0; i < 2; ++i)
        for (int e = 0; e < np[i] * 3; ++e) pc[i][e] /= scale_global;

    /* ---- InitialMatching: nearest source feature of every target feature, and back; cross check */
    int *ji = (int *)malloc((size_t)(nft > 0 ? nft : 1) * sizeof(int));
    int *ij = (int *)malloc((size_t)(nfs > 0 ? nfs : 1) * sizeof(int));
    double *dtmp = (double *)malloc((size_t)((nfs > nft ? nfs : nft) + 1) * sizeof(double));
    oracle_knn(feat_s, nfs, 33, feat_t, nft, 1, 0, ji, dtmp);
    oracle_knn(feat_t, nft, 33, feat_s, nfs, 1, 0, ij, dtmp);
    int (*corres)[2] = (int (*)[2])malloc((size_t)(nfs > 0 ? nfs : 1) * sizeof(int[2]));
    int ncorr = 0;
    for (int i = 0; i < nfs; ++i) {          /* std::map order = ascending i */
        const int j = ij[i];
        if (j < 0 || ji[j] != i) continue;
        if (i >= ns || j >= nt) continue;    /* out-of-range in Open3D (UB); dropped here */
        corres[ncorr][0] = i; corres[ncorr][1] = j; ++ncorr;
    }
    const int n_initial = ncorr;

    /* ---- AdvancedMatching: swap so that `fi` is the larger cloud, tuple test, swap back */
    int fi = 0, fj = 1, swapped = 0;
    if (np[fj] > np[fi]) {
        fi = 1; fj = 0; swapped = 1;
        for (int c = 0; c < ncorr; ++c) { const int t = corres[c][0]; corres[c][0] = corres[c][1]; corres[c][1] = t; }
    }
    int trials_run = 0;
    if (tuple_test) {
        int (*tup)[2] = (int (*)[2])malloc((size_t)(3 * (max_tuple_count > 0 ? max_tuple_count : 1)) * sizeof(int[2]));
        int cnt = 0;
        const long long number_of_trial = (long long)ncorr * 100;
        for (long long t = 0; t < number_of_trial && cnt < max_tuple_count; ++t) {
            uint32_t r[4] = {(uint32_t)t, (uint32_t)((uint64_t)t >> 32), 0u, 0u};
            philox4x32_10(r, (uint32_t)tuple_key, (uint32_t)(tuple_key >> 32));
            const int r0 = (int)(((uint64_t)r[0] * (uint64_t)ncorr) >> 32);
            const int r1 = (int)(((uint64_t)r[1] * (uint64_t)ncorr) >> 32);
            const int r2 = (int)(((uint64_t)r[2] * (uint64_t)ncorr) >> 32);
            const int idi0 = corres[r0][0], idj0 = corres[r0][1];
            const int idi1 = corres[r1][0], idj1 = corres[r1][1];
            const int idi2 = corres[r2][0], idj2 = corres[r2][1];
            const double *pi0 = pc[fi] + (size_t)idi0 * 3, *pi1 = pc[fi] + (size_t)idi1 * 3, *pi2 = pc[fi] + (size_t)idi2 * 3;
            const double *pj0 = pc[fj] + (size_t)idj0 * 3, *pj1 = pc[fj] + (size_t)idj1 * 3, *pj2 = pc[fj] + (size_t)idj2 * 3;
            const double li0 = dist3(pi0, pi1), li1 = dist3(pi1, pi2), li2 = dist3(pi2, pi0);
            const double lj0 = dist3(pj0, pj1), lj1 = dist3(pj1, pj2), lj2 = dist3(pj2, pj0);
            ++trials_run;
            if ((li0 * tuple_scale < lj0) && (lj0 < li0 / tuple_scale) && (li1 * tuple_scale < lj1) &&
                (lj1 < li1 / tuple_scale) && (li2 * tuple_scale < lj2) && (lj2 < li2 / tuple_scale)) {
                tup[cnt * 3][0] = idi0; tup[cnt * 3][1] = idj0;
                tup[cnt * 3 + 1][0] = idi1; tup[cnt * 3 + 1][1] = idj1;
                tup[cnt * 3 + 2][0] = idi2; tup[cnt * 3 + 2][1] = idj2;
                ++cnt;
            }
        }
        free(corres);
        corres = tup;
        ncorr = cnt * 3;
    }
    if (swapped)
        for (int c = 0; c < ncorr; ++c) { const int t = corres[c][0]; corres[c][0] = corres[c][1]; corres[c][1] = t; }

    /* ---- OptimizePairwiseRegistration: graduated non-convexity, Geman-McClure weights */
    double trans[16];
    for (int k = 0; k < 16; ++k) trans[k] = (k % 5 == 0) ? 1.0 : 0.0;
    if (ncorr >= 10) {
        double par = scale_start;
        double *q = (double *)malloc((size_t)nt * 3 * sizeof(double));   /* moving copy of cloud 1 */
        memcpy(q, pc[1], (size_t)nt * 3 * sizeof(double));
        for (int itr = 0; itr < iteration_number; ++itr) {
            double JTJ[36], JTr[6];
            for (int k = 0; k < 36; ++k) JTJ[k] = 0.0;
            for (int k = 0; k < 6; ++k) JTr[k] = 0.0;
            for (int c = 0; c < ncorr; ++c) {
                const double *pp = pc[0] + (size_t)corres[c][0] * 3;
                const double *qq = q + (size_t)corres[c][1] * 3;
                const double rpq[3] = {pp[0] - qq[0], pp[1] - qq[1], pp[2] - qq[2]};
                const double temp = par / (dot3(rpq, rpq) + par);
                const double s = temp * temp;
                const double J[3][6] = {{0, -qq[2], qq[1], -1, 0, 0},
                                        {qq[2], 0, -qq[0], 0, -1, 0},
                                        {-qq[1], qq[0], 0, 0, 0, -1}};
                for (int r = 0; r < 3; ++r) {
                    for (int a = 0; a < 6; ++a) {
                        for (int b = 0; b < 6; ++b) JTJ[a * 6 + b] += J[r][a] * J[r][b] * s;
                        JTr[a] += J[r][a] * rpq[r] * s;
                    }
                }
            }
            double A[36], x[6], delta[16];
            for (int k = 0; k < 36; ++k) A[k] = -JTJ[k];
            ldlt_solve6(A, JTr, x);
            vec6_to_mat4(x, delta);
            mat4_mul(delta, trans, trans);
            transform_cloud(q, NULL, nt, delta);
            if (decrease_mu && itr % 4 == 0 && par > max_corr_dist) par /= division_factor;
        }
        free(q);
    }

    /* ---- GetTransformationOriginalScale, then its inverse (source -> target) */
    double tt[3];
    for (int r = 0; r < 3; ++r)
        tt[r] = -(trans[r * 4] * mean[1][0] + trans[r * 4 + 1] * mean[1][1] + trans[r * 4 + 2] * mean[1][2]) +
                trans[r * 4 + 3] * scale_global + mean[0][r];
    double T[16];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[r * 4 + c] = trans[c * 4 + r];          /* R^T */
        T[r * 4 + 3] = -(trans[0 * 4 + r] * tt[0] + trans[1 * 4 + r] * tt[1] + trans[2 * 4 + r] * tt[2]);
    }
    T[12] = T[13] = T[14] = 0.0; T[15] = 1.0;
    memcpy(T_out, T, sizeof(T));

    /* ---- EvaluateRegistration(source_orig, target_orig, max_corr_dist, T) */
    oracle_evaluate(source, ns, target, nt, max_corr_dist, T, fitness_out, rmse_out, n_corr_out, NULL);
    if (info_out) { info_out[0] = n_initial; info_out[1] = ncorr; info_out[2] = trials_run; info_out[3] = swapped; }
    if (corres_out) memcpy(corres_out, corres, (size_t)ncorr * sizeof(int[2]));
    free(corres); free(ji); free(ij); free(dtmp); free(pc[0]); free(pc[1]);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Down-samplers                                   */
/* ------------------------------------------------------------------------ */

/* Preprocessor/Downsamplers/farthestDownsampler.py:33-54: distances start at 1e6; per step
 * dist = min(dist, cdist(point, cloud)) (scipy euclidean: sqrt of the coordinate-order sum of
 * squares), next = np.argmax(dist) (first maximum).  out_idx[0] = first.                    */
ORACLE_API int oracle_fps(const double *pts, int n, int first, int k, int *out_idx)
{
    if (n <= 0 || k <= 0 || first < 0 || first >= n) return -1;
    double *dist = (double *)malloc((size_t)n * sizeof(double));
    for (int i = 0; i < n; ++i) dist[i] = 1e6;
    int cur = first;
    out_idx[0] = cur;
    for (int s = 0; s + 1 < k; ++s) {
        const double *p = pts + (size_t)cur * 3;
        int best = 0;
        double bv = -1.0;
        for (int i = 0; i < n; ++i) {
            const double d = sqrt(sqdist(pts + (size_t)i * 3, p, 3));
            if (d < dist[i]) dist[i] = d;
            if (dist[i] > bv) { bv = dist[i]; best = i; }
        }
        cur = best;
        out_idx[s + 1] = cur;
    }
    free(dist);
    return 0;
}

typedef struct { long long c[3]; int idx; } VoxelRec;
static int voxel_cmp(const void *a, const void *b)
{
    const VoxelRec *x = (const VoxelRec *)a, *y = (const VoxelRec *)b;
    for (int k = 0; k < 3; ++k) if (x->c[k] != y->c[k]) return x->c[k] < y->c[k] ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

/* PointCloud::VoxelDownSample(voxel_size) (voxelDownsampler.py:83,95): voxel index
 * floor((p - (min_bound - voxel_size/2)) / voxel_size), per-voxel average with the points added in
 * index order.  Output in lexicographic voxel-index order (Open3D: unordered_map order).
 * Returns the number of voxels; out (n x 3) holds that many points.                         */
ORACLE_API int oracle_voxel_downsample(const double *pts, int n, double voxel_size, double *out, int *first_idx)
{
    if (n <= 0 || !(voxel_size > 0)) return -1;
    double mn[3] = {INFINITY, INFINITY, INFINITY};
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) if (pts[(size_t)i * 3 + k] < mn[k]) mn[k] = pts[(size_t)i * 3 + k];
    double vmin[3];
    for (int k = 0; k < 3; ++k) vmin[k] = mn[k] - voxel_size * 0.5;
    VoxelRec *rec = (VoxelRec *)malloc((size_t)n * sizeof(VoxelRec));
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) rec[i].c[k] = (long long)floor((pts[(size_t)i * 3 + k] - vmin[k]) / voxel_size);
        rec[i].idx = i;
    }
    qsort(rec, n, sizeof(VoxelRec), voxel_cmp);
    int m = 0;
    for (int i = 0; i < n;) {
        int j = i;
        double s[3] = {0, 0, 0};
        while (j < n && rec[j].c[0] == rec[i].c[0] && rec[j].c[1] == rec[i].c[1] && rec[j].c[2] == rec[i].c[2]) {
            const double *p = pts + (size_t)rec[j].idx * 3;
            s[0] += p[0]; s[1] += p[1]; s[2] += p[2];
            ++j;
        }
        const double cnt = (double)(j - i);
        out[(size_t)m * 3] = s[0] / cnt; out[(size_t)m * 3 + 1] = s[1] / cnt; out[(size_t)m * 3 + 2] = s[2] / cnt;
        if (first_idx) first_idx[m] = rec[i].idx;
        ++m;
        i = j;
    }
    free(rec);
    return m;
}

ORACLE_API int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORACLE_API void oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
