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
 *   PARITY UNPINNED: the reference ships no tests and no golden vectors for
 *   this path, and Open3D itself cannot be run here.  This oracle is pinned
 *   only against independent libraries (scipy cKDTree, numpy eigh/inv/sqrtm)
 *   and against the reference's own importable numpy blocks
 *   (tests/golden/make_golden.py).
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
        double angle = acos(half_det) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        double beta2 = cos(angle) * 2;
        double beta0 = cos(angle + two_thirds_pi) * 2;
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
