// Per-point neighbourhood covariance -> normal -> GICP covariance.
//
// Replaces PointCloud::EstimateNormals (KNN or Hybrid search, fast closed-form
// eigenvector) and InitializePointCloudForGeneralizedICP of Open3D 0.18.0, as
// reached from Optimizer/generalizedICP.py:59-70 and
// Optimizer/fastGlobalOptimizer.py:118-127.
// One thread per point, float64 throughout: the one-pass covariance
// E[ab]-E[a]E[b] cancels catastrophically in float32.
#include "common.cuh"

int orpcd_launch_knn(const double *d_points, int n, long long points_stride, const double *d_queries,
                     int m, long long queries_stride, int batch, int k, int *d_idx, double *d_d2,
                     void *d_scratch, size_t scratch_bytes, cudaStream_t st);
size_t orpcd_knn_scratch_bytes(int n, int m, int batch, int k);

namespace {

__device__ __forceinline__ void cross3(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double *a, const double *b)
{
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

// Eigenvector of symmetric A for eigenvalue ev: best-conditioned row cross product.
__device__ void eigvec_from_rows(const double *A, double ev, double *out)
{
    const double r0[3] = {A[0] - ev, A[1], A[2]};
    const double r1[3] = {A[1], A[4] - ev, A[5]};
    const double r2[3] = {A[2], A[5], A[8] - ev};
    double c01[3], c02[3], c12[3];
    cross3(r0, r1, c01);
    cross3(r0, r2, c02);
    cross3(r1, r2, c12);
    const double d0 = dot3(c01, c01), d1 = dot3(c02, c02), d2 = dot3(c12, c12);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    const double *c = imax == 0 ? c01 : (imax == 1 ? c02 : c12);
    const double s = sqrt(imax == 0 ? d0 : (imax == 1 ? d1 : d2));
    out[0] = c[0] / s; out[1] = c[1] / s; out[2] = c[2] / s;
}

// Second eigenvector from the 2x2 problem in the plane orthogonal to e0.
__device__ void eigvec_in_plane(const double *A, const double *e0, double ev1, double *out)
{
    double U[3], V[3];
    if (fabs(e0[0]) > fabs(e0[1])) {
        const double inv = 1 / sqrt(e0[0] * e0[0] + e0[2] * e0[2]);
        U[0] = -e0[2] * inv; U[1] = 0; U[2] = e0[0] * inv;
    } else {
        const double inv = 1 / sqrt(e0[1] * e0[1] + e0[2] * e0[2]);
        U[0] = 0; U[1] = e0[2] * inv; U[2] = -e0[1] * inv;
    }
    cross3(e0, U, V);
    const double AU[3] = {A[0] * U[0] + A[1] * U[1] + A[2] * U[2], A[1] * U[0] + A[4] * U[1] + A[5] * U[2],
                          A[2] * U[0] + A[5] * U[1] + A[8] * U[2]};
    const double AV[3] = {A[0] * V[0] + A[1] * V[1] + A[2] * V[2], A[1] * V[0] + A[4] * V[1] + A[5] * V[2],
                          A[2] * V[0] + A[5] * V[1] + A[8] * V[2]};
    double m00 = U[0] * AU[0] + U[1] * AU[1] + U[2] * AU[2] - ev1;
    double m01 = U[0] * AV[0] + U[1] * AV[1] + U[2] * AV[2];
    double m11 = V[0] * AV[0] + V[1] * AV[1] + V[2] * AV[2] - ev1;
    const double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    if (a00 >= a11) {
        if (fmax(a00, a01) > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1 / sqrt(1 + m01 * m01); m01 *= m00; }
            else            { m00 /= m01; m01 = 1 / sqrt(1 + m00 * m00); m00 *= m01; }
            for (int i = 0; i < 3; ++i) out[i] = m01 * U[i] - m00 * V[i];
        } else {
            for (int i = 0; i < 3; ++i) out[i] = U[i];
        }
    } else {
        if (fmax(a11, a01) > 0) {
            if (a11 >= a01) { m01 /= m11; m11 = 1 / sqrt(1 + m01 * m01); m01 *= m11; }
            else            { m11 /= m01; m01 = 1 / sqrt(1 + m11 * m11); m11 *= m01; }
            for (int i = 0; i < 3; ++i) out[i] = m11 * U[i] - m01 * V[i];
        } else {
            for (int i = 0; i < 3; ++i) out[i] = U[i];
        }
    }
}

// Closed-form eigenvector of the smallest eigenvalue of the symmetric C (9, row major).
__device__ void smallest_eigvec(const double *C, double *n)
{
    double A[9];
    double mc = C[0];
    for (int i = 1; i < 9; ++i) mc = C[i] > mc ? C[i] : mc;
    if (mc == 0) { n[0] = n[1] = n[2] = 0; return; }
    for (int i = 0; i < 9; ++i) A[i] = C[i] / mc;
    const double norm = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (norm > 0) {
        const double q = (A[0] + A[4] + A[8]) / 3;
        const double b00 = A[0] - q, b11 = A[4] - q, b22 = A[8] - q;
        const double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
        const double c00 = b11 * b22 - A[5] * A[5];
        const double c01 = A[1] * b22 - A[5] * A[2];
        const double c02 = A[1] * A[5] - b11 * A[2];
        const double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
        double half_det = det * 0.5;
        half_det = fmin(fmax(half_det, -1.0), 1.0);
        const double angle = det_acos(half_det) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        const double beta2 = det_cos(angle) * 2;
        const double beta0 = det_cos(angle + two_thirds_pi) * 2;
        const double beta1 = -(beta0 + beta2);
        const double ev0 = q + p * beta0, ev1 = q + p * beta1, ev2 = q + p * beta2;
        double e0[3], e1[3], e2[3];
        if (half_det >= 0) {
            eigvec_from_rows(A, ev2, e2);
            if (ev2 < ev0 && ev2 < ev1) { n[0] = e2[0]; n[1] = e2[1]; n[2] = e2[2]; return; }
            eigvec_in_plane(A, e2, ev1, e1);
            if (ev1 < ev0 && ev1 < ev2) { n[0] = e1[0]; n[1] = e1[1]; n[2] = e1[2]; return; }
            cross3(e1, e2, n);
        } else {
            eigvec_from_rows(A, ev0, e0);
            if (ev0 < ev1 && ev0 < ev2) { n[0] = e0[0]; n[1] = e0[1]; n[2] = e0[2]; return; }
            eigvec_in_plane(A, e0, ev1, e1);
            if (ev1 < ev0 && ev1 < ev2) { n[0] = e1[0]; n[1] = e1[1]; n[2] = e1[2]; return; }
            cross3(e0, e1, n);
        }
    } else {
        const double a0 = A[0] * mc, a1 = A[4] * mc, a2 = A[8] * mc;
        if (a0 < a1 && a0 < a2) { n[0] = 1; n[1] = 0; n[2] = 0; }
        else if (a1 < a0 && a1 < a2) { n[0] = 0; n[1] = 1; n[2] = 0; }
        else { n[0] = 0; n[1] = 0; n[2] = 1; }
    }
}

// C = Rx diag(eps,1,1) Rx^T with Rx the rotation taking e1 onto the normal
// (identity when the normal is within ~8 degrees of -e1), upper triangle.
__device__ void gicp_cov_from_normal(const double *nrm, double eps, double *c6, double *axis)
{
    double R[9];
    const double c = nrm[0];
    const double v[3] = {0.0, -nrm[2], nrm[1]};
    if (c < -0.99) {
        for (int i = 0; i < 9; ++i) R[i] = 0;
        R[0] = R[4] = R[8] = 1;
    } else {
        const double sv[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
        double sv2[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                sv2[i * 3 + j] = sv[i * 3 + 0] * sv[0 * 3 + j] + sv[i * 3 + 1] * sv[1 * 3 + j] +
                                 sv[i * 3 + 2] * sv[2 * 3 + j];
        const double f = 1 / (1 + c);
        for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + sv[i] + sv2[i] * f;
    }
    if (axis) { axis[0] = R[0]; axis[1] = R[3]; axis[2] = R[6]; }  // R e1
    if (!c6) return;
    double RD[9];
    for (int i = 0; i < 3; ++i) { RD[i * 3] = R[i * 3] * eps; RD[i * 3 + 1] = R[i * 3 + 1]; RD[i * 3 + 2] = R[i * 3 + 2]; }
    int o = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j)
            c6[o++] = RD[i * 3] * R[j * 3] + RD[i * 3 + 1] * R[j * 3 + 1] + RD[i * 3 + 2] * R[j * 3 + 2];
}

__global__ void __launch_bounds__(128)
normals_cov_kernel(const double *__restrict__ points, int n, long long points_stride,
                   const int *__restrict__ nn_idx, const double *__restrict__ nn_d2, int knn,
                   double radius2, double gicp_eps, double *__restrict__ normals,
                   double *__restrict__ cov6, double *__restrict__ gicp_cov6,
                   double *__restrict__ gicp_axis, int *__restrict__ nn_out, const int *__restrict__ order)
{
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = order ? order[t] : t;     // spatial processing order: neighbouring threads gather the same points (L1)
    const double *pts = points + (size_t)b * points_stride;
    const size_t row = ((size_t)b * n + i);
    const int *idx = nn_idx + row * knn;
    const double *d2 = nn_d2 + row * knn;
    // cumulants in neighbour (ascending distance) order
    double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (int j = 0; j < knn; ++j) {
        const int id = idx[j];
        if (id < 0) break;
        if (radius2 > 0 && !(d2[j] < radius2)) break;
        const double x = pts[(size_t)id * 3], y = pts[(size_t)id * 3 + 1], z = pts[(size_t)id * 3 + 2];
        c[0] += x; c[1] += y; c[2] += z;
        c[3] += x * x; c[4] += x * y; c[5] += x * z;
        c[6] += y * y; c[7] += y * z; c[8] += z * z;
        ++cnt;
    }
    if (nn_out)
        for (int j = 0; j < knn; ++j) nn_out[row * knn + j] = j < cnt ? idx[j] : -1;
    double C[9];
    if (cnt >= 3) {
        const double nn = (double)cnt;
        for (int j = 0; j < 9; ++j) c[j] /= nn;
        C[0] = c[3] - c[0] * c[0];
        C[4] = c[6] - c[1] * c[1];
        C[8] = c[8] - c[2] * c[2];
        C[1] = C[3] = c[4] - c[0] * c[1];
        C[2] = C[6] = c[5] - c[0] * c[2];
        C[5] = C[7] = c[7] - c[1] * c[2];
    } else {
        for (int j = 0; j < 9; ++j) C[j] = 0;
        C[0] = C[4] = C[8] = 1;
    }
    double nr[3];
    smallest_eigvec(C, nr);
    if (sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]) == 0.0) { nr[0] = 0; nr[1] = 0; nr[2] = 1; }
    if (normals) { normals[row * 3] = nr[0]; normals[row * 3 + 1] = nr[1]; normals[row * 3 + 2] = nr[2]; }
    if (cov6) {
        double *o = cov6 + row * 6;
        o[0] = C[0]; o[1] = C[1]; o[2] = C[2]; o[3] = C[4]; o[4] = C[5]; o[5] = C[8];
    }
    if (gicp_cov6 || gicp_axis)
        gicp_cov_from_normal(nr, gicp_eps, gicp_cov6 ? gicp_cov6 + row * 6 : nullptr,
                             gicp_axis ? gicp_axis + row * 3 : nullptr);
}

__global__ void apply_scale_kernel(const double *__restrict__ pts, int n, const double *__restrict__ scales,
                                   double *__restrict__ out)
{
    const int s = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * 3) return;
    out[(size_t)s * n * 3 + e] = __dmul_rn(pts[e], scales[s * 3 + e % 3]);
}

}  // namespace

ORPCD_API size_t orpcd_normals_workspace_bytes(int n, int batch, int knn)
{
    if (n <= 0 || batch <= 0 || knn < 1) return 0;
    const size_t nn = (size_t)n * batch * knn;
    return orpcd_align_up(nn * sizeof(double), 256) + orpcd_align_up(nn * sizeof(int), 256) +
           orpcd_knn_scratch_bytes(n, n, batch, knn);
}

ORPCD_API int orpcd_normals_cov(const double *d_points, int n, long long points_stride, int batch,
                                int knn, double radius, double gicp_eps, void *d_workspace,
                                size_t workspace_bytes, double *d_normals, double *d_cov6,
                                double *d_gicp_cov6, double *d_gicp_axis, int *d_nn_idx,
                                orpcd_stream_t stream)
{
    if (!d_points || n <= 0 || batch <= 0 || knn < 1 || knn > ORPCD_MAX_KNN) return ORPCD_ERR_INVALID;
    if ((d_gicp_cov6 || d_gicp_axis) && !(gicp_eps > 0)) return ORPCD_ERR_INVALID;
    if (!d_workspace || workspace_bytes < orpcd_normals_workspace_bytes(n, batch, knn) ||
        ((uintptr_t)d_workspace & 255))
        return ORPCD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nn = (size_t)n * batch * knn;
    double *d2 = (double *)d_workspace;
    int *idx = (int *)((unsigned char *)d_workspace + orpcd_align_up(nn * sizeof(double), 256));
    unsigned char *scratch = (unsigned char *)idx + orpcd_align_up(nn * sizeof(int), 256);
    int rc = orpcd_launch_knn(d_points, n, points_stride, d_points, n, points_stride, batch, knn, idx, d2, scratch,
                              orpcd_knn_scratch_bytes(n, n, batch, knn), st);
    if (rc != ORPCD_OK) return rc;
    dim3 grid((n + 127) / 128, batch);
    normals_cov_kernel<<<grid, 128, 0, st>>>(d_points, n, points_stride, idx, d2, knn,
                                             radius > 0 ? radius * radius : 0.0, gicp_eps, d_normals,
                                             d_cov6, d_gicp_cov6, d_gicp_axis, d_nn_idx, nullptr);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API int orpcd_normals_cov_from_knn(const double *d_points, int n, long long points_stride, int batch,
                                         int knn, const int *d_knn_idx, const double *d_knn_d2, double radius,
                                         double gicp_eps, double *d_normals, double *d_cov6,
                                         double *d_gicp_cov6, double *d_gicp_axis, int *d_nn_idx,
                                         const int *d_order, orpcd_stream_t stream)
{
    if (!d_points || !d_knn_idx || !d_knn_d2 || n <= 0 || batch <= 0 || knn < 1) return ORPCD_ERR_INVALID;
    if ((d_gicp_cov6 || d_gicp_axis) && !(gicp_eps > 0)) return ORPCD_ERR_INVALID;
    dim3 grid((n + 127) / 128, batch);
    normals_cov_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d_points, n, points_stride, d_knn_idx, d_knn_d2, knn,
                                                               radius > 0 ? radius * radius : 0.0, gicp_eps,
                                                               d_normals, d_cov6, d_gicp_cov6, d_gicp_axis, d_nn_idx, d_order);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

// out[u][i] = points[i] @ R_u + t_u (row-vector convention of Aligner/Aligner.py:183-185) with the arithmetic of
// numpy's np.dot(source, rotation) + translation: OpenBLAS' dgemm micro-kernels accumulate over k with fused
// multiply-adds, acc = x R0j; acc = fma(y, R1j, acc); acc = fma(z, R2j, acc), then one rounded add of t_j
// (checked bit for bit against numpy in tests/ and by the Aligner before it relies on it).  One thread per point.
__global__ void __launch_bounds__(256)
apply_rigid_kernel(const double *__restrict__ pts, int n, const double *__restrict__ rot, const double *__restrict__ trans,
                   double *__restrict__ out)
{
    const int u = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *R = rot + (size_t)u * 9, *t = trans + (size_t)u * 3;
    const double x = pts[(size_t)i * 3], y = pts[(size_t)i * 3 + 1], z = pts[(size_t)i * 3 + 2];
    double *o = out + ((size_t)u * n + i) * 3;
#pragma unroll
    for (int j = 0; j < 3; ++j)
        o[j] = __dadd_rn(fma(z, R[6 + j], fma(y, R[3 + j], __dmul_rn(x, R[j]))), t[j]);
}

ORPCD_API int orpcd_apply_rigid_batch(const double *d_points, int n, const double *d_rot, const double *d_trans,
                                      int n_units, double *d_out, orpcd_stream_t stream)
{
    if (!d_points || !d_rot || !d_trans || !d_out || n < 0 || n_units < 0 || n_units > 65535) return ORPCD_ERR_INVALID;
    if (n == 0 || n_units == 0) return ORPCD_OK;
    apply_rigid_kernel<<<dim3((n + 255) / 256, n_units), 256, 0, (cudaStream_t)stream>>>(d_points, n, d_rot, d_trans, d_out);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API int orpcd_apply_scale_batch(const double *d_points, int n, const double *d_scales,
                                      int n_scales, double *d_out, orpcd_stream_t stream)
{
    if (!d_points || !d_scales || !d_out || n < 0 || n_scales < 0) return ORPCD_ERR_INVALID;
    if (n == 0 || n_scales == 0) return ORPCD_OK;
    dim3 grid((n * 3 + 255) / 256, n_scales);
    apply_scale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_points, n, d_scales, d_out);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}
