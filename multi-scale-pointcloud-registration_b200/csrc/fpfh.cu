// FPFH descriptors (Rusu et al.) as Open3D 0.18.0 computes them, float64.
//
// Replaces open3d.pipelines.registration.compute_fpfh_feature(cloud,
// KDTreeSearchParamHybrid(radius, max_nn)) as called from
// Optimizer/fastGlobalOptimizer.py:130-143.  Algorithm: SURVEY.md Appendix A.7
// (Feature.cpp: ComputePairFeatures, ComputeSPFHFeature, ComputeFPFHFeature).
//
// Both passes run over neighbour lists computed elsewhere (orpcd_knn_exact /
// orpcd_knn_rerank: exact float64, sorted by (distance, index)); the Hybrid search is
// the k-NN list truncated at d2 < radius^2.  One thread per point; every sum is taken in
// neighbour order like the CPU code; the libm calls that feed discontinuous decisions (the acos
// comparison that picks the pair's source point, the atan2 whose result is binned by floor()) use the
// deterministic det_acos / det_atan2 shared with the oracle: SPFH and FPFH are bit-identical to it.
#include "common.cuh"

namespace {

constexpr int FP_THREADS = 128;
constexpr int FP_DIM = 33;
// SPFH rows are stored padded: 3 histograms x 12 doubles (11 bins + a zero), 288 bytes = nine 32-byte sectors, so that
// the FPFH pass -- one gather of a neighbour's histogram per (neighbour, histogram), bound by the L1's tag rate --
// reads a histogram as six aligned 16-byte loads (three sectors) instead of eleven 8-byte loads of a row that
// starts at any multiple of 8.
constexpr int SP_HIST = 12;
constexpr int SP_ROW = 3 * SP_HIST;     // 36 doubles

__device__ __forceinline__ int hist_bin(double x)
{
    int h = (int)floor(x);
    if (h < 0) h = 0;
    if (h >= 11) h = 10;
    return h;
}

// ComputePairFeatures(p1, n1, p2, n2): f0 = atan2 angle, f1, f2; returns false for the zero vector.
__device__ __forceinline__ bool pair_features(const double *p1, const double *n1, const double *p2,
                                              const double *n2, double &f0, double &f1, double &f2)
{
    double dx = p2[0] - p1[0], dy = p2[1] - p1[1], dz = p2[2] - p1[2];
    f0 = f1 = f2 = 0.0;
    const double len = sqrt(dx * dx + dy * dy + dz * dz);
    if (len == 0.0) return false;
    double ax = n1[0], ay = n1[1], az = n1[2], bx = n2[0], by = n2[1], bz = n2[2];
    const double angle1 = (ax * dx + ay * dy + az * dz) / len;
    const double angle2 = (bx * dx + by * dy + bz * dz) / len;
    double g2;
    if (det_acos(fabs(angle1)) > det_acos(fabs(angle2))) {
        ax = n2[0]; ay = n2[1]; az = n2[2];
        bx = n1[0]; by = n1[1]; bz = n1[2];
        dx *= -1.0; dy *= -1.0; dz *= -1.0;
        g2 = -angle2;
    } else {
        g2 = angle1;
    }
    // v = dp x a
    double vx = dy * az - dz * ay, vy = dz * ax - dx * az, vz = dx * ay - dy * ax;
    const double vn = sqrt(vx * vx + vy * vy + vz * vz);
    if (vn == 0.0) return false;
    vx /= vn; vy /= vn; vz /= vn;
    // w = a x v
    const double wx = ay * vz - az * vy, wy = az * vx - ax * vz, wz = ax * vy - ay * vx;
    f2 = g2;
    f1 = vx * bx + vy * by + vz * bz;
    f0 = det_atan2(wx * bx + wy * by + wz * bz, ax * bx + ay * by + az * bz);
    return true;
}

// SPFH: 3 x 11-bin histograms of the pair features to the neighbours (self excluded).
__global__ void __launch_bounds__(FP_THREADS)
spfh_kernel(const double *__restrict__ points, const double *__restrict__ normals, int n, long long stride,
            const int *__restrict__ nn_idx, const double *__restrict__ nn_d2, int list_k, int knn, double radius2,
            double *__restrict__ spfh)
{
    __shared__ double hist[FP_DIM][FP_THREADS];   // bin-major: conflict free
    const int b = blockIdx.y, tid = threadIdx.x;
    const int i = blockIdx.x * FP_THREADS + tid;
    if (i >= n) return;
    const double PI = 3.14159265358979323846;
    const double *pts = points + (size_t)b * stride;
    const double *nrm = normals + (size_t)b * stride;
    const size_t row = (size_t)b * n + i;
    const int *idx = nn_idx + row * list_k;
    const double *d2 = nn_d2 + row * list_k;
#pragma unroll
    for (int j = 0; j < FP_DIM; ++j) hist[j][tid] = 0.0;
    int cnt = 0;
    for (; cnt < knn; ++cnt) {
        if (idx[cnt] < 0) break;
        if (radius2 > 0 && !(d2[cnt] < radius2)) break;
    }
    if (cnt > 1) {
        const double incr = 100.0 / (double)(cnt - 1);
        const double *p1 = pts + (size_t)i * 3, *n1 = nrm + (size_t)i * 3;
        for (int k = 1; k < cnt; ++k) {
            const int id = idx[k];
            double f0, f1, f2;
            pair_features(p1, n1, pts + (size_t)id * 3, nrm + (size_t)id * 3, f0, f1, f2);
            hist[hist_bin(11 * (f0 + PI) / (2.0 * PI))][tid] += incr;
            hist[11 + hist_bin(11 * (f1 + 1.0) * 0.5)][tid] += incr;
            hist[22 + hist_bin(11 * (f2 + 1.0) * 0.5)][tid] += incr;
        }
    }
    double *o = spfh + row * SP_ROW;
#pragma unroll
    for (int h = 0; h < 3; ++h) {
#pragma unroll
        for (int j = 0; j < 11; ++j) o[h * SP_HIST + j] = hist[h * 11 + j][tid];
        o[h * SP_HIST + 11] = 0.0;
    }
}

// FPFH: own SPFH + the 1/d2-weighted, per-histogram renormalised SPFH of the neighbours.
__global__ void __launch_bounds__(FP_THREADS, 8)
fpfh_kernel(int n, const int *__restrict__ nn_idx, const double *__restrict__ nn_d2, int list_k, int knn,
            double radius2, const double *__restrict__ spfh, double *__restrict__ fpfh)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * FP_THREADS + threadIdx.x;
    if (i >= n) return;
    const size_t row = (size_t)b * n + i;
    const int *idx = nn_idx + row * list_k;
    const double *d2 = nn_d2 + row * list_k;
    const double *sp = spfh + (size_t)b * n * SP_ROW;
    int cnt = 0;
    for (; cnt < knn; ++cnt) {
        if (idx[cnt] < 0) break;
        if (radius2 > 0 && !(d2[cnt] < radius2)) break;
    }
    double *o = fpfh + row * FP_DIM;
    if (cnt > 1) {
        const double *self = sp + (size_t)i * SP_ROW;
        // one pass per histogram (the three are independent, each with its own normaliser): 11 running sums in
        // registers instead of 33 double the resident warps of a kernel that waits on the gathers of the rows
#pragma unroll 1
        for (int h = 0; h < 3; ++h) {
            double acc[11];
#pragma unroll
            for (int j = 0; j < 11; ++j) acc[j] = 0.0;
            double sum = 0.0;
            for (int k = 1; k < cnt; ++k) {
                const double dist = d2[k];
                if (dist == 0.0) continue;
                // the neighbour's histogram h: 96 bytes, 32-byte aligned
                const double2 *s2 = reinterpret_cast<const double2 *>(sp + (size_t)idx[k] * SP_ROW + h * SP_HIST);
                double s[12];
#pragma unroll
                for (int j = 0; j < 6; ++j) { const double2 t2 = __ldg(s2 + j); s[2 * j] = t2.x; s[2 * j + 1] = t2.y; }
                // 11 correctly rounded quotients s[j] / dist from ONE division: with y = RN(1 / dist), q = s y,
                // r = s - q dist (exact in an FMA), q' = q + r y is RN(s / dist) (Markstein) unless dist's
                // significand is all ones -- that case takes the plain division.  Histogram values and distances
                // are far from the overflow / subnormal ranges.
                const double y = 1.0 / dist;
                const bool plain = (__double2hiint(dist) & 0x000fffff) == 0x000fffff && (unsigned)__double2loint(dist) == 0xffffffffu;
                // (the branch sits OUTSIDE the loop over the bins: written as a select per bin, the compiler evaluated
                // `sj / dist` for every bin, and div.rn.f64 leaves its inline fast path for a ~70-instruction
                // subroutine whenever the numerator is zero -- most histogram bins are: 63 % of the kernel's
                // instructions)
                if (!plain) {
#pragma unroll
                    for (int j = 0; j < 11; ++j) {
                        const double sj = s[j];
                        const double q = __dmul_rn(sj, y);
                        const double val = fma(fma(-q, dist, sj), y, q);
                        sum += val;
                        acc[j] += val;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 11; ++j) {
                        const double val = s[j] / dist;
                        sum += val;
                        acc[j] += val;
                    }
                }
            }
            if (sum != 0.0) sum = 100.0 / sum;
#pragma unroll
            for (int j = 0; j < 11; ++j) o[h * 11 + j] = acc[j] * sum + self[h * SP_HIST + j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < FP_DIM; ++j) o[j] = 0.0;
    }
}

}  // namespace

ORPCD_API int orpcd_fpfh(const double *d_points, const double *d_normals, int n, long long points_stride,
                         int batch, const int *d_nn_idx, const double *d_nn_d2, int list_k, int knn, double radius,
                         double *d_spfh, double *d_fpfh, orpcd_stream_t stream)
{
    if (!d_points || !d_normals || !d_nn_idx || !d_nn_d2 || !d_spfh || !d_fpfh) return ORPCD_ERR_INVALID;
    if (n <= 0 || batch <= 0 || knn < 1 || list_k < knn) return ORPCD_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((n + FP_THREADS - 1) / FP_THREADS, batch);
    const double r2 = radius > 0 ? radius * radius : 0.0;
    spfh_kernel<<<grid, FP_THREADS, 0, st>>>(d_points, d_normals, n, points_stride, d_nn_idx, d_nn_d2, list_k, knn, r2,
                                             d_spfh);
    ORPCD_CHECK_LAUNCH();
    fpfh_kernel<<<grid, FP_THREADS, 0, st>>>(n, d_nn_idx, d_nn_d2, list_k, knn, r2, d_spfh, d_fpfh);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}
