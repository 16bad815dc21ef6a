// Exact float64 k-nearest-neighbour search (brute force, shared-memory tiled)
// and the statistical-outlier-removal block built on it.
//
// Replaces open3d KDTreeFlann::SearchKNN / PointCloud::RemoveStatisticalOutliers
// as reached from Preprocessor/Outliers/sor.py:69-71 and the 20-NN covariance
// stage of registration_generalized_icp (Optimizer/generalizedICP.py:59-70).
// Every distance is computed in float64 with Open3D's operation order, so the
// (distance, index)-sorted result lists are bit-identical to the CPU path; the
// k > 1 searches are a few percent of a registration, the 1-NN correspondence
// sweep (gicp.cu) is the hot one and has its own FP32-filter kernel.
#include "common.cuh"

namespace {

constexpr int KNN_BLOCK = 128;  // queries per block (one per thread)
constexpr int KNN_TILE = 512;   // points staged per shared-memory tile

// One thread per query; its k best (d2, idx) pairs live in shared memory, slot-major so
// a warp's accesses to one slot are conflict free.  The set is kept UNSORTED with the
// position of its worst member cached: a better point overwrites the worst slot and the
// new worst is found by one scan (a sorted insert costs k dependent shifts, and any
// lane inserting makes the whole warp wait).  Points are visited in ascending index order
// and replace only on a strict '<', so the kept set is the k smallest under (d2, idx)
// lexicographic order; it is sorted once at the end.
__global__ void __launch_bounds__(KNN_BLOCK)
knn_f64_kernel(const double *__restrict__ points, int n, long long points_stride,
               const double *__restrict__ queries, int m, long long queries_stride, int k,
               int *__restrict__ out_idx, double *__restrict__ out_d2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *tx = reinterpret_cast<double *>(smem_raw);
    double *ty = tx + KNN_TILE;
    double *tz = ty + KNN_TILE;
    double *ld = tz + KNN_TILE;                          // [k][KNN_BLOCK]
    int *li = reinterpret_cast<int *>(ld + (size_t)k * KNN_BLOCK);  // [k][KNN_BLOCK]

    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const int qi = blockIdx.x * KNN_BLOCK + tid;
    const double *pts = points + (size_t)b * points_stride;
    const double *qs = queries + (size_t)b * queries_stride;
    const bool active = qi < m;
    double qx = 0, qy = 0, qz = 0;
    if (active) {
        qx = qs[(size_t)qi * 3];
        qy = qs[(size_t)qi * 3 + 1];
        qz = qs[(size_t)qi * 3 + 2];
    }
    int cnt = 0, wpos = 0;
    double worst = CUDART_INF;

    for (int base = 0; base < n; base += KNN_TILE) {
        const int tn = min(KNN_TILE, n - base);
        __syncthreads();
        for (int e = tid; e < tn * 3; e += KNN_BLOCK) {
            const double v = pts[(size_t)base * 3 + e];
            const int p = e / 3, c = e - p * 3;
            (c == 0 ? tx : (c == 1 ? ty : tz))[p] = v;
        }
        __syncthreads();
        if (!active) continue;
        for (int j0 = 0; j0 < tn; j0 += 4) {
            double dd[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int j = min(j0 + r, tn - 1);
                dd[r] = sqdist3_rn(qx, qy, qz, tx[j], ty[j], tz[j]);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int j = j0 + r;
                const double d = dd[r];
                if (j >= tn) continue;
                if (cnt < k) {
                    ld[(size_t)cnt * KNN_BLOCK + tid] = d;
                    li[(size_t)cnt * KNN_BLOCK + tid] = base + j;
                    ++cnt;
                    if (cnt < k) continue;
                } else if (d < worst) {
                    ld[(size_t)wpos * KNN_BLOCK + tid] = d;
                    li[(size_t)wpos * KNN_BLOCK + tid] = base + j;
                } else {
                    continue;
                }
                // rescan for the lexicographically largest (d2, idx) member
                double wd = ld[tid];
                int wi = li[tid], wp = 0;
                for (int s = 1; s < k; ++s) {
                    const double sd = ld[(size_t)s * KNN_BLOCK + tid];
                    const int si = li[(size_t)s * KNN_BLOCK + tid];
                    if (sd > wd || (sd == wd && si > wi)) { wd = sd; wi = si; wp = s; }
                }
                worst = wd;
                wpos = wp;
            }
        }
    }
    if (!active) return;
    // sort the kept set by (d2, idx): insertion sort in shared memory
    for (int a = 1; a < cnt; ++a) {
        const double d = ld[(size_t)a * KNN_BLOCK + tid];
        const int id = li[(size_t)a * KNN_BLOCK + tid];
        int pos = a;
        while (pos > 0) {
            const double pd = ld[(size_t)(pos - 1) * KNN_BLOCK + tid];
            const int pj = li[(size_t)(pos - 1) * KNN_BLOCK + tid];
            if (!(d < pd || (d == pd && id < pj))) break;
            ld[(size_t)pos * KNN_BLOCK + tid] = pd;
            li[(size_t)pos * KNN_BLOCK + tid] = pj;
            --pos;
        }
        ld[(size_t)pos * KNN_BLOCK + tid] = d;
        li[(size_t)pos * KNN_BLOCK + tid] = id;
    }
    int *oi = out_idx + ((size_t)b * m + qi) * k;
    double *od = out_d2 + ((size_t)b * m + qi) * k;
    for (int s = 0; s < k; ++s) {
        oi[s] = s < cnt ? li[(size_t)s * KNN_BLOCK + tid] : -1;
        od[s] = s < cnt ? ld[(size_t)s * KNN_BLOCK + tid] : CUDART_INF;
    }
}

size_t knn_smem_bytes(int k)
{
    return (size_t)3 * KNN_TILE * sizeof(double) + (size_t)k * KNN_BLOCK * (sizeof(double) + sizeof(int));
}

// avg_i = sum_j sqrt(d2_ij) / cnt_i  in ascending-distance order (A.2)
__global__ void sor_avg_kernel(const double *__restrict__ d2, int n, int k, double *__restrict__ avg)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *row = d2 + (size_t)i * k;
    double s = 0.0;
    int cnt = 0;
    for (int j = 0; j < k; ++j) {
        const double v = row[j];
        if (v == CUDART_INF) break;
        s = __dadd_rn(s, __dsqrt_rn(v));
        ++cnt;
    }
    avg[i] = cnt > 0 ? __ddiv_rn(s, (double)cnt) : -1.0;
}

// Sequential, index-ordered mean / Bessel std / threshold: bit-exact with the
// CPU loop (the strict '<' against the threshold decides the mask).
__global__ void sor_stats_kernel(const double *__restrict__ avg, int n, double std_ratio,
                                 double *__restrict__ stats)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    long long valid = 0;
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        const double a = avg[i];
        if (a > 0) { ++valid; sum = __dadd_rn(sum, a); }
    }
    if (valid == 0) { stats[0] = 0; stats[1] = 0; stats[2] = -1.0; stats[3] = 0; return; }
    const double mean = __ddiv_rn(sum, (double)valid);
    double sq = 0.0;
    for (int i = 0; i < n; ++i) {
        const double a = avg[i];
        if (a > 0) { const double d = __dsub_rn(a, mean); sq = __dadd_rn(sq, __dmul_rn(d, d)); }
    }
    const double sd = __dsqrt_rn(__ddiv_rn(sq, (double)(valid - 1)));
    stats[0] = mean;
    stats[1] = sd;
    stats[2] = __dadd_rn(mean, __dmul_rn(std_ratio, sd));
    stats[3] = 0;
}

__global__ void sor_mask_kernel(const double *__restrict__ avg, int n, double *__restrict__ stats,
                                unsigned char *__restrict__ mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double thr = stats[2];
    int keep = 0;
    if (i < n) {
        const double a = avg[i];
        keep = (a > 0 && a < thr) ? 1 : 0;
        mask[i] = (unsigned char)keep;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&stats[3], (double)__popc(bal));
}


// ---------------------------------------------------------------------------
// k-NN of a moved copy of a cloud from the neighbour lists of the base cloud.
//
// The multi-start loop registers `attempts` rigid images of ONE source cloud
// (Aligner/Aligner.py:179-185) and Open3D recomputes the 20-NN graph of each.
// A rigid motion keeps every pair distance up to rounding, so the k nearest
// neighbours of the moved cloud lie inside the kc > k nearest of the base
// cloud; each thread re-ranks those kc candidates with exact float64 distances
// of the MOVED coordinates.  The shortcut is certified per point: every
// non-candidate has base d2 >= bound_i, hence moved d2 >= factor * bound_i; if
// the k-th re-ranked distance is not strictly below that, the point goes to a
// work list and gets an exact scan.  The same holds for an anisotropically
// scaled cloud with factor = min(s)^2.
// ---------------------------------------------------------------------------
constexpr int RR_BLOCK = 128;

__global__ void __launch_bounds__(RR_BLOCK)
knn_rerank_kernel(const int *__restrict__ cand, const double *__restrict__ bound, int kc,
                  const double *__restrict__ points, int n, long long stride,
                  const double *__restrict__ factor, int k, int *__restrict__ out_idx,
                  double *__restrict__ out_d2, int *__restrict__ worklist, int *__restrict__ wcount)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *ld = reinterpret_cast<double *>(smem_raw);                  // [k][RR_BLOCK]
    int *li = reinterpret_cast<int *>(ld + (size_t)k * RR_BLOCK);       // [k][RR_BLOCK]
    const int b = blockIdx.y, tid = threadIdx.x;
    const int i = blockIdx.x * RR_BLOCK + tid;
    if (i >= n) return;
    const double *pts = points + (size_t)b * stride;
    const double qx = pts[(size_t)i * 3], qy = pts[(size_t)i * 3 + 1], qz = pts[(size_t)i * 3 + 2];
    const int *row = cand + (size_t)i * kc;
    int cnt = 0;
    for (int s = 0; s < kc; ++s) {
        const int j = row[s];
        if (j < 0) break;
        const double d = sqdist3_rn(qx, qy, qz, pts[(size_t)j * 3], pts[(size_t)j * 3 + 1], pts[(size_t)j * 3 + 2]);
        if (cnt == k) {
            const double wd = ld[(size_t)(k - 1) * RR_BLOCK + tid];
            const int wj = li[(size_t)(k - 1) * RR_BLOCK + tid];
            if (!(d < wd || (d == wd && j < wj))) continue;
        }
        int pos = (cnt < k) ? cnt++ : k - 1;
        while (pos > 0) {
            const double pd = ld[(size_t)(pos - 1) * RR_BLOCK + tid];
            const int pj = li[(size_t)(pos - 1) * RR_BLOCK + tid];
            if (!(d < pd || (d == pd && j < pj))) break;
            ld[(size_t)pos * RR_BLOCK + tid] = pd;
            li[(size_t)pos * RR_BLOCK + tid] = pj;
            --pos;
        }
        ld[(size_t)pos * RR_BLOCK + tid] = d;
        li[(size_t)pos * RR_BLOCK + tid] = j;
    }
    const size_t o = ((size_t)b * n + i) * k;
    for (int s = 0; s < k; ++s) {
        out_idx[o + s] = s < cnt ? li[(size_t)s * RR_BLOCK + tid] : -1;
        out_d2[o + s] = s < cnt ? ld[(size_t)s * RR_BLOCK + tid] : CUDART_INF;
    }
    // certificate: k-th distance strictly below anything a non-candidate can reach
    const double lim = factor[b] * bound[i];
    const bool ok = cnt == k && ld[(size_t)(k - 1) * RR_BLOCK + tid] < lim;
    if (!ok) worklist[atomicAdd(wcount, 1)] = b * n + i;
}

// Exact fallback, one warp per listed (cloud, point): k rounds of "smallest
// (d2, index) lexicographically above the previous pick" over the whole cloud.
__global__ void __launch_bounds__(128)
knn_rerank_fix_kernel(const double *__restrict__ points, int n, long long stride, int k,
                      int *__restrict__ out_idx, double *__restrict__ out_d2,
                      const int *__restrict__ worklist, const int *__restrict__ wcount)
{
    const int count = *wcount;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int w = gw; w < count; w += nw) {
        const int item = worklist[w];
        const int b = item / n, i = item - b * n;
        const double *pts = points + (size_t)b * stride;
        const double qx = pts[(size_t)i * 3], qy = pts[(size_t)i * 3 + 1], qz = pts[(size_t)i * 3 + 2];
        double pd = -1.0;
        int pj = -1;
        const size_t o = ((size_t)b * n + i) * k;
        for (int s = 0; s < k; ++s) {
            double best = CUDART_INF;
            int bj = 0x7fffffff;
            for (int j = lane; j < n; j += 32) {
                const double d = sqdist3_rn(qx, qy, qz, pts[(size_t)j * 3], pts[(size_t)j * 3 + 1], pts[(size_t)j * 3 + 2]);
                const bool above = d > pd || (d == pd && j > pj);
                if (above && (d < best || (d == best && j < bj))) { best = d; bj = j; }
            }
            for (int off = 16; off > 0; off >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, best, off);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
                if (od < best || (od == best && oj < bj)) { best = od; bj = oj; }
            }
            const bool found = bj != 0x7fffffff;
            if (lane == 0) {
                out_idx[o + s] = found ? bj : -1;
                out_d2[o + s] = found ? best : CUDART_INF;
            }
            if (!found) {
                for (int r = s + 1; r < k; ++r)
                    if (lane == 0) { out_idx[o + r] = -1; out_d2[o + r] = CUDART_INF; }
                break;
            }
            pd = best;
            pj = bj;
        }
    }
}

// bound_i = d2 of the last candidate of point i (+inf when the list is short:
// then it already holds every point of the cloud)
__global__ void knn_bound_kernel(const int *__restrict__ cand, const double *__restrict__ cand_d2, int n, int kc,
                                 double *__restrict__ bound)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bound[i] = cand[(size_t)i * kc + kc - 1] >= 0 ? cand_d2[(size_t)i * kc + kc - 1] : CUDART_INF;
}

}  // namespace

int orpcd_launch_knn(const double *d_points, int n, long long points_stride, const double *d_queries,
                     int m, long long queries_stride, int batch, int k, int *d_idx, double *d_d2,
                     cudaStream_t st)
{
    if (!d_points || !d_queries || !d_idx || !d_d2 || n < 0 || m < 0 || batch < 0 || k < 1 ||
        k > ORPCD_MAX_KNN)
        return ORPCD_ERR_INVALID;
    if (m == 0 || batch == 0) return ORPCD_OK;
    const size_t smem = knn_smem_bytes(k);
    if (cudaFuncSetAttribute(knn_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
        return ORPCD_ERR_CUDA;
    dim3 grid((m + KNN_BLOCK - 1) / KNN_BLOCK, batch);
    knn_f64_kernel<<<grid, KNN_BLOCK, smem, st>>>(d_points, n, points_stride, d_queries, m,
                                                  queries_stride, k, d_idx, d_d2);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API int orpcd_knn_exact(const double *d_points, int n, long long points_stride,
                              const double *d_queries, int m, long long queries_stride, int batch,
                              int k, int *d_idx, double *d_d2, orpcd_stream_t stream)
{
    return orpcd_launch_knn(d_points, n, points_stride, d_queries, m, queries_stride, batch, k, d_idx,
                            d_d2, (cudaStream_t)stream);
}

ORPCD_API size_t orpcd_sor_workspace_bytes(int n, int nb_neighbors)
{
    if (n <= 0 || nb_neighbors < 1) return 0;
    const size_t nn = (size_t)n * nb_neighbors;
    return orpcd_align_up(nn * sizeof(double), 256) + orpcd_align_up(nn * sizeof(int), 256) +
           orpcd_align_up((size_t)n * sizeof(double), 256);
}

ORPCD_API int orpcd_sor_mask(const double *d_points, int n, int nb_neighbors, double std_ratio,
                             void *d_workspace, size_t workspace_bytes, unsigned char *d_mask,
                             double *d_avg, double *d_stats, orpcd_stream_t stream)
{
    if (!d_points || !d_mask || !d_stats || n <= 0 || nb_neighbors < 1 ||
        nb_neighbors > ORPCD_MAX_KNN || !(std_ratio > 0))
        return ORPCD_ERR_INVALID;
    if (!d_workspace || workspace_bytes < orpcd_sor_workspace_bytes(n, nb_neighbors) ||
        ((uintptr_t)d_workspace & 255))
        return ORPCD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nn = (size_t)n * nb_neighbors;
    unsigned char *w = (unsigned char *)d_workspace;
    double *d2 = (double *)w;
    w += orpcd_align_up(nn * sizeof(double), 256);
    int *idx = (int *)w;
    w += orpcd_align_up(nn * sizeof(int), 256);
    double *avg = d_avg ? d_avg : (double *)w;
    int rc = orpcd_launch_knn(d_points, n, 0, d_points, n, 0, 1, nb_neighbors, idx, d2, st);
    if (rc != ORPCD_OK) return rc;
    const int tb = 128, gb = (n + tb - 1) / tb;
    sor_avg_kernel<<<gb, tb, 0, st>>>(d2, n, nb_neighbors, avg);
    ORPCD_CHECK_LAUNCH();
    sor_stats_kernel<<<1, 32, 0, st>>>(avg, n, std_ratio, d_stats);
    ORPCD_CHECK_LAUNCH();
    sor_mask_kernel<<<gb, tb, 0, st>>>(avg, n, d_stats, d_mask);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API size_t orpcd_knn_rerank_workspace_bytes(int n, int batch)
{
    if (n <= 0 || batch <= 0) return 0;
    return orpcd_align_up((size_t)n * batch * sizeof(int), 256) + 256;
}

ORPCD_API int orpcd_knn_candidate_bound(const int *d_cand_idx, const double *d_cand_d2, int n, int kc,
                                        double *d_bound, orpcd_stream_t stream)
{
    if (!d_cand_idx || !d_cand_d2 || !d_bound || n <= 0 || kc < 1) return ORPCD_ERR_INVALID;
    knn_bound_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_cand_idx, d_cand_d2, n, kc, d_bound);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API int orpcd_knn_rerank(const int *d_cand_idx, const double *d_cand_bound, int kc,
                               const double *d_points, int n, long long points_stride, int batch,
                               const double *d_factor, int k, int *d_idx, double *d_d2, void *d_workspace,
                               size_t workspace_bytes, orpcd_stream_t stream)
{
    if (!d_cand_idx || !d_cand_bound || !d_points || !d_factor || !d_idx || !d_d2 || n <= 0 || batch <= 0 ||
        k < 1 || k > ORPCD_MAX_KNN || kc < k || (long long)n * batch > 0x7fffffffLL)
        return ORPCD_ERR_INVALID;
    if (!d_workspace || workspace_bytes < orpcd_knn_rerank_workspace_bytes(n, batch) || ((uintptr_t)d_workspace & 255))
        return ORPCD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int *wcount = (int *)d_workspace;
    int *worklist = (int *)((unsigned char *)d_workspace + 256);
    if (cudaMemsetAsync(wcount, 0, sizeof(int), st) != cudaSuccess) return ORPCD_ERR_CUDA;
    const size_t smem = (size_t)k * RR_BLOCK * (sizeof(double) + sizeof(int));
    if (cudaFuncSetAttribute(knn_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return ORPCD_ERR_CUDA;
    dim3 grid((n + RR_BLOCK - 1) / RR_BLOCK, batch);
    knn_rerank_kernel<<<grid, RR_BLOCK, smem, st>>>(d_cand_idx, d_cand_bound, kc, d_points, n, points_stride, d_factor,
                                                    k, d_idx, d_d2, worklist, wcount);
    ORPCD_CHECK_LAUNCH();
    knn_rerank_fix_kernel<<<148 * 2, 128, 0, st>>>(d_points, n, points_stride, k, d_idx, d_d2, worklist, wcount);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}
