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
#include "sort.cuh"

namespace {

constexpr int KNN_TILE = 512;   // points staged per shared-memory tile
// queries per block (one per thread): the k-best sets take k * 12 bytes of shared memory per query, so the block
// shrinks as k grows (k <= 128: 128 queries, k <= 256: 64, k <= ORPCD_MAX_KNN = 512: 32)
__host__ __device__ constexpr int knn_block_for(int k) { return k <= 128 ? 128 : (k <= 256 ? 64 : 32); }

// One thread per query; its k best (d2, idx) pairs live in shared memory, slot-major so
// a warp's accesses to one slot are conflict free.  The set is kept UNSORTED with the
// position of its worst member cached: a better point overwrites the worst slot and the
// new worst is found by one scan (a sorted insert costs k dependent shifts, and any
// lane inserting makes the whole warp wait).  Points are visited in ascending index order
// and replace only on a strict '<', so the kept set is the k smallest under (d2, idx)
// lexicographic order; it is sorted once at the end.
template <int KNN_BLOCK>
__global__ void __launch_bounds__(KNN_BLOCK)
knn_f64_kernel(const double *__restrict__ points, int n, long long points_stride,
               const double *__restrict__ queries, int m, long long queries_stride, int k,
               int split_len, int *__restrict__ out_idx, double *__restrict__ out_d2)
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

    // blockIdx.z = split: this block scans points [p0, p1) and writes list (query, split); the lists of a
    // query are merged by knn_merge_kernel (the (d2, idx) order is total, so split + merge is exact)
    const int n_splits = gridDim.z, split = blockIdx.z;
    const int p0 = split * split_len, p1 = min(n, p0 + split_len);
    for (int base = p0; base < p1; base += KNN_TILE) {
        const int tn = min(KNN_TILE, p1 - base);
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
    int *oi = out_idx + (((size_t)b * m + qi) * n_splits + split) * k;
    double *od = out_d2 + (((size_t)b * m + qi) * n_splits + split) * k;
    for (int s = 0; s < k; ++s) {
        oi[s] = s < cnt ? li[(size_t)s * KNN_BLOCK + tid] : -1;
        od[s] = s < cnt ? ld[(size_t)s * KNN_BLOCK + tid] : CUDART_INF;
    }
}

// S sorted lists per query -> the k smallest under (d2, idx), sorted.  One thread per query.
constexpr int KNN_MAX_SPLITS = 8;
__global__ void __launch_bounds__(128)
knn_merge_kernel(const int *__restrict__ part_idx, const double *__restrict__ part_d2, long long n_queries, int n_splits,
                 int k, int *__restrict__ out_idx, double *__restrict__ out_d2)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    const int *pi = part_idx + (size_t)q * n_splits * k;
    const double *pd = part_d2 + (size_t)q * n_splits * k;
    int head[KNN_MAX_SPLITS], hi[KNN_MAX_SPLITS];
    double hd[KNN_MAX_SPLITS];
#pragma unroll
    for (int s = 0; s < KNN_MAX_SPLITS; ++s) {
        head[s] = 0;
        hi[s] = s < n_splits ? pi[(size_t)s * k] : -1;
        hd[s] = s < n_splits ? pd[(size_t)s * k] : CUDART_INF;
    }
    for (int r = 0; r < k; ++r) {
        int bs = -1, bi = -1;
        double bd = CUDART_INF;
#pragma unroll
        for (int s = 0; s < KNN_MAX_SPLITS; ++s)
            if (hi[s] >= 0 && (bs < 0 || hd[s] < bd || (hd[s] == bd && hi[s] < bi))) { bs = s; bd = hd[s]; bi = hi[s]; }
        out_idx[(size_t)q * k + r] = bi;
        out_d2[(size_t)q * k + r] = bd;
        if (bs < 0) continue;
#pragma unroll
        for (int s = 0; s < KNN_MAX_SPLITS; ++s)
            if (s == bs) {
                ++head[s];
                hi[s] = head[s] < k ? pi[(size_t)s * k + head[s]] : -1;
                hd[s] = head[s] < k ? pd[(size_t)s * k + head[s]] : CUDART_INF;
            }
    }
}

// ----------------------------------------------------------------------------
// Self k-NN with chunk culling (queries == points, the case of SOR, normals, GICP covariances and FPFH).
// The cloud is put in Hilbert order (sort.cuh); every 32 consecutive points form a chunk with a bounding
// sphere.  One warp owns the 32 queries of one chunk, one query per lane, k-best lists in shared memory as
// in knn_f64_kernel.  It scans its own chunk and its neighbours along the curve until every list is full,
// then walks outwards over the remaining chunks and skips every chunk whose sphere lies farther from the
// patch than the worst kept distance of any lane (and, per lane, farther than that lane's own worst).  Skips
// are strict with slack, so a skipped chunk holds no point that could enter a list -- not even on an exact
// tie -- and the lists equal the brute-force result under the (d2, original index) order.
// ----------------------------------------------------------------------------
// warps per block: 4 for k <= 128, 2 for k <= 256, 1 up to ORPCD_MAX_KNN (the lists take k * 384 bytes per warp)
__host__ __device__ constexpr int kc_warps_for(int k) { return k <= 128 ? 4 : (k <= 256 ? 2 : 1); }

// sorted copy of the cloud + chunk spheres.  grid (chunks, batch), 32 threads.
__global__ void __launch_bounds__(32)
knn_cull_prep_kernel(const double *__restrict__ points, int n, const int *__restrict__ perm, int n_pad,
                     double *__restrict__ sorted, double4 *__restrict__ spheres)
{
    const int b = blockIdx.y, ch = blockIdx.x, lane = threadIdx.x;
    const int pos = ch * 32 + lane;
    const int i = perm[(size_t)b * n_pad + pos];
    double x = 0, y = 0, z = 0;
    if (i >= 0) {
        const double *p = points + ((size_t)b * n + i) * 3;
        x = p[0]; y = p[1]; z = p[2];
    }
    double *o = sorted + ((size_t)b * n_pad + pos) * 3;
    o[0] = x; o[1] = y; o[2] = z;
    double lo[3] = {i >= 0 ? x : CUDART_INF, i >= 0 ? y : CUDART_INF, i >= 0 ? z : CUDART_INF};
    double hi[3] = {i >= 0 ? x : -CUDART_INF, i >= 0 ? y : -CUDART_INF, i >= 0 ? z : -CUDART_INF};
    for (int k = 0; k < 3; ++k)
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o2));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o2));
        }
    const double cx = 0.5 * (lo[0] + hi[0]), cy = 0.5 * (lo[1] + hi[1]), cz = 0.5 * (lo[2] + hi[2]);
    double r = i >= 0 ? sqrt((x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz)) : 0.0;
    for (int o2 = 16; o2 > 0; o2 >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o2));
    if (lane == 0) spheres[(size_t)b * (n_pad / 32) + ch] = make_double4(cx, cy, cz, r * (1.0 + 1e-12) + 1e-300);
}

template <int KC_WARPS>
__global__ void __launch_bounds__(KC_WARPS * 32)
knn_cull_kernel(const double *__restrict__ sorted, const double4 *__restrict__ spheres, const int *__restrict__ perm,
                int n, int n_pad, int k, int *__restrict__ out_idx, double *__restrict__ out_d2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, n_chunks = n_pad / 32;
    const int c0 = blockIdx.x * KC_WARPS + warp;   // this warp's chunk = its 32 queries
    if (c0 >= n_chunks) return;
    double *ld = reinterpret_cast<double *>(smem_raw) + (size_t)warp * k * 32;                        // [k][32]
    int *li = reinterpret_cast<int *>(smem_raw + (size_t)KC_WARPS * k * 32 * sizeof(double)) + (size_t)warp * k * 32;
    const double *pts = sorted + (size_t)b * n_pad * 3;
    const int *pm = perm + (size_t)b * n_pad;
    const double4 *sph = spheres + (size_t)b * n_chunks;
    const int qorig = pm[c0 * 32 + lane];
    const bool active = qorig >= 0;
    const double qx = pts[(size_t)(c0 * 32 + lane) * 3], qy = pts[(size_t)(c0 * 32 + lane) * 3 + 1],
                 qz = pts[(size_t)(c0 * 32 + lane) * 3 + 2];
    int cnt = 0, wpos = 0, worst_i = 0x7fffffff;
    double worst = CUDART_INF;

    // all 32 points of chunk ch against this lane's query (broadcast loads)
    auto scan = [&](int ch) {
        const double *cp = pts + (size_t)ch * 96;
        const int *ci = pm + ch * 32;
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
            const int id = ci[j];
            const double d = sqdist3_rn(qx, qy, qz, cp[j * 3], cp[j * 3 + 1], cp[j * 3 + 2]);
            if (id < 0 || !active) continue;
            if (cnt < k) {
                ld[cnt * 32 + lane] = d;
                li[cnt * 32 + lane] = id;
                ++cnt;
                if (cnt < k) continue;
            } else if (d < worst || (d == worst && id < worst_i)) {
                ld[wpos * 32 + lane] = d;
                li[wpos * 32 + lane] = id;
            } else {
                continue;
            }
            // rescan for the lexicographically largest (d2, idx) member
            double wd = ld[lane];
            int wi = li[lane], wp = 0;
            for (int s = 1; s < k; ++s) {
                const double sd = ld[s * 32 + lane];
                const int si = li[s * 32 + lane];
                if (sd > wd || (sd == wd && si > wi)) { wd = sd; wi = si; wp = s; }
            }
            worst = wd; worst_i = wi; wpos = wp;
        }
    };

    // 1) own chunk and its neighbours along the curve until every list is full (or the cloud is exhausted)
    int lo = c0, hi = c0;   // chunks [lo, hi] are done
    scan(c0);
    while ((hi - lo + 1) * 32 < k + 32 && (lo > 0 || hi < n_chunks - 1)) {
        if (hi < n_chunks - 1) scan(++hi);
        if (lo > 0) scan(--lo);
    }
    // 2) outwards, 32 chunks per step on each side, culled
    const double4 ps = sph[c0];
    auto side = [&](int first, int count, int step) {   // chunks first, first+step, ... (count of them)
        double wmax = active && cnt >= k ? worst : (active ? CUDART_INF : 0.0);
        for (int o = 16; o > 0; o >>= 1) wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        bool need = false;
        double4 sc = make_double4(0, 0, 0, 0);
        if (lane < count) {
            sc = sph[first + lane * step];
            const double dx = sc.x - ps.x, dy = sc.y - ps.y, dz = sc.z - ps.z;
            const double lb = sqrt(dx * dx + dy * dy + dz * dz) - sc.w - ps.w;
            need = !(lb > 0.0 && lb * lb * (1.0 - 1e-9) > wmax);
        }
        unsigned bal = __ballot_sync(0xffffffffu, need);
        while (bal) {
            const int l = __ffs(bal) - 1;
            bal &= bal - 1;
            const double sx = __shfl_sync(0xffffffffu, sc.x, l), sy = __shfl_sync(0xffffffffu, sc.y, l);
            const double sz = __shfl_sync(0xffffffffu, sc.z, l), sr = __shfl_sync(0xffffffffu, sc.w, l);
            const double dx = sx - qx, dy = sy - qy, dz = sz - qz;
            const double lb = sqrt(dx * dx + dy * dy + dz * dz) - sr;
            const bool mine = active && !(cnt >= k && lb > 0.0 && lb * lb * (1.0 - 1e-9) > worst);
            if (__any_sync(0xffffffffu, mine)) scan(first + l * step);
        }
    };
    while (lo > 0 || hi < n_chunks - 1) {
        if (hi < n_chunks - 1) {
            const int cntr = min(32, n_chunks - 1 - hi);
            side(hi + 1, cntr, 1);
            hi += cntr;
        }
        if (lo > 0) {
            const int cntl = min(32, lo);
            side(lo - 1, cntl, -1);
            lo -= cntl;
        }
    }
    if (!active) return;
    // sort the kept set by (d2, idx)
    for (int a = 1; a < cnt; ++a) {
        const double d = ld[a * 32 + lane];
        const int id = li[a * 32 + lane];
        int pos = a;
        while (pos > 0) {
            const double pd = ld[(pos - 1) * 32 + lane];
            const int pj = li[(pos - 1) * 32 + lane];
            if (!(d < pd || (d == pd && id < pj))) break;
            ld[pos * 32 + lane] = pd;
            li[pos * 32 + lane] = pj;
            --pos;
        }
        ld[pos * 32 + lane] = d;
        li[pos * 32 + lane] = id;
    }
    int *oi = out_idx + ((size_t)b * n + qorig) * k;
    double *od = out_d2 + ((size_t)b * n + qorig) * k;
    for (int s = 0; s < k; ++s) {
        oi[s] = s < cnt ? li[s * 32 + lane] : -1;
        od[s] = s < cnt ? ld[s * 32 + lane] : CUDART_INF;
    }
}

// ----------------------------------------------------------------------------
// Self k-NN, one WARP per query (k <= 64): the form that fills the GPU at 5000 points.  knn_cull_kernel above gives a
// warp 32 queries and a serial insertion per lane -- 157 warps per 5000-point cloud, 7.7 % achieved occupancy, more
// than half of its instructions in the rescans for the worst list member.  Here the 32 lanes score the 32 points of a
// chunk at once against ONE query; points that beat the current bound tau = (d2, index) of the k-th best are appended
// to a buffer in shared memory with a ballot / popc prefix; when the buffer cannot take another chunk it is sorted
// (bitonic, one compare-exchange per lane and step) and cut back to its k best, which tightens tau.  Chunks are
// visited own chunk first, then along the Hilbert curve until k points are in hand, then the rest of the cloud under
// two-level bounding-sphere culling against tau (strict tests with slack: a skipped chunk holds no point that could
// enter the list, not even on an exact tie).  Same (d2, original index) order as every other k-NN here: bit-exact.
// ----------------------------------------------------------------------------
constexpr int WQ_WARPS = 8;

// one sphere around each group of 32 chunk spheres.  grid (n_super, batch), 32 threads.
__global__ void __launch_bounds__(32)
knn_super_sphere_kernel(const double4 *__restrict__ spheres, int n_chunks, int n_super, double4 *__restrict__ super)
{
    const int b = blockIdx.y, g = blockIdx.x, lane = threadIdx.x;
    const int ch = g * 32 + lane;
    const bool valid = ch < n_chunks;
    const double4 t = valid ? spheres[(size_t)b * n_chunks + ch] : make_double4(0, 0, 0, 0);
    double lo[3] = {valid ? t.x - t.w : CUDART_INF, valid ? t.y - t.w : CUDART_INF, valid ? t.z - t.w : CUDART_INF};
    double hi[3] = {valid ? t.x + t.w : -CUDART_INF, valid ? t.y + t.w : -CUDART_INF, valid ? t.z + t.w : -CUDART_INF};
    for (int k = 0; k < 3; ++k)
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    const double cx = 0.5 * (lo[0] + hi[0]), cy = 0.5 * (lo[1] + hi[1]), cz = 0.5 * (lo[2] + hi[2]);
    double r = valid ? sqrt((t.x - cx) * (t.x - cx) + (t.y - cy) * (t.y - cy) + (t.z - cz) * (t.z - cz)) + t.w : 0.0;
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    if (lane == 0) super[(size_t)b * n_super + g] = make_double4(cx, cy, cz, r * (1.0 + 1e-12) + 1e-300);
}

// C = buffer capacity per query (64 for k <= 32, 128 for k <= 64)
template <int C>
__global__ void __launch_bounds__(WQ_WARPS * 32)
knn_wq_kernel(const double *__restrict__ sorted, const double4 *__restrict__ spheres, const double4 *__restrict__ super,
              const int *__restrict__ perm, int n, int n_pad, int n_super, int k, int *__restrict__ out_idx,
              double *__restrict__ out_d2)
{
    __shared__ double bd_s[WQ_WARPS][C];
    __shared__ int bi_s[WQ_WARPS][C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, n_chunks = n_pad / 32;
    const int qpos = blockIdx.x * WQ_WARPS + warp;         // the query's position in Hilbert order
    if (qpos >= n_pad) return;
    const int *pm = perm + (size_t)b * n_pad;
    const int qorig = pm[qpos];
    if (qorig < 0) return;                                  // padding
    const double *pts = sorted + (size_t)b * n_pad * 3;
    const double4 *sph = spheres + (size_t)b * n_chunks;
    const double4 *ssph = super + (size_t)b * n_super;
    double *bd = bd_s[warp];
    int *bi = bi_s[warp];
    const double qx = pts[(size_t)qpos * 3], qy = pts[(size_t)qpos * 3 + 1], qz = pts[(size_t)qpos * 3 + 2];
    int cnt = 0;
    double tau_d = CUDART_INF;      // the k-th best so far (inf until k points are in hand)
    int tau_i = 0x7fffffff;
    const unsigned lt_mask = (1u << lane) - 1u;

    // buffer -> its k best under (d2, index), sorted; tau tightened
    auto compact = [&]() {
        for (int e = cnt + lane; e < C; e += 32) { bd[e] = CUDART_INF; bi[e] = 0x7fffffff; }
        __syncwarp();
        for (int kk = 2; kk <= C; kk <<= 1)
            for (int j = kk >> 1; j > 0; j >>= 1) {
#pragma unroll
                for (int t = lane; t < C / 2; t += 32) {
                    const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;   // partners at distance j
                    const bool up = (lo & kk) == 0;
                    const double dl = bd[lo], dh = bd[hi];
                    const int il = bi[lo], ih = bi[hi];
                    const bool gt = dl > dh || (dl == dh && il > ih);
                    if (gt == up) { bd[lo] = dh; bd[hi] = dl; bi[lo] = ih; bi[hi] = il; }
                }
                __syncwarp();
            }
        cnt = min(cnt, k);
        if (cnt >= k) { tau_d = bd[k - 1]; tau_i = bi[k - 1]; }
    };
    // the 32 points of chunk ch against the query; the ones that beat tau are appended
    auto scan = [&](int ch) {
        if (cnt + 32 > C) compact();
        const int id = pm[ch * 32 + lane];
        const double *cp = pts + ((size_t)ch * 32 + lane) * 3;
        const double d = sqdist3_rn(qx, qy, qz, cp[0], cp[1], cp[2]);
        const bool keep = id >= 0 && (d < tau_d || (d == tau_d && id < tau_i));
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = cnt + __popc(bal & lt_mask);
            bd[pos] = d;
            bi[pos] = id;
        }
        cnt += __popc(bal);
        __syncwarp();
    };

    // 1) own chunk and its neighbours along the curve until k points are in hand (or the cloud is exhausted)
    const int c0 = qpos >> 5;
    int lo = c0, hi = c0;
    scan(c0);
    while ((hi - lo + 1) * 32 < k + 32 && (lo > 0 || hi < n_chunks - 1)) {
        if (hi < n_chunks - 1) scan(++hi);
        if (lo > 0) scan(--lo);
    }
    compact();
    // 2) every other chunk, culled on two levels against the ball of the k-th best
    auto far = [&](const double4 &sc) {      // true: the sphere holds no point that could enter the list
        const double dx = sc.x - qx, dy = sc.y - qy, dz = sc.z - qz;
        const double lb = sqrt(dx * dx + dy * dy + dz * dz) - sc.w;
        return cnt >= k && lb > 0.0 && lb * lb * (1.0 - 1e-9) > tau_d;
    };
    for (int g0 = 0; g0 < n_super; g0 += 32) {
        bool gneed = false;
        if (g0 + lane < n_super) gneed = !far(ssph[g0 + lane]);
        unsigned gbal = __ballot_sync(0xffffffffu, gneed);
        while (gbal) {
            const int gl = __ffs(gbal) - 1;
            gbal &= gbal - 1;
            const int ch = (g0 + gl) * 32 + lane;
            bool need = false;
            if (ch < n_chunks && (ch < lo || ch > hi)) need = !far(sph[ch]);
            unsigned bal = __ballot_sync(0xffffffffu, need);
            while (bal) {
                const int l = __ffs(bal) - 1;
                bal &= bal - 1;
                const int cc = (g0 + gl) * 32 + l;
                // tau may have tightened since the lane-parallel test: one more look before the 32 loads
                const double4 sc = sph[cc];
                if (!far(sc)) scan(cc);
            }
        }
    }
    compact();
    int *oi = out_idx + ((size_t)b * n + qorig) * k;
    double *od = out_d2 + ((size_t)b * n + qorig) * k;
    for (int s2 = lane; s2 < k; s2 += 32) {
        oi[s2] = s2 < cnt ? bi[s2] : -1;
        od[s2] = s2 < cnt ? bd[s2] : CUDART_INF;
    }
}

struct CullLayout {
    size_t off_perm, off_sorted, off_sph, off_super, off_keys, off_bbox, total;
    int n_pad, n_super;
};
CullLayout knn_cull_layout(int n, int batch)
{
    CullLayout L;
    L.n_pad = (n + 31) / 32 * 32;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += orpcd_align_up(bytes, 256); return r; };
    L.off_perm = take((size_t)batch * L.n_pad * sizeof(int));
    L.off_sorted = take((size_t)batch * L.n_pad * 3 * sizeof(double));
    L.off_sph = take((size_t)batch * (L.n_pad / 32) * sizeof(double4));
    L.n_super = (L.n_pad / 32 + 31) / 32;
    L.off_super = take((size_t)batch * L.n_super * sizeof(double4));
    L.off_keys = take((size_t)batch * sort_pad(n) * sizeof(unsigned long long));
    L.off_bbox = take((size_t)batch * 6 * sizeof(double));
    L.total = o;
    return L;
}
constexpr int KNN_CULL_MIN_N = 2048;   // below this the brute-force kernel is as fast

size_t knn_smem_bytes(int k)
{
    return (size_t)3 * KNN_TILE * sizeof(double) + (size_t)k * knn_block_for(k) * (sizeof(double) + sizeof(int));
}

template <int B>
int launch_knn_f64(dim3 grid, size_t smem, cudaStream_t st, const double *points, int n, long long points_stride,
                   const double *queries, int m, long long queries_stride, int k, int split_len, int *out_idx,
                   double *out_d2)
{
    if (cudaFuncSetAttribute(knn_f64_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return ORPCD_ERR_CUDA;
    knn_f64_kernel<B><<<grid, B, smem, st>>>(points, n, points_stride, queries, m, queries_stride, k, split_len, out_idx,
                                            out_d2);
    return cudaGetLastError() == cudaSuccess ? ORPCD_OK : ORPCD_ERR_CUDA;
}

int launch_knn_f64_any(int gy, int gz, size_t smem, cudaStream_t st, const double *points, int n,
                       long long points_stride, const double *queries, int m, long long queries_stride, int k,
                       int split_len, int *out_idx, double *out_d2)
{
    const int B = knn_block_for(k);
    const dim3 grid((m + B - 1) / B, gy, gz);
    if (B == 128) return launch_knn_f64<128>(grid, smem, st, points, n, points_stride, queries, m, queries_stride, k, split_len, out_idx, out_d2);
    if (B == 64) return launch_knn_f64<64>(grid, smem, st, points, n, points_stride, queries, m, queries_stride, k, split_len, out_idx, out_d2);
    return launch_knn_f64<32>(grid, smem, st, points, n, points_stride, queries, m, queries_stride, k, split_len, out_idx, out_d2);
}

template <int W>
int launch_knn_cull(int n_chunks, int batch, cudaStream_t st, const double *sorted, const double4 *sph, const int *perm,
                    int n, int n_pad, int k, int *out_idx, double *out_d2)
{
    const size_t sm = (size_t)W * k * 32 * (sizeof(double) + sizeof(int));
    if (cudaFuncSetAttribute(knn_cull_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
        return ORPCD_ERR_CUDA;
    knn_cull_kernel<W><<<dim3((n_chunks + W - 1) / W, batch), W * 32, sm, st>>>(sorted, sph, perm, n, n_pad, k, out_idx,
                                                                              out_d2);
    return cudaGetLastError() == cudaSuccess ? ORPCD_OK : ORPCD_ERR_CUDA;
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
constexpr int RR_STEP = 8;     // candidates gathered per step

__global__ void __launch_bounds__(RR_BLOCK)
knn_rerank_kernel(const int *__restrict__ cand, const double *__restrict__ bound, int kc,
                  const double *__restrict__ points, int n, long long stride,
                  const double *__restrict__ factor, int k, int *__restrict__ out_idx,
                  double *__restrict__ out_d2, int *__restrict__ worklist, int *__restrict__ wcount,
                  const int *__restrict__ order, const double *__restrict__ cand_d2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *ld = reinterpret_cast<double *>(smem_raw);                  // [k][RR_BLOCK]
    int *li = reinterpret_cast<int *>(ld + (size_t)k * RR_BLOCK);       // [k][RR_BLOCK]
    const int b = blockIdx.y, tid = threadIdx.x;
    const int t = blockIdx.x * RR_BLOCK + tid;
    if (t >= n) return;
    // processing order: with the base cloud's spatial (Hilbert) order the threads of a block handle neighbouring
    // points, whose candidate lists overlap -- most of the 48 random 24-byte gathers per point then hit L1
    const int i = order ? order[t] : t;
    const double *pts = points + (size_t)b * stride;
    const double qx = pts[(size_t)i * 3], qy = pts[(size_t)i * 3 + 1], qz = pts[(size_t)i * 3 + 2];
    const int *row = cand + (size_t)i * kc;
    const double *row_d2 = cand_d2 ? cand_d2 + (size_t)i * kc : nullptr;
    const double flo = factor[b];
    int cnt = 0;
    // RR_STEP candidates per step: their indices and coordinates are fetched together (the gathers are what this
    // kernel waits for), then inserted one after the other in list order
    for (int s0 = 0; s0 < kc; s0 += RR_STEP) {
        // the candidates are in ascending order of their BASE distance and a moved distance is at least factor x the
        // base one: once k are in hand and the next candidate's bound already exceeds the k-th moved distance, no later
        // candidate can enter the list (strictly: a tie goes on scanning).  Rigid copies stop after ~k + 8 of the 48.
        if (row_d2 && cnt == k && flo * row_d2[s0] > ld[(size_t)(k - 1) * RR_BLOCK + tid]) break;
        int jj[RR_STEP];
        double cx[RR_STEP], cy[RR_STEP], cz[RR_STEP];
#pragma unroll
        for (int r = 0; r < RR_STEP; ++r) jj[r] = s0 + r < kc ? row[s0 + r] : -1;
#pragma unroll
        for (int r = 0; r < RR_STEP; ++r) {
            const int jr = jj[r] < 0 ? 0 : jj[r];
            cx[r] = pts[(size_t)jr * 3]; cy[r] = pts[(size_t)jr * 3 + 1]; cz[r] = pts[(size_t)jr * 3 + 2];
        }
        bool stop = false;
#pragma unroll
        for (int r = 0; r < RR_STEP; ++r) {
        const int j = jj[r];
        if (j < 0) stop = true;
        if (stop) continue;
        const double d = sqdist3_rn(qx, qy, qz, cx[r], cy[r], cz[r]);
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
        if (stop) break;
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

// Splits of the point range per query: enough blocks to fill the GPU when there are few queries
// (one 5000-point cloud is 157 warps for 148 SMs), at least 512 points per split.
int orpcd_knn_splits(int n, int m, int batch)
{
    const long long warps = ((long long)m * batch + 31) / 32;
    long long s = (148LL * 24 + warps - 1) / (warps > 0 ? warps : 1);
    if (s > KNN_MAX_SPLITS) s = KNN_MAX_SPLITS;
    if (s > n / 512) s = n / 512;
    return s < 1 ? 1 : (int)s;
}

size_t orpcd_knn_scratch_bytes(int n, int m, int batch, int k)
{
    const int s = orpcd_knn_splits(n, m, batch);
    size_t split = 0;
    if (s > 1) {
        const size_t e = (size_t)m * batch * s * k;
        split = orpcd_align_up(e * sizeof(double), 256) + orpcd_align_up(e * sizeof(int), 256);
    }
    const size_t cull = (n == m && n >= KNN_CULL_MIN_N) ? knn_cull_layout(n, batch).total : 0;
    return split > cull ? split : cull;
}

// d_scratch (orpcd_knn_scratch_bytes) may be NULL: the search then runs unsplit.
int orpcd_launch_knn(const double *d_points, int n, long long points_stride, const double *d_queries,
                     int m, long long queries_stride, int batch, int k, int *d_idx, double *d_d2,
                     void *d_scratch, size_t scratch_bytes, cudaStream_t st)
{
    if (!d_points || !d_queries || !d_idx || !d_d2 || n < 0 || m < 0 || batch < 0 || k < 1 ||
        k > ORPCD_MAX_KNN)
        return ORPCD_ERR_INVALID;
    if (m == 0 || batch == 0) return ORPCD_OK;
    const size_t smem = knn_smem_bytes(k);
    const bool have_scratch = d_scratch && scratch_bytes >= orpcd_knn_scratch_bytes(n, m, batch, k) &&
                              !((uintptr_t)d_scratch & 255);
    if (have_scratch && d_queries == d_points && m == n && queries_stride == points_stride && n >= KNN_CULL_MIN_N &&
        (batch == 1 || points_stride == (long long)n * 3)) {
        // self query: Hilbert order, chunk spheres, culled scan
        const CullLayout L = knn_cull_layout(n, batch);
        unsigned char *w = (unsigned char *)d_scratch;
        int *perm = (int *)(w + L.off_perm);
        double *sorted = (double *)(w + L.off_sorted);
        double4 *sph = (double4 *)(w + L.off_sph);
        if (spatial_perm(d_points, n, batch, L.n_pad, (double *)(w + L.off_bbox), (unsigned long long *)(w + L.off_keys),
                         perm, st) != 0)
            return ORPCD_ERR_CUDA;
        const int n_chunks = L.n_pad / 32;
        knn_cull_prep_kernel<<<dim3(n_chunks, batch), 32, 0, st>>>(d_points, n, perm, L.n_pad, sorted, sph);
        if (k <= 64) {      // one warp per query
            double4 *ssph = (double4 *)(w + L.off_super);
            knn_super_sphere_kernel<<<dim3(L.n_super, batch), 32, 0, st>>>(sph, n_chunks, L.n_super, ssph);
            const dim3 grid((L.n_pad + WQ_WARPS - 1) / WQ_WARPS, batch);
            if (k <= 32)
                knn_wq_kernel<64><<<grid, WQ_WARPS * 32, 0, st>>>(sorted, sph, ssph, perm, n, L.n_pad, L.n_super, k, d_idx, d_d2);
            else
                knn_wq_kernel<128><<<grid, WQ_WARPS * 32, 0, st>>>(sorted, sph, ssph, perm, n, L.n_pad, L.n_super, k, d_idx, d_d2);
            ORPCD_CHECK_LAUNCH();
            return ORPCD_OK;
        }
        const int W = kc_warps_for(k);
        return W == 4 ? launch_knn_cull<4>(n_chunks, batch, st, sorted, sph, perm, n, L.n_pad, k, d_idx, d_d2)
                      : (W == 2 ? launch_knn_cull<2>(n_chunks, batch, st, sorted, sph, perm, n, L.n_pad, k, d_idx, d_d2)
                                : launch_knn_cull<1>(n_chunks, batch, st, sorted, sph, perm, n, L.n_pad, k, d_idx, d_d2));
    }
    int splits = orpcd_knn_splits(n, m, batch);
    if (!d_scratch || scratch_bytes < orpcd_knn_scratch_bytes(n, m, batch, k) || ((uintptr_t)d_scratch & 255)) splits = 1;
    if (splits <= 1)
        return launch_knn_f64_any(batch, 1, smem, st, d_points, n, points_stride, d_queries, m, queries_stride, k,
                                  n > 0 ? n : 1, d_idx, d_d2);
    const size_t e = (size_t)m * batch * splits * k;
    double *pd = (double *)d_scratch;
    int *pi = (int *)((unsigned char *)d_scratch + orpcd_align_up(e * sizeof(double), 256));
    if (launch_knn_f64_any(batch, splits, smem, st, d_points, n, points_stride, d_queries, m, queries_stride, k,
                           (n + splits - 1) / splits, pi, pd) != ORPCD_OK)
        return ORPCD_ERR_CUDA;
    const long long nq = (long long)m * batch;
    knn_merge_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(pi, pd, nq, splits, k, d_idx, d_d2);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API size_t orpcd_knn_exact_workspace_bytes(int n, int m, int batch, int k)
{
    if (n <= 0 || m <= 0 || batch <= 0 || k < 1) return 0;
    return orpcd_knn_scratch_bytes(n, m, batch, k);
}

ORPCD_API int orpcd_knn_exact(const double *d_points, int n, long long points_stride,
                              const double *d_queries, int m, long long queries_stride, int batch,
                              int k, int *d_idx, double *d_d2, void *d_workspace, size_t workspace_bytes,
                              orpcd_stream_t stream)
{
    return orpcd_launch_knn(d_points, n, points_stride, d_queries, m, queries_stride, batch, k, d_idx,
                            d_d2, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

ORPCD_API size_t orpcd_sor_workspace_bytes(int n, int nb_neighbors)
{
    if (n <= 0 || nb_neighbors < 1) return 0;
    const size_t nn = (size_t)n * nb_neighbors;
    return orpcd_align_up(nn * sizeof(double), 256) + orpcd_align_up(nn * sizeof(int), 256) +
           orpcd_align_up((size_t)n * sizeof(double), 256) + orpcd_knn_scratch_bytes(n, n, 1, nb_neighbors);
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
    w += orpcd_align_up((size_t)n * sizeof(double), 256);
    int rc = orpcd_launch_knn(d_points, n, 0, d_points, n, 0, 1, nb_neighbors, idx, d2, w,
                              orpcd_knn_scratch_bytes(n, n, 1, nb_neighbors), st);
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

ORPCD_API size_t orpcd_spatial_order_workspace_bytes(int n)
{
    if (n <= 0) return 0;
    return orpcd_align_up((size_t)sort_pad(n) * sizeof(unsigned long long), 256) + 256;
}

ORPCD_API int orpcd_spatial_order(const double *d_points, int n, void *d_workspace, size_t workspace_bytes,
                                  int *d_order, orpcd_stream_t stream)
{
    if (!d_points || !d_order || n <= 0) return ORPCD_ERR_INVALID;
    if (!d_workspace || workspace_bytes < orpcd_spatial_order_workspace_bytes(n) || ((uintptr_t)d_workspace & 255))
        return ORPCD_ERR_WORKSPACE;
    unsigned char *w = (unsigned char *)d_workspace;
    double *bbox = (double *)w;                                     // 6 doubles
    unsigned long long *keys = (unsigned long long *)(w + 256);
    return spatial_perm(d_points, n, 1, n, bbox, keys, d_order, (cudaStream_t)stream) == 0 ? ORPCD_OK : ORPCD_ERR_CUDA;
}

ORPCD_API int orpcd_knn_rerank(const int *d_cand_idx, const double *d_cand_bound, int kc,
                               const double *d_points, int n, long long points_stride, int batch,
                               const double *d_factor, int k, int *d_idx, double *d_d2, void *d_workspace,
                               size_t workspace_bytes, const int *d_order, const double *d_cand_d2,
                               orpcd_stream_t stream)
{
    if (!d_cand_idx || !d_cand_bound || !d_points || !d_factor || !d_idx || !d_d2 || n <= 0 || batch <= 0 ||
        k < 1 || k > 64 || kc < k || (long long)n * batch > 0x7fffffffLL)   /* lists: k * 1536 bytes of shared memory */
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
                                                    k, d_idx, d_d2, worklist, wcount, d_order, d_cand_d2);
    ORPCD_CHECK_LAUNCH();
    knn_rerank_fix_kernel<<<148 * 2, 128, 0, st>>>(d_points, n, points_stride, k, d_idx, d_d2, worklist, wcount);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}
