// Down-sampling blocks of the preprocessing pipeline.
//
// orpcd_farthest_point_sample  replaces the numpy/scipy loop of
//     Preprocessor/Downsamplers/farthestDownsampler.py:33-54 (PointNet++ farthest point sampling:
//     k sequential steps of  dist = min(dist, cdist(point, cloud));  next = argmax(dist)).
// orpcd_voxel_downsample       replaces open3d PointCloud.voxel_down_sample as called from
//     Preprocessor/Downsamplers/voxelDownsampler.py:83,95 (voxel-grid averaging).
//
// Both are bit-exact restatements in float64: scipy's euclidean is sqrt(((dx*dx)+dy*dy)+dz*dz),
// np.argmax returns the first maximum; Open3D averages the points of a voxel in index order.
#include "sort.cuh"

namespace {

// ----------------------------------------------------------------------------
// farthest point sampling: one persistent block, the running distance of each point lives in
// registers (FPS_PER_THREAD points per thread; larger clouds keep them in an L2-resident workspace),
// the cloud is re-read from L2 every step
// ----------------------------------------------------------------------------
constexpr int FPS_THREADS = 1024;
constexpr int FPS_PER_THREAD = 16;   // register-resident distances: clouds up to 16384 points, else L2

template <bool IN_REGS>
__global__ void __launch_bounds__(FPS_THREADS) fps_kernel(const double *__restrict__ pts, int n, int first, int k,
                                                         double *__restrict__ dist_ws, int *__restrict__ out)
{
    __shared__ double s_val[FPS_THREADS / 32];
    __shared__ int s_idx[FPS_THREADS / 32];
    __shared__ int s_cur;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double dist[IN_REGS ? FPS_PER_THREAD : 1];
    if (IN_REGS) {
#pragma unroll
        for (int r = 0; r < FPS_PER_THREAD; ++r) dist[r] = 1e6;
    } else {
        for (int i = tid; i < n; i += FPS_THREADS) dist_ws[i] = 1e6;
    }
    int cur = first;
    if (tid == 0) out[0] = cur;
    for (int s = 0; s + 1 < k; ++s) {
        const double px = pts[(size_t)cur * 3], py = pts[(size_t)cur * 3 + 1], pz = pts[(size_t)cur * 3 + 2];
        double bv = -1.0;
        int bi = 0x7fffffff;
        if (IN_REGS) {
#pragma unroll
            for (int r = 0; r < FPS_PER_THREAD; ++r) {
                const int i = r * FPS_THREADS + tid;
                if (i < n) {
                    const double d = sqrt(sqdist3_rn(pts[(size_t)i * 3], pts[(size_t)i * 3 + 1], pts[(size_t)i * 3 + 2],
                                                     px, py, pz));
                    const double nd = fmin(d, dist[r]);
                    dist[r] = nd;
                    if (nd > bv) { bv = nd; bi = i; }   // ascending i: the first maximum wins
                }
            }
        } else {
            for (int i = tid; i < n; i += FPS_THREADS) {
                const double d = sqrt(sqdist3_rn(pts[(size_t)i * 3], pts[(size_t)i * 3 + 1], pts[(size_t)i * 3 + 2],
                                                 px, py, pz));
                const double nd = fmin(d, dist_ws[i]);
                dist_ws[i] = nd;
                if (nd > bv) { bv = nd; bi = i; }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_val[wid] = bv; s_idx[wid] = bi; }
        __syncthreads();
        if (wid == 0) {
            bv = s_val[lane];
            bi = s_idx[lane];
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { s_cur = bi; out[s + 1] = bi; }
        }
        __syncthreads();
        cur = s_cur;
    }
}

// ----------------------------------------------------------------------------
// voxel-grid down-sampling
// ----------------------------------------------------------------------------
struct VoxelGrid {
    double vmin[3];      // voxel_min_bound = min_bound - voxel_size / 2
    long long dim[3];    // cells per axis
    int idx_bits;        // low bits of a sort word that hold the point index
    int error;           // 1: the (cell id, point index) pair does not fit one 64-bit word
    int n_voxels;
};

__global__ void voxel_grid_kernel(const double *__restrict__ bbox, int n, double voxel_size, VoxelGrid *g)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int bits = 1;
    while ((1ll << bits) < n) ++bits;
    g->idx_bits = bits;
    double cells = 1.0;
    for (int k = 0; k < 3; ++k) {
        g->vmin[k] = bbox[k] - voxel_size * 0.5;
        const double top = floor((bbox[3 + k] - g->vmin[k]) / voxel_size);
        g->dim[k] = (long long)top + 1;
        cells *= (double)g->dim[k];
    }
    // error 1: Open3D's own condition ("voxel_size is too small": voxel_size * INT_MAX < extent);
    // error 2: a limit of THIS implementation -- the dense linear cell id and the point index share one 64-bit sort
    // word, so dim_x * dim_y * dim_z must stay below 2^(63 - idx_bits) (Open3D hashes the occupied cells and has
    // no such limit; a cloud of extent 100 with a 1e-4 voxel would hit it)
    bool o3d_small = false;
    for (int k = 0; k < 3; ++k) o3d_small = o3d_small || (voxel_size * 2147483647.0 < bbox[3 + k] - bbox[k]);
    g->error = o3d_small ? 1 : (!(cells < ldexp(1.0, 63 - bits)) ? 2 : 0);
    g->n_voxels = 0;
}

// word = (linear cell id << idx_bits) | point index; padding sorts last
__global__ void __launch_bounds__(256) voxel_key_kernel(const double *__restrict__ pts, int n, int pad2, double voxel_size,
                                                       const VoxelGrid *__restrict__ g, unsigned long long *__restrict__ keys)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= pad2) return;
    unsigned long long w = ~0ull;
    if (i < n && !g->error) {
        long long c[3];
        for (int k = 0; k < 3; ++k) c[k] = (long long)floor((pts[(size_t)i * 3 + k] - g->vmin[k]) / voxel_size);
        const unsigned long long cell = ((unsigned long long)c[0] * g->dim[1] + c[1]) * g->dim[2] + c[2];
        w = (cell << g->idx_bits) | (unsigned)i;
    }
    keys[i] = w;
}

// single block: head flags -> output slots (ordered), then one thread per voxel averages its points in
// index order (AccumulatedPoint::AddPoint order) and writes slot-th output point
__global__ void __launch_bounds__(1024) voxel_average_kernel(const double *__restrict__ pts, int n,
                                                            const unsigned long long *__restrict__ keys,
                                                            VoxelGrid *g, int *__restrict__ slot_ws,
                                                            double *__restrict__ out, int *__restrict__ out_index_first)
{
    __shared__ int warp_tot[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (g->error) {
        if (tid == 0) g->n_voxels = -g->error;     // -1: Open3D's "voxel_size is too small"; -2: sort-word limit
        return;
    }
    const int shift = g->idx_bits;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        int head = 0;
        if (i < n) head = (i == 0) || ((keys[i] >> shift) != (keys[i - 1] >> shift));
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        const int within = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            if (w < wid) before += warp_tot[w];
            total += warp_tot[w];
        }
        if (i < n) slot_ws[i] = head ? s_base + before + within : -1;
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    if (tid == 0) g->n_voxels = s_base;
    for (int i = tid; i < n; i += 1024) {
        const int slot = slot_ws[i];
        if (slot < 0) continue;
        const unsigned long long cell = keys[i] >> shift;
        const unsigned long long mask = (1ull << shift) - 1ull;
        double sx = 0.0, sy = 0.0, sz = 0.0;
        int cnt = 0;
        for (int j = i; j < n && (keys[j] >> shift) == cell; ++j) {
            const int p = (int)(keys[j] & mask);
            sx += pts[(size_t)p * 3]; sy += pts[(size_t)p * 3 + 1]; sz += pts[(size_t)p * 3 + 2];
            ++cnt;
        }
        out[(size_t)slot * 3] = sx / (double)cnt;
        out[(size_t)slot * 3 + 1] = sy / (double)cnt;
        out[(size_t)slot * 3 + 2] = sz / (double)cnt;
        if (out_index_first) out_index_first[slot] = (int)(keys[i] & mask);
    }
}

}  // namespace

ORPCD_API int orpcd_farthest_point_sample(const double *d_points, int n, int first_index, int sample_size,
                                          double *d_dist_workspace, int *d_out_indices, orpcd_stream_t stream)
{
    if (!d_points || !d_out_indices || n <= 0 || sample_size <= 0 || first_index < 0 || first_index >= n)
        return ORPCD_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= FPS_THREADS * FPS_PER_THREAD) {
        fps_kernel<true><<<1, FPS_THREADS, 0, st>>>(d_points, n, first_index, sample_size, nullptr, d_out_indices);
    } else {
        if (!d_dist_workspace) return ORPCD_ERR_WORKSPACE;
        fps_kernel<false><<<1, FPS_THREADS, 0, st>>>(d_points, n, first_index, sample_size, d_dist_workspace,
                                                     d_out_indices);
    }
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API size_t orpcd_voxel_downsample_workspace_bytes(int n)
{
    if (n <= 0) return 0;
    return orpcd_align_up(sizeof(VoxelGrid), 256) + orpcd_align_up(6 * sizeof(double), 256) +
           orpcd_align_up((size_t)sort_pad(n) * sizeof(unsigned long long), 256) +
           orpcd_align_up((size_t)n * sizeof(int), 256);
}

ORPCD_API int orpcd_voxel_downsample(const double *d_points, int n, double voxel_size, void *d_workspace,
                                     size_t workspace_bytes, double *d_out_points, int *d_out_first_index,
                                     int *d_out_count, orpcd_stream_t stream)
{
    if (!d_points || !d_out_points || !d_out_count || n <= 0 || !(voxel_size > 0)) return ORPCD_ERR_INVALID;
    if (!d_workspace || workspace_bytes < orpcd_voxel_downsample_workspace_bytes(n) || ((uintptr_t)d_workspace & 255))
        return ORPCD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *w = (unsigned char *)d_workspace;
    VoxelGrid *g = (VoxelGrid *)w;
    w += orpcd_align_up(sizeof(VoxelGrid), 256);
    double *bbox = (double *)w;
    w += orpcd_align_up(6 * sizeof(double), 256);
    unsigned long long *keys = (unsigned long long *)w;
    const int pad2 = sort_pad(n);
    w += orpcd_align_up((size_t)pad2 * sizeof(unsigned long long), 256);
    int *slots = (int *)w;
    bbox_kernel<<<1, 256, 0, st>>>(d_points, n, bbox);
    voxel_grid_kernel<<<1, 32, 0, st>>>(bbox, n, voxel_size, g);
    voxel_key_kernel<<<pad2 / 256, 256, 0, st>>>(d_points, n, pad2, voxel_size, g, keys);
    if (bitonic_sort_u64(keys, pad2, 1, st) != 0) return ORPCD_ERR_CUDA;
    voxel_average_kernel<<<1, 1024, 0, st>>>(d_points, n, keys, g, slots, d_out_points, d_out_first_index);
    ORPCD_CHECK_LAUNCH();
    if (cudaMemcpyAsync(d_out_count, &g->n_voxels, sizeof(int), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return ORPCD_ERR_CUDA;
    return ORPCD_OK;
}
