// Bounding boxes, Hilbert keys and a bitonic sort on 64-bit key|index words, shared by the
// registration kernels (spatial ordering of clouds, gicp.cu) and the voxel-grid down-sampler
// (downsample.cu).  Header-only: every translation unit gets its own copy in its anonymous namespace.
#pragma once
#include "common.cuh"

namespace {

// ----------------------------------------------------------------------------
// spatial (Hilbert) ordering of point sets: consecutive indices are then close in space, which is
// what lets a warp of the 1-NN sweep skip whole 32-target chunks
// ----------------------------------------------------------------------------
constexpr int SORT_TILE = 1024;

__global__ void __launch_bounds__(256) bbox_kernel(const double *__restrict__ pts, int n, double *__restrict__ bbox)
{
    __shared__ double red[6][8];
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double *p = pts + (size_t)s * n * 3;
    double mn[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, mx[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
    for (int i = tid; i < n; i += 256)
        for (int k = 0; k < 3; ++k) {
            const double v = p[(size_t)i * 3 + k];
            mn[k] = fmin(mn[k], v);
            mx[k] = fmax(mx[k], v);
        }
    for (int k = 0; k < 3; ++k)
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    if (lane == 0)
        for (int k = 0; k < 3; ++k) { red[k][wid] = mn[k]; red[3 + k][wid] = mx[k]; }
    __syncthreads();
    if (tid < 3) {
        double a = CUDART_INF, b = -CUDART_INF;
        for (int w = 0; w < 8; ++w) { a = fmin(a, red[tid][w]); b = fmax(b, red[3 + tid][w]); }
        bbox[(size_t)s * 6 + tid] = a;
        bbox[(size_t)s * 6 + 3 + tid] = b;
    }
}

__device__ __forceinline__ unsigned part1by2(unsigned x)
{
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

// key = (30-bit Morton code << 32) | index; padding entries sort last.  grid (pad2/256, sets)
__global__ void __launch_bounds__(256) morton_kernel(const double *__restrict__ pts, int n, int pad2,
                                                    const double *__restrict__ bbox, unsigned long long *__restrict__ keys)
{
    const int s = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    if (i >= pad2) return;
    unsigned long long key = ~0ull;
    if (i < n) {
        const double *b = bbox + (size_t)s * 6;
        const double ext = fmax(fmax(b[3] - b[0], b[4] - b[1]), b[5] - b[2]);
        const double inv = ext > 0 ? 1024.0 / ext : 0.0;
        const double *p = pts + ((size_t)s * n + i) * 3;
        unsigned q[3];
        for (int k = 0; k < 3; ++k) {
            const int v = (int)((p[k] - b[k]) * inv);
            q[k] = (unsigned)(v < 0 ? 0 : (v > 1023 ? 1023 : v));
        }
        // Hilbert index (Skilling, "Programming the Hilbert curve", 2004: axes -> transpose), 10 bits per
        // axis: unlike the Z-order it has no long jumps, so runs of consecutive points stay compact
        {
            const unsigned M = 1u << 9;
            for (unsigned Qb = M; Qb > 1; Qb >>= 1) {
                const unsigned P = Qb - 1;
                for (int k = 0; k < 3; ++k) {
                    if (q[k] & Qb) q[0] ^= P;
                    else { const unsigned t = (q[0] ^ q[k]) & P; q[0] ^= t; q[k] ^= t; }
                }
            }
            for (int k = 1; k < 3; ++k) q[k] ^= q[k - 1];
            unsigned t = 0;
            for (unsigned Qb = M; Qb > 1; Qb >>= 1)
                if (q[2] & Qb) t ^= Qb - 1;
            for (int k = 0; k < 3; ++k) q[k] ^= t;
        }
        // interleave the transposed form: bit b of axis k -> bit 3*b + (2 - k)
        const unsigned code = (part1by2(q[0]) << 2) | (part1by2(q[1]) << 1) | part1by2(q[2]);
        key = ((unsigned long long)code << 32) | (unsigned)i;
    }
    keys[(size_t)s * pad2 + i] = key;
}

// bitonic network on 64-bit keys: FULL sorts each 1024-element tile, otherwise the last ten
// steps (j = 512..1) of merge phase k run in shared memory.  grid (pad2/1024, sets), 512 threads
template <bool FULL>
__global__ void __launch_bounds__(512) bitonic_local_kernel(unsigned long long *__restrict__ keys, int pad2, int k_phase)
{
    __shared__ unsigned long long sh[SORT_TILE];
    unsigned long long *base = keys + (size_t)blockIdx.y * pad2 + (size_t)blockIdx.x * SORT_TILE;
    const int t = threadIdx.x;
    sh[t] = base[t];
    sh[t + 512] = base[t + 512];
    __syncthreads();
    for (int k = FULL ? 2 : k_phase; k <= (FULL ? SORT_TILE : k_phase); k <<= 1) {
        for (int j = (FULL ? k : SORT_TILE) >> 1; j > 0; j >>= 1) {
            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            const int l = i | j;
            const bool up = (((blockIdx.x * SORT_TILE + i) & k) == 0);
            const unsigned long long a = sh[i], b = sh[l];
            if ((a > b) == up) { sh[i] = b; sh[l] = a; }
            __syncthreads();
        }
    }
    base[t] = sh[t];
    base[t + 512] = sh[t + 512];
}

// one global compare-exchange step (j >= 1024).  grid (pad2/512, sets), 256 threads
__global__ void __launch_bounds__(256) bitonic_global_kernel(unsigned long long *__restrict__ keys, int pad2, int j, int k)
{
    unsigned long long *base = keys + (size_t)blockIdx.y * pad2;
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= pad2 / 2) return;
    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const int l = i | j;
    const bool up = ((i & k) == 0);
    const unsigned long long a = base[i], b = base[l];
    if ((a > b) == up) { base[i] = b; base[l] = a; }
}

// grid (ceil(out_n/256), sets): perm[pos] = original index at sorted position pos (-1 for padding)
__global__ void __launch_bounds__(256) perm_kernel(const unsigned long long *__restrict__ keys, int n, int pad2,
                                                  int out_n, int *__restrict__ perm)
{
    const int s = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    if (i >= out_n) return;
    perm[(size_t)s * out_n + i] = i < n ? (int)(unsigned)(keys[(size_t)s * pad2 + i] & 0xffffffffull) : -1;
}

inline int sort_pad(int n)
{
    int p = SORT_TILE;
    while (p < n) p <<= 1;
    return p;
}

// perm [sets][out_n] <- Morton order of pts [sets][n][3]
inline int spatial_perm(const double *pts, int n, int sets, int out_n, double *bbox, unsigned long long *keys, int *perm,
                 cudaStream_t st)
{
    const int pad2 = sort_pad(n);
    bbox_kernel<<<sets, 256, 0, st>>>(pts, n, bbox);
    morton_kernel<<<dim3(pad2 / 256, sets), 256, 0, st>>>(pts, n, pad2, bbox, keys);
    bitonic_local_kernel<true><<<dim3(pad2 / SORT_TILE, sets), 512, 0, st>>>(keys, pad2, 0);
    for (int k = 2 * SORT_TILE; k <= pad2; k <<= 1) {
        for (int j = k >> 1; j >= SORT_TILE; j >>= 1)
            bitonic_global_kernel<<<dim3((pad2 / 2 + 255) / 256, sets), 256, 0, st>>>(keys, pad2, j, k);
        bitonic_local_kernel<false><<<dim3(pad2 / SORT_TILE, sets), 512, 0, st>>>(keys, pad2, k);
    }
    perm_kernel<<<dim3((out_n + 255) / 256, sets), 256, 0, st>>>(keys, n, pad2, out_n, perm);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}


// sort `sets` independent arrays of pad2 (power of two >= SORT_TILE) 64-bit words ascending
inline int bitonic_sort_u64(unsigned long long *keys, int pad2, int sets, cudaStream_t st)
{
    bitonic_local_kernel<true><<<dim3(pad2 / SORT_TILE, sets), 512, 0, st>>>(keys, pad2, 0);
    for (int k = 2 * SORT_TILE; k <= pad2; k <<= 1) {
        for (int j = k >> 1; j >= SORT_TILE; j >>= 1)
            bitonic_global_kernel<<<dim3((pad2 / 2 + 255) / 256, sets), 256, 0, st>>>(keys, pad2, j, k);
        bitonic_local_kernel<false><<<dim3(pad2 / SORT_TILE, sets), 512, 0, st>>>(keys, pad2, k);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace
