// Batched Generalized-ICP for sm_100a: every (multi-start x scale-candidate)
// inner registration of the Aligner's pattern search advances together, one
// launch sequence per Gauss-Newton iteration for all of them.
//
// Replaces open3d.pipelines.registration.registration_generalized_icp as called
// from Optimizer/generalizedICP.py:59-70 (once per start, Aligner/Aligner.py:188-190).
// Algorithm: SURVEY.md Appendix A.4-A.5.
//
// Default path: ONE launch per Gauss-Newton round for all units, gicp_round_kernel (see its header comment):
// every warp owns a patch of 64 Hilbert-consecutive source points and runs the pending rigid update, the
// chunk-culled FP32 1-NN sweep (packed fma.rn.f32x2 on |p|^2 - 2 q.p, FMNMX3 minima per 8-target sub-chunk),
// the FP32 certificate with the float64 re-rank, the exact float64 search for what the certificate leaves open,
// and the Mahalanobis residual / Jacobian sums; the last warp of a unit adds the per-patch partials in patch
// order, tests convergence, solves the 6x6 system (LDLT) and composes the transform.
//
// The same arithmetic as separate kernels (sweep_mode 1..4, kept for the dense brute-force sweeps and A/B runs):
//   update_kernel     applies the pending rigid update to the stored float64 source
//                     points and covariances (or covariance axes) and writes the
//                     centred FP32 copy the sweep reads.
//   nn_sweep_kernel / nn_sweep_cull_kernel
//                     FP32 filter on the FMA pipe: per query (best chunk min, best chunk,
//                     second-best chunk min), over all targets or over the chunks a patch
//                     cannot exclude.
//   nn_resolve_kernel proves with an FP32 error bound that the best chunk holds the
//                     exact nearest neighbour, then a warp-wide ballot over that
//                     chunk finds the single target inside the certified margin
//                     and its float64 distance is taken in Open3D's operation
//                     order; anything ambiguous at FP32 resolution goes to a work
//                     list.
//   nn_fix_kernel     exact float64 search, one warp per listed query.
//   reduce_kernel     residual / Jacobian sums per correspondence, block partials, the
//                     last block of a unit sums them in index order, tests convergence,
//                     solves the 6x6 system and composes the transform.
#include <vector>
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "sort.cuh"

namespace {

// ----------------------------------------------------------------------------
// configuration
// ----------------------------------------------------------------------------
constexpr int NN_THREADS = 128;
constexpr int NN_CH = 32;                        // targets per chunk (re-rank granularity)
constexpr int NN_TILE_CH = 32;                   // chunks per shared-memory tile
constexpr int NN_TILE = NN_CH * NN_TILE_CH;      // 1024 targets = 16 KB per stage
constexpr int RED_THREADS = 128;
constexpr int RED_VALS = 29;                     // 21 JTJ + 6 JTr + sum d2 + count
constexpr int RED_STRIDE = 32;
constexpr int RED_PER_THREAD = 4;                // correspondences accumulated per thread before the shuffles
constexpr float F32_EPS = 5.9604644775390625e-8f;  // 2^-24, unit roundoff
#ifndef ORPCD_HALF_FROM_ROUND
#define ORPCD_HALF_FROM_ROUND 32   // fused rounds: 32-point patches from this round on (see round_body)
#endif
#ifndef ORPCD_SPIN_NS
#define ORPCD_SPIN_NS 500          // back-off of a worker waiting on an empty ring slot (a patch takes tens of microseconds)
#endif
#ifndef ORPCD_PLANE_MMA
#define ORPCD_PLANE_MMA 1          // EST 1: Jacobian sums on the float64 tensor cores (0: per-lane accumulators + butterfly)
#endif

struct UnitState {
    double T[16];       // accumulated transform (row-major 4x4, column convention)
    double U[12];       // pending update [R|t] not yet applied to points / covariances
    double fitness;     // result of the previous Evaluate
    double rmse;
    int iters;
    int done;
    int has_update;
    int n_corr;
    int src_set;
    int tgt_set;
    int status;
    int pad;
};

struct TgtMeta {
    double cx, cy, cz;  // centre subtracted before rounding to float
    double rmax;        // max |p - c| over the set (slightly inflated)
};

struct Ctx {
    int n_units, n_src, n_tgt, nt_pad, n_chunks, splits, chunks_per_split;
    int max_iter, red_blocks, cov_mode, cov_stride, ignore_done;
    double r2, rel_fit, rel_rmse, kappa;
    UnitState *state;
    double *cur_pts;     // [units][n_src][3]
    double *cur_cov;     // [units][n_src][6] covariances (cov_mode 0) or [..][3] axes (cov_mode 1)
    float4 *cur_qf;      // [units][n_src] centred FP32 copy of cur_pts (w unused)
    int *corr;           // [units][n_src]
    double *corr_d2;     // [units][n_src]
    float *part_m1;      // [units][splits][n_src]
    int *part_c1;
    float *part_m2;
    float4 *tgt4;        // [tgt_sets][nt_pad]  (-2x, -2y, -2z, |p|^2), centred
    TgtMeta *tgt_meta;   // [tgt_sets]
    double *red_partial; // [units][red_blocks][RED_STRIDE_MMA] (the unfused reduce_kernel uses the first RED_STRIDE of each row... see its indexing)
    unsigned *red_counter;  // [units][red_groups + 1]: [0] groups closed in the unit's round, [1 + g] arrivals in group g
    int red_groups;         // counter stride - 1 (>= groups of the finest patch size)
    int *worklist;       // [units*n_src]
    int *worklist_count; // [max_iter+2]
    int *done_count;     // [1]
    int *round_active;   // [max_iter+2] units still running in each round
    unsigned long long *visit_count;  // [max_iter+2] (query, target) pairs the sweep evaluated per round
    int *sperm;          // [src_sets][n_src]  spatially sorted position -> original source index
    int *tperm;          // [tgt_sets][nt_pad] spatially sorted position -> original target index (-1 = padding)
    float4 *chunk_sph;   // [tgt_sets][n_chunks] bounding sphere (centre, radius) of each 32-target chunk, centred coords
    float4 *super_sph;   // [tgt_sets][n_super] bounding sphere of each group of 32 chunks (1024 targets)
    float4 *chunk_half;  // [tgt_sets][n_chunks] half extents (hx, hy, hz, 0) of each chunk's bounding box about the sphere's centre
    int n_super;
    unsigned long long *sortbuf;  // Morton keys being sorted
    double *bbox;        // [max(src_sets, tgt_sets)][6]
    float dcap;          // max correspondence distance (inflated): nothing farther is ever needed
    const double *tgt_pts;   // [tgt_sets][n_tgt][3]
    const double *tgt_cov;   // [tgt_sets][n_tgt][cov_stride]
    double *out_T, *out_stats, *trace;
    int *out_status;
    // work queue of the persistent round kernel: ring of (unit * n_patches + patch + 1), 0 = empty slot
    unsigned *q_slots;
    unsigned *q_ctrl;    // [0] head (tickets handed out), [32] tail (items published), [64] all-units-done flag
    unsigned long long *q_stats;   // ORPCD_QUEUE_STATS: [0] start time; [64..127] idle ns per 250 us bucket; [128..191] items per bucket
    unsigned q_mask;     // ring capacity - 1 (capacity = power of two >= n_units * n_patches)
    int n_patches;       // 64-point patches (warps' work items) per unit and round
    int n_patches2;      // 32-point patches per unit and round (rounds >= half_from_round); also the item stride of a unit
    int half_from_round; // from this round on a unit's work items are 32-point patches (see round_body)
};
constexpr int Q_HEAD = 0, Q_TAIL = 32, Q_DONE = 64;

// ----------------------------------------------------------------------------
// small PTX helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pack_f2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// a * {b,b} + {c,c}
__device__ __forceinline__ uint64_t fma2_ss(uint64_t a, float b, float c)
{
    uint64_t d;
    asm("{\n\t.reg .b64 rb, rc;\n\tmov.b64 rb, {%2, %2};\n\tmov.b64 rc, {%3, %3};\n\t"
        "fma.rn.f32x2 %0, %1, rb, rc;\n\t}"
        : "=l"(d) : "l"(a), "f"(b), "f"(c));
    return d;
}
// a * {b,b} + c
__device__ __forceinline__ uint64_t fma2_sv(uint64_t a, float b, uint64_t c)
{
    uint64_t d;
    asm("{\n\t.reg .b64 rb;\n\tmov.b64 rb, {%2, %2};\n\tfma.rn.f32x2 %0, %1, rb, %3;\n\t}"
        : "=l"(d) : "l"(a), "f"(b), "l"(c));
    return d;
}
// packed FP32 pairs: a + {b,b}, a * b, a * b + c
__device__ __forceinline__ uint64_t add2_s(uint64_t a, float b)
{
    uint64_t d;
    asm("{\n\t.reg .b64 rb;\n\tmov.b64 rb, {%2, %2};\n\tadd.rn.f32x2 %0, %1, rb;\n\t}" : "=l"(d) : "l"(a), "f"(b));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
{
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// min / max over the 32 lanes of a (converged) warp; the same value in every lane.  No NaN inputs on these paths.
#ifndef ORPCD_CREDUX
#define ORPCD_CREDUX 1
#endif
__device__ __forceinline__ float warp_min_f32(float v)
{
#if ORPCD_CREDUX
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
#else
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
#endif
}
__device__ __forceinline__ float warp_max_f32(float v)
{
#if ORPCD_CREDUX
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
#else
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
#endif
}
__device__ __forceinline__ float min3f(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------
// set-up kernels
// ----------------------------------------------------------------------------
// Release / acquire without wiping the L1.  `__threadfence()` is `fence.sc.gpu`: SASS MEMBAR.SC.GPU + CCTL.IVALL, i.e.
// every fence INVALIDATES THE SM'S WHOLE L1 -- two per work item with 20 worker warps per SM emptied the L1 (target
// tables, chunk spheres, permutations: the immutable data the sweep lives on) every microsecond.  The queue only
// needs (1) a release before the atomic that hands data on -- `atom.release.gpu` is MEMBAR.ALL.GPU + the atomic, no
// invalidate; `__syncwarp()` before it orders the other lanes' stores ahead of lane 0's release (cumulativity) -- and
// (2) on the consuming side nothing at all: every load of data that another SM wrote during the launch is a strong
// L2 load (`__ldcg`, SASS LDG.STRONG.GPU) whose address depends on the value the consuming atomic returned, and the
// L1 never holds a line of that data.  ORPCD_LIGHT_FENCE=0 restores the fences for A/B runs.
#ifndef ORPCD_LIGHT_FENCE
#define ORPCD_LIGHT_FENCE 1
#endif
__device__ __forceinline__ unsigned atom_add_release(unsigned *p, unsigned v)
{
#if ORPCD_LIGHT_FENCE
    unsigned r;
    asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
    return r;
#else
    return atomicAdd(p, v);
#endif
}
__device__ __forceinline__ void fence_release_side()     // all lanes, before __syncwarp() + atom_add_release by lane 0
{
#if !ORPCD_LIGHT_FENCE
    __threadfence();
#endif
}
__device__ __forceinline__ void fence_acquire_side()     // after the consuming atomic
{
#if ORPCD_LIGHT_FENCE
    asm volatile("" ::: "memory");
#else
    __threadfence();
#endif
}

__global__ void init_units_kernel(Ctx c, const int *__restrict__ unit_src, const int *__restrict__ unit_tgt,
                                  const double *__restrict__ src_pts, const double *__restrict__ src_cov,
                                  const double *__restrict__ init_T)
{
    const int u = blockIdx.y;
    const int ss = unit_src[u];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        UnitState &s = c.state[u];
        for (int k = 0; k < 16; ++k) s.T[k] = init_T ? init_T[(size_t)u * 16 + k] : ((k % 5 == 0) ? 1.0 : 0.0);
        for (int k = 0; k < 12; ++k) s.U[k] = s.T[k];
        s.fitness = 0; s.rmse = 0; s.iters = 0; s.done = 0; s.has_update = init_T ? 1 : 0; s.n_corr = 0;
        s.src_set = ss; s.tgt_set = unit_tgt[u]; s.status = 0; s.pad = 0;
        for (int k = 0; k <= c.red_groups; ++k) c.red_counter[(size_t)u * (c.red_groups + 1) + k] = 0;
        if (u == 0) {
            *c.done_count = 0;
            for (int k = 0; k < c.max_iter + 2; ++k) { c.worklist_count[k] = 0; c.round_active[k] = 0; c.visit_count[k] = 0ull; }
        }
    }
    const int cs = src_cov ? c.cov_stride : 0;
    const double *sp = src_pts + (size_t)ss * c.n_src * 3;
    const double *sc = src_cov + (size_t)ss * c.n_src * cs;
    const int *perm = c.sperm + (size_t)ss * c.n_src;
    double *dp = c.cur_pts + (size_t)u * c.n_src * 3;
    double *dc = c.cur_cov + (size_t)u * c.n_src * cs;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < c.n_src; p += gridDim.x * blockDim.x) {
        const int i = perm[p];   // unit-local arrays live in spatially sorted order
        for (int k = 0; k < 3; ++k) dp[(size_t)p * 3 + k] = sp[(size_t)i * 3 + k];
        for (int k = 0; k < cs; ++k) dc[(size_t)p * cs + k] = sc[(size_t)i * cs + k];
    }
}

// perm of set 0 copied to sets 1.. (src_rigid_copies).  grid (ceil(n/256), sets - 1)
__global__ void __launch_bounds__(256) replicate_perm_kernel(int *__restrict__ perm, int n)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) perm[(size_t)(blockIdx.y + 1) * n + i] = perm[i];
}

// one block per target set: bounding-box centre, then centred float4 table.
__global__ void __launch_bounds__(1024) prep_targets_kernel(Ctx c)
{
    __shared__ double red[6][32];
    __shared__ double ctr[3];
    __shared__ float rmax_s[32];
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double *pts = c.tgt_pts + (size_t)s * c.n_tgt * 3;
    double mn[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, mx[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
    for (int i = tid; i < c.n_tgt; i += blockDim.x)
        for (int k = 0; k < 3; ++k) {
            const double v = pts[(size_t)i * 3 + k];
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
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a = fmin(a, red[tid][w]); b = fmax(b, red[3 + tid][w]); }
        ctr[tid] = c.n_tgt > 0 ? 0.5 * (a + b) : 0.0;
    }
    __syncthreads();
    const double cx = ctr[0], cy = ctr[1], cz = ctr[2];
    float4 *out = c.tgt4 + (size_t)s * c.nt_pad;
    const int *perm = c.tperm + (size_t)s * c.nt_pad;
    float rm = 0.f;
    for (int pos = tid; pos < c.nt_pad; pos += blockDim.x) {
        float4 v;
        const int i = perm[pos];   // table in spatially sorted order
        if (i >= 0) {
            const float fx = (float)(pts[(size_t)i * 3] - cx);
            const float fy = (float)(pts[(size_t)i * 3 + 1] - cy);
            const float fz = (float)(pts[(size_t)i * 3 + 2] - cz);
            const double w = (double)fx * fx + (double)fy * fy + (double)fz * fz;
            v = make_float4(-2.f * fx, -2.f * fy, -2.f * fz, (float)w);
            rm = fmaxf(rm, (float)sqrt(w));
        } else {
            v = make_float4(0.f, 0.f, 0.f, 3.0e38f);  // sentinel: never the minimum
        }
        out[pos] = v;
    }
    for (int o = 16; o > 0; o >>= 1) rm = fmaxf(rm, __shfl_xor_sync(0xffffffffu, rm, o));
    if (lane == 0) rmax_s[wid] = rm;
    __syncthreads();   // also orders the table writes before the sphere pass below
    if (tid == 0) {
        float r = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, rmax_s[w]);
        TgtMeta m;
        m.cx = cx; m.cy = cy; m.cz = cz;
        m.rmax = (double)r * (1.0 + 1e-6) + 1e-300;
        c.tgt_meta[s] = m;
    }
    // bounding sphere of every 32-target chunk (centred FP32 coordinates, radius rounded up)
    float4 *sph = c.chunk_sph + (size_t)s * c.n_chunks;
    for (int ch = wid; ch < c.n_chunks; ch += (int)(blockDim.x >> 5)) {
        const float4 t = out[(size_t)ch * NN_CH + lane];
        const bool valid = t.w < 1.0e38f;
        const float x = -0.5f * t.x, y = -0.5f * t.y, z = -0.5f * t.z;
        float lo[3] = {valid ? x : CUDART_INF_F, valid ? y : CUDART_INF_F, valid ? z : CUDART_INF_F};
        float hi[3] = {valid ? x : -CUDART_INF_F, valid ? y : -CUDART_INF_F, valid ? z : -CUDART_INF_F};
        for (int k = 0; k < 3; ++k)
            for (int o = 16; o > 0; o >>= 1) {
                lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
                hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
            }
        const float mx = 0.5f * (lo[0] + hi[0]), my = 0.5f * (lo[1] + hi[1]), mz = 0.5f * (lo[2] + hi[2]);
        float r = valid ? sqrtf((x - mx) * (x - mx) + (y - my) * (y - my) + (z - mz) * (z - mz)) : 0.f;
        for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
        if (lane == 0) {
            sph[ch] = make_float4(mx, my, mz, r * (1.f + 1e-5f) + 1e-30f);
            // the chunk's bounding box as half extents about the same centre, rounded up
            c.chunk_half[(size_t)s * c.n_chunks + ch] =
                make_float4(fmaxf(hi[0] - mx, mx - lo[0]) * (1.f + 1e-6f) + 1e-30f, fmaxf(hi[1] - my, my - lo[1]) * (1.f + 1e-6f) + 1e-30f,
                            fmaxf(hi[2] - mz, mz - lo[2]) * (1.f + 1e-6f) + 1e-30f, 0.f);
        }
    }
    __syncthreads();
    // second level: one sphere around each group of 32 chunk spheres
    float4 *ssph = c.super_sph + (size_t)s * c.n_super;
    for (int g = wid; g < c.n_super; g += (int)(blockDim.x >> 5)) {
        const int ch = g * 32 + lane;
        const bool valid = ch < c.n_chunks;
        const float4 t = valid ? sph[ch] : make_float4(0.f, 0.f, 0.f, 0.f);
        float lo[3] = {valid ? t.x - t.w : CUDART_INF_F, valid ? t.y - t.w : CUDART_INF_F, valid ? t.z - t.w : CUDART_INF_F};
        float hi[3] = {valid ? t.x + t.w : -CUDART_INF_F, valid ? t.y + t.w : -CUDART_INF_F, valid ? t.z + t.w : -CUDART_INF_F};
        for (int k = 0; k < 3; ++k)
            for (int o = 16; o > 0; o >>= 1) {
                lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
                hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
            }
        const float mx = 0.5f * (lo[0] + hi[0]), my = 0.5f * (lo[1] + hi[1]), mz = 0.5f * (lo[2] + hi[2]);
        float r = valid ? sqrtf((t.x - mx) * (t.x - mx) + (t.y - my) * (t.y - my) + (t.z - mz) * (t.z - mz)) + t.w : 0.f;
        for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
        if (lane == 0) ssph[g] = make_float4(mx, my, mz, r * (1.f + 1e-5f) + 1e-30f);
    }
}

// ----------------------------------------------------------------------------
// per-round update: p <- U p, C <- R C R^T (or axis <- R axis), centred FP32 copy
// (PointCloud::Transform of the moving source, Registration.cpp)
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) update_kernel(Ctx c, int round)
{
    const int u = blockIdx.y;
    const UnitState &st = c.state[u];
    if (st.done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n_src) return;
    const size_t oi = (size_t)u * c.n_src + i;
    double *p = c.cur_pts + oi * 3;
    double x = p[0], y = p[1], z = p[2];
    if (st.has_update) {
        transform_point(st.U, x, y, z);
        p[0] = x; p[1] = y; p[2] = z;
        if (c.cov_mode == 0) {
            double *cp = c.cur_cov + oi * 6;
            store_sym3(cp, rotate_sym3(st.U, load_sym3(cp)));
        } else if (c.cov_mode == 1) {
            double *ap = c.cur_cov + oi * 3;
            const double ax = ap[0], ay = ap[1], az = ap[2];
            ap[0] = st.U[0] * ax + st.U[1] * ay + st.U[2] * az;
            ap[1] = st.U[4] * ax + st.U[5] * ay + st.U[6] * az;
            ap[2] = st.U[8] * ax + st.U[9] * ay + st.U[10] * az;
        }
    }
    const TgtMeta tm = c.tgt_meta[st.tgt_set];
    // upper bound of this point's nearest-neighbour distance that matters: the distance to the
    // target it was matched to in the previous round, never more than the correspondence cut-off
    float dub = c.dcap;
    if (round >= 1) {
        const int j = c.corr[oi];
        if (j >= 0) {
            const double *tp = c.tgt_pts + ((size_t)st.tgt_set * c.n_tgt + j) * 3;
            const double dx = x - tp[0], dy = y - tp[1], dz = z - tp[2];
            const float d = __double2float_ru(sqrt(dx * dx + dy * dy + dz * dz)) * (1.f + 1e-6f);
            dub = fminf(dub, d);
        }
    }
    c.cur_qf[oi] = make_float4((float)(x - tm.cx), (float)(y - tm.cy), (float)(z - tm.cz), dub);
}

// ----------------------------------------------------------------------------
// 1-NN sweep: FP32 filter on the FMA pipe.  Q queries per thread (Q/2 packed pairs).
// ----------------------------------------------------------------------------
template <int Q, int TJ, int MODE>
__global__ void __launch_bounds__(NN_THREADS) nn_sweep_kernel(Ctx c, int round)
{
    __shared__ __align__(16) float4 tile[2][NN_TILE];
    const int u = blockIdx.z, split = blockIdx.y, tid = threadIdx.x;
    const UnitState &st = c.state[u];
    if (st.done && !c.ignore_done) return;
    const float4 *tg = c.tgt4 + (size_t)st.tgt_set * c.nt_pad;

    // this thread's Q queries: i = qbase + k*NN_THREADS + tid
    const int qbase = blockIdx.x * (NN_THREADS * Q);
    const float4 *qsrc = c.cur_qf + (size_t)u * c.n_src;
    uint64_t qx[Q / 2], qy[Q / 2], qz[Q / 2];
#pragma unroll
    for (int h = 0; h < Q / 2; ++h) {
        const int i0 = qbase + (2 * h) * NN_THREADS + tid, i1 = i0 + NN_THREADS;
        const float4 a = i0 < c.n_src ? qsrc[i0] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 b = i1 < c.n_src ? qsrc[i1] : make_float4(0.f, 0.f, 0.f, 0.f);
        qx[h] = pack_f2(a.x, b.x);
        qy[h] = pack_f2(a.y, b.y);
        qz[h] = pack_f2(a.z, b.z);
    }
    float m1[Q], m2[Q];
    int c1[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) { m1[k] = CUDART_INF_F; m2[k] = CUDART_INF_F; c1[k] = 0; }

    const int cb = split * c.chunks_per_split;
    const int ce = min(c.n_chunks, cb + c.chunks_per_split);
    const int n_tiles = (ce - cb + NN_TILE_CH - 1) / NN_TILE_CH;

    auto issue = [&](int t) {
        if (t < n_tiles) {
            const int c0 = cb + t * NN_TILE_CH;
            const int cnt = min(NN_TILE_CH, ce - c0) * NN_CH;
            const float4 *src = tg + (size_t)c0 * NN_CH;
            float4 *dst = tile[t & 1];
            for (int e = tid; e < cnt; e += NN_THREADS) cp_async16(dst + e, src + e);
        }
        cp_async_commit();
    };
    issue(0);
    for (int t = 0; t < n_tiles; ++t) {
        issue(t + 1);
        cp_async_wait<1>();
        __syncthreads();
        const int c0 = cb + t * NN_TILE_CH;
        const int nch = min(NN_TILE_CH, ce - c0);
        const float4 *tl = tile[t & 1];
        for (int ch = 0; ch < nch; ++ch) {
            const float4 *tc = tl + ch * NN_CH;
            float mc[Q];
#pragma unroll
            for (int k = 0; k < Q; ++k) mc[k] = CUDART_INF_F;
#pragma unroll
            for (int j = 0; j < NN_CH; j += TJ) {
                // TJ targets per step: consecutive FFMA2s share the packed query operand so it
                // is served from the operand-reuse cache (register-bank reads bound the FFMA2
                // rate otherwise: 2 even + 2 odd registers per cycle pair).
                float4 t[TJ];
#pragma unroll
                for (int r = 0; r < TJ; ++r) t[r] = tc[j + r];
#pragma unroll
                for (int h = 0; h < Q / 2; ++h) {
                    if (MODE == 0) {
                        // all three stages packed
                        uint64_t a[TJ];
#pragma unroll
                        for (int r = 0; r < TJ; ++r) a[r] = fma2_ss(qz[h], t[r].z, t[r].w);
#pragma unroll
                        for (int r = 0; r < TJ; ++r) a[r] = fma2_sv(qy[h], t[r].y, a[r]);
#pragma unroll
                        for (int r = 0; r < TJ; ++r) a[r] = fma2_sv(qx[h], t[r].x, a[r]);
#pragma unroll
                        for (int r = 0; r < TJ; r += 2) {
                            float l0, h0, l1, h1;
                            unpack_f2(a[r], l0, h0);
                            unpack_f2(a[r + 1], l1, h1);
                            mc[2 * h] = min3f(mc[2 * h], l0, l1);
                            mc[2 * h + 1] = min3f(mc[2 * h + 1], h0, h1);
                        }
                    } else {
                        // stage 1 packed (two broadcast scalars: 2 even + 2 odd register reads, full
                        // rate); stages 2-3 scalar FFMA (accumulator + query register, target scalar
                        // from the operand-reuse cache) -- a packed FFMA2 with a register-pair
                        // accumulator needs 5 register reads and issues every 3 cycles.
                        float lo[TJ], hi[TJ];
                        float qyl, qyh, qxl, qxh, qzl, qzh;
                        unpack_f2(qy[h], qyl, qyh);
                        unpack_f2(qx[h], qxl, qxh);
                        unpack_f2(qz[h], qzl, qzh);
#pragma unroll
                        for (int r = 0; r < TJ; ++r) {
                            if (MODE == 1) {
                                unpack_f2(fma2_ss(qz[h], t[r].z, t[r].w), lo[r], hi[r]);
                            } else {
                                lo[r] = __fmaf_rn(qzl, t[r].z, t[r].w);
                                hi[r] = __fmaf_rn(qzh, t[r].z, t[r].w);
                            }
                        }
#pragma unroll
                        for (int r = 0; r < TJ; ++r) {
                            lo[r] = __fmaf_rn(qyl, t[r].y, lo[r]);
                            hi[r] = __fmaf_rn(qyh, t[r].y, hi[r]);
                        }
#pragma unroll
                        for (int r = 0; r < TJ; ++r) {
                            lo[r] = __fmaf_rn(qxl, t[r].x, lo[r]);
                            hi[r] = __fmaf_rn(qxh, t[r].x, hi[r]);
                        }
#pragma unroll
                        for (int r = 0; r < TJ; r += 2) {
                            mc[2 * h] = min3f(mc[2 * h], lo[r], lo[r + 1]);
                            mc[2 * h + 1] = min3f(mc[2 * h + 1], hi[r], hi[r + 1]);
                        }
                    }
                }
            }
            const int chunk = c0 + ch;
#pragma unroll
            for (int k = 0; k < Q; ++k) {
                m2[k] = fminf(m2[k], fmaxf(m1[k], mc[k]));
                c1[k] = mc[k] < m1[k] ? chunk : c1[k];
                m1[k] = fminf(m1[k], mc[k]);
            }
        }
        __syncthreads();
    }
    const size_t pbase = ((size_t)u * c.splits + split) * c.n_src;
#pragma unroll
    for (int k = 0; k < Q; ++k) {
        const int i = qbase + k * NN_THREADS + tid;
        if (i < c.n_src) {
            c.part_m1[pbase + i] = m1[k];
            c.part_c1[pbase + i] = c1[k];
            c.part_m2[pbase + i] = m2[k];
        }
    }
    if (tid == 0 && !c.ignore_done) {
        const int nq = min(NN_THREADS * Q, c.n_src - qbase);
        atomicAdd(c.visit_count + round, (unsigned long long)(ce - cb) * NN_CH * nq);
    }
}

// ----------------------------------------------------------------------------
// 1-NN sweep with chunk culling.  Same FP32 filter arithmetic as nn_sweep_kernel, but every
// warp owns 32*Q spatially consecutive queries (a compact patch), reads the target chunks it
// needs straight through L1 (broadcast float4 loads) and skips every 32-target chunk whose
// bounding sphere lies farther from the patch than any of its queries can still be matched:
//   dist(patch centre, chunk centre) - r_chunk  >  r_patch + max_q d_ub(q)
// where d_ub(q) is the distance to q's previous correspondence (capped at the correspondence
// cut-off).  Skipped chunks provably hold no nearest neighbour, so the (best chunk, second
// best) certificate over the visited chunks is still a certificate over all targets.
// ----------------------------------------------------------------------------
// warps per block (template parameter W): warps are independent workers, 2 by default
#ifndef ORPCD_PH1_HOIST
#define ORPCD_PH1_HOIST 1
#endif
#ifndef ORPCD_PH5_HOIST
#define ORPCD_PH5_HOIST 1
#endif
#ifndef ORPCD_SWEEP_ROLLED
#define ORPCD_SWEEP_ROLLED 1   // per-chunk arithmetic of the culled sweep rolled over its 8-target sub-chunks (code size)
#endif
#ifndef ORPCD_CULL_AABB
#define ORPCD_CULL_AABB 1      // chunk bounding boxes (0: bounding spheres) in the visit-list tests of the culled sweep
#endif
constexpr int CULL_RING = 4;       // chunk buffers per warp (512 B each), filled by cp.async
constexpr int CULL_LIST = 1024;    // visit-list capacity per pass (chunk ids)
constexpr int RED_STRIDE_MMA = 128;   // doubles per patch partial of the tensor-core Jacobian sums (two 8x8 tiles)

// a worker warp's private shared memory in the fused round kernels: the chunk ring and the visit list of the sweep,
// then (the sweep is over by then) the 96 x 8 float64 records of the Jacobian sums, then the unit totals
struct __align__(16) WarpSmem {
    float4 ring[CULL_RING][NN_CH];      // 2 KB
    unsigned short vl[CULL_LIST];       // 2 KB
};
static_assert(sizeof(WarpSmem) == 4096, "WarpSmem holds 64 records of 8 doubles");
// (shared memory is kept small on purpose: it comes out of the same 256 KB as the L1 cache, and the target tables,
// chunk spheres and permutations of the sweep live on L1 hits -- 6 KB per warp with the large carve-out measured
// 14.9 ms per 180-unit batch against 12.8 ms with 4 KB)

// The sweep of one warp over its patch.  qv[k] = (centred FP32 query, d_ub) of query k*32+lane of the patch
// (lanes past the end of the cloud carry a copy of another query of the same patch); chunks [cb, ce) of the
// target table; results per query: best chunk minimum m1, its chunk c1, second-best chunk minimum m2.
// Returns the number of chunks the warp visited.  `ring` / `vl` are this warp's private shared memory.
// SUB = false: (m1, c1, m2) fold per 32-target chunk, c1 = chunk id.  SUB = true: fold per 8-target sub-chunk,
// c1 = chunk * 4 + sub-chunk, m2 = second-best sub-chunk minimum (the certificate works at either granularity; the
// finer one leaves the caller 8 candidates to re-score instead of 32).
// HALF (32-point patches, Q = 2): both half-warps hold the SAME pair of queries per lane (32 queries per warp) and
// split the 32 targets of a chunk by parity -- lane l takes targets 2 i + (l >> 4) -- so a warp spends half the FMA
// work per chunk on half the queries; each half folds the 8-target sub-chunks it sees, id = chunk * 4 + parity * 2 +
// (first | second 16 targets), and the caller merges the two halves' (m1, c1, m2).
// (`half` is a run-time argument: the set-up and the visit list are the same code for both patch sizes and exist once
// in the kernel -- the persistent kernel is bound by instruction fetch, see DESIGN 4.1 -- and only the per-chunk
// arithmetic, sweep_chunk<>, is instantiated per layout.)
template <int Q, bool SUB, bool HALF>
__device__ __forceinline__ void sweep_chunk(const float4 *__restrict__ tc, int chunk, int par, const uint64_t (&qx)[Q / 2],
                                            const uint64_t (&qy)[Q / 2], const uint64_t (&qz)[Q / 2], float (&m1)[Q],
                                            int (&c1)[Q], float (&m2)[Q])
{
    constexpr int TJ = 4;
    static_assert(!HALF || (Q == 2 && SUB), "the half-warp split is written for one query pair per lane and sub-chunk folds");
    float mc[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) mc[q] = CUDART_INF_F;
    if constexpr (SUB) {
        // one 8-target sub-chunk per trip of a ROLLED loop (4 trips, 2 for the half-warp split): the unrolled form was
        // 150 instructions per instantiation of straight-line code in the hottest loop of an instruction-fetch-bound kernel
        constexpr int STRIDE = HALF ? 2 : 1;
        const float4 *ts = tc + par;
        int id = chunk * 4 + par * 2;
#if ORPCD_SWEEP_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int sc = 0; sc < (HALF ? 2 : 4); ++sc, ts += 8 * STRIDE, ++id) {
#pragma unroll
            for (int j = 0; j < 8; j += TJ) {
                float4 t[TJ];
#pragma unroll
                for (int r = 0; r < TJ; ++r) t[r] = ts[(j + r) * STRIDE];
#pragma unroll
                for (int h = 0; h < Q / 2; ++h) {
                    uint64_t a[TJ];
#pragma unroll
                    for (int r = 0; r < TJ; ++r) a[r] = fma2_ss(qz[h], t[r].z, t[r].w);
#pragma unroll
                    for (int r = 0; r < TJ; ++r) a[r] = fma2_sv(qy[h], t[r].y, a[r]);
#pragma unroll
                    for (int r = 0; r < TJ; ++r) a[r] = fma2_sv(qx[h], t[r].x, a[r]);
#pragma unroll
                    for (int r = 0; r < TJ; r += 2) {
                        float l0, h0, l1, h1;
                        unpack_f2(a[r], l0, h0);
                        unpack_f2(a[r + 1], l1, h1);
                        mc[2 * h] = min3f(mc[2 * h], l0, l1);
                        mc[2 * h + 1] = min3f(mc[2 * h + 1], h0, h1);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                m2[q] = fminf(m2[q], fmaxf(m1[q], mc[q]));
                c1[q] = mc[q] < m1[q] ? id : c1[q];
                m1[q] = fminf(m1[q], mc[q]);
                mc[q] = CUDART_INF_F;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < NN_CH; j += TJ) {
            float4 t[TJ];
#pragma unroll
            for (int r = 0; r < TJ; ++r) t[r] = tc[j + r];
#pragma unroll
            for (int h = 0; h < Q / 2; ++h) {
                uint64_t a[TJ];
#pragma unroll
                for (int r = 0; r < TJ; ++r) a[r] = fma2_ss(qz[h], t[r].z, t[r].w);
#pragma unroll
                for (int r = 0; r < TJ; ++r) a[r] = fma2_sv(qy[h], t[r].y, a[r]);
#pragma unroll
                for (int r = 0; r < TJ; ++r) a[r] = fma2_sv(qx[h], t[r].x, a[r]);
#pragma unroll
                for (int r = 0; r < TJ; r += 2) {
                    float l0, h0, l1, h1;
                    unpack_f2(a[r], l0, h0);
                    unpack_f2(a[r + 1], l1, h1);
                    mc[2 * h] = min3f(mc[2 * h], l0, l1);
                    mc[2 * h + 1] = min3f(mc[2 * h + 1], h0, h1);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            m2[q] = fminf(m2[q], fmaxf(m1[q], mc[q]));
            c1[q] = mc[q] < m1[q] ? chunk : c1[q];
            m1[q] = fminf(m1[q], mc[q]);
        }
    }
}

template <int Q, bool SUB>
__device__ __forceinline__ int cull_sweep_core(const Ctx &c, const float4 *__restrict__ tg, const float4 *__restrict__ sph,
                                               const float4 *__restrict__ ssph, int n_super, float rmax_tgt, int cb,
                                               int ce, const float4 (&qv)[Q],
                                               float4 (*ring)[NN_CH], unsigned short *vl, float (&m1)[Q], int (&c1)[Q],
                                               float (&m2)[Q], bool HALF = false)
{
    const int lane = threadIdx.x & 31;
    const int par = HALF ? (lane >> 4) : 0;      // which targets of a chunk this half-warp takes
    uint64_t qx[Q / 2], qy[Q / 2], qz[Q / 2];
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    float D = 0.f, rw = 0.f, pcx, pcy, pcz;
    float dq[Q];
    {
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            dq[k] = qv[k].w;
            lo[0] = fminf(lo[0], qv[k].x); hi[0] = fmaxf(hi[0], qv[k].x);
            lo[1] = fminf(lo[1], qv[k].y); hi[1] = fmaxf(hi[1], qv[k].y);
            lo[2] = fminf(lo[2], qv[k].z); hi[2] = fmaxf(hi[2], qv[k].z);
            D = fmaxf(D, qv[k].w);
        }
#pragma unroll
        for (int h = 0; h < Q / 2; ++h) {
            qx[h] = pack_f2(qv[2 * h].x, qv[2 * h + 1].x);
            qy[h] = pack_f2(qv[2 * h].y, qv[2 * h + 1].y);
            qz[h] = pack_f2(qv[2 * h].z, qv[2 * h + 1].z);
        }
        // warp-wide min / max in ONE instruction each (redux.sync.f32, sm_100a: SASS CREDUX) instead of five
        // shuffle + min steps: the patch's bounding box, largest reach and radius are eight reductions per item
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = warp_min_f32(lo[k]); hi[k] = warp_max_f32(hi[k]); }
        D = warp_max_f32(D);
        pcx = 0.5f * (lo[0] + hi[0]); pcy = 0.5f * (lo[1] + hi[1]); pcz = 0.5f * (lo[2] + hi[2]);
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            const float dx = qv[k].x - pcx, dy = qv[k].y - pcy, dz = qv[k].z - pcz;
            rw = fmaxf(rw, sqrtf(dx * dx + dy * dy + dz * dz));
        }
        rw = warp_max_f32(rw);
    }
    // FP32 slack: every quantity below carries a few ulps of the coordinate magnitude
    const float mag = fabsf(pcx) + fabsf(pcy) + fabsf(pcz) + rw + rmax_tgt;
    const float reach_box = D * (1.f + 4e-6f) + 8e-6f * mag;   // a query is inside the box: only its own reach counts
#pragma unroll
    for (int k = 0; k < Q; ++k) dq[k] = dq[k] * (1.f + 4e-6f) + 8e-6f * mag;   // per-query reach, inflated
    uint64_t dq2[Q / 2];
#pragma unroll
    for (int h = 0; h < Q / 2; ++h) dq2[h] = pack_f2(dq[2 * h], dq[2 * h + 1]);
#if ORPCD_CULL_AABB
    const float4 *__restrict__ half = c.chunk_half + (sph - c.chunk_sph);   // same (target set, chunk) indexing as `sph`
    uint64_t dlim2[Q / 2];     // squared reach of each query, with the slack of the squared comparison
#pragma unroll
    for (int h = 0; h < Q / 2; ++h) dlim2[h] = mul2(mul2(dq2[h], dq2[h]), pack_f2(1.f + 1e-6f, 1.f + 1e-6f));
#endif

#pragma unroll
    for (int k = 0; k < Q; ++k) { m1[k] = CUDART_INF_F; m2[k] = CUDART_INF_F; c1[k] = 0; }

    int visited = 0;
    for (int pass = cb; pass < ce; pass += CULL_LIST) {
        // 1) visit list of this pass.  Coarse test, lane-parallel over chunks: chunk sphere against the
        //    patch sphere inflated by the largest d_ub.  Survivors get the exact test: a chunk is needed
        //    iff SOME query q of the patch has it within reach, dist(q, chunk centre) <= d_ub(q) + r_chunk.
        const int pe = min(ce, pass + CULL_LIST);
        int nv = 0;
        // level 1: the 32 groups of 32 chunks of this pass (pass boundaries are multiples of 1024 chunks)
        bool gneed = false;
        {
            const int g = pass / 32 + lane;
            if (g < n_super && g * 32 < pe) {
                const float4 sp = __ldg(ssph + g);
                const float bx = fmaxf(fmaxf(lo[0] - sp.x, sp.x - hi[0]), 0.f), by = fmaxf(fmaxf(lo[1] - sp.y, sp.y - hi[1]), 0.f);
                const float bz = fmaxf(fmaxf(lo[2] - sp.z, sp.z - hi[2]), 0.f);
                gneed = !(sqrtf(__fmaf_rn(bx, bx, __fmaf_rn(by, by, bz * bz))) - sp.w > reach_box);
            }
        }
        unsigned gbal = __ballot_sync(0xffffffffu, gneed);
        while (gbal) {
            const int gl = __ffs(gbal) - 1;
            gbal &= gbal - 1;
            const int base = pass + gl * 32;
            const int chunk = base + lane;
            bool need = false;
            float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
#if ORPCD_CULL_AABB
            float4 hf = make_float4(0.f, 0.f, 0.f, 0.f);
            if (chunk < pe) {
                sp = __ldg(sph + chunk);
                hf = __ldg(half + chunk);
                // distance between the chunk's bounding box and the patch's
                const float bx = fmaxf(fmaxf(lo[0] - (sp.x + hf.x), (sp.x - hf.x) - hi[0]), 0.f);
                const float by = fmaxf(fmaxf(lo[1] - (sp.y + hf.y), (sp.y - hf.y) - hi[1]), 0.f);
                const float bz = fmaxf(fmaxf(lo[2] - (sp.z + hf.z), (sp.z - hf.z) - hi[2]), 0.f);
                need = !(sqrtf(__fmaf_rn(bx, bx, __fmaf_rn(by, by, bz * bz))) > reach_box);
            }
#else
            if (chunk < pe) {
                sp = __ldg(sph + chunk);
                // distance from the chunk centre to the patch's bounding box (tighter than its bounding sphere)
                const float bx = fmaxf(fmaxf(lo[0] - sp.x, sp.x - hi[0]), 0.f), by = fmaxf(fmaxf(lo[1] - sp.y, sp.y - hi[1]), 0.f);
                const float bz = fmaxf(fmaxf(lo[2] - sp.z, sp.z - hi[2]), 0.f);
                need = !(sqrtf(__fmaf_rn(bx, bx, __fmaf_rn(by, by, bz * bz))) - sp.w > reach_box);
            }
#endif
            const unsigned bal = __ballot_sync(0xffffffffu, need);
            if (bal) {
                // The exact test of the box survivors.  Their spheres are staged in the (still idle) first ring slot
                // and read back as one broadcast LDS.128 per candidate; every lane tests its own query pair against
                // two candidates per step (independent chains) and keeps a private bit mask; ONE warp-wide OR per
                // group of 32 chunks then replaces the vote + branch + lane-0 store per candidate of the first
                // version (which was 13 % of the persistent kernel's stall samples).  List order: ascending chunk id.
#if ORPCD_CULL_AABB
                // exact test against the chunk's BOUNDING BOX: some query q of the patch has dist(q, box) <= d_ub(q).
                // (Hilbert-consecutive targets of a surface scan form flat pieces: their boxes hold a third less empty
                // space than their spheres, 17.7 -> 12.2 listed chunks per patch on the bench clouds.)
                ring[0][lane] = sp;
                ring[1][lane] = hf;
                __syncwarp();
                unsigned mymask = 0u, b = bal;
                while (b) {
                    const int l0 = __ffs(b) - 1;
                    b &= b - 1;
                    const int l1 = b ? __ffs(b) - 1 : l0;   // (a repeat of the first candidate when it was the last)
                    b &= b - 1;
                    const float4 s0 = ring[0][l0], s1 = ring[0][l1], h0 = ring[1][l0], h1 = ring[1][l1];
                    bool mine0 = false, mine1 = false;
#pragma unroll
                    for (int h = 0; h < Q / 2; ++h) {
                        // both queries of the pair: sum_k max(|q_k - m_k| - h_k, 0)^2 <= d_ub(q)^2 (1 + 1e-6)
                        float ex[2][2], ey[2][2], ez[2][2];
                        unpack_f2(add2_s(qx[h], -s0.x), ex[0][0], ex[0][1]);
                        unpack_f2(add2_s(qy[h], -s0.y), ey[0][0], ey[0][1]);
                        unpack_f2(add2_s(qz[h], -s0.z), ez[0][0], ez[0][1]);
                        unpack_f2(add2_s(qx[h], -s1.x), ex[1][0], ex[1][1]);
                        unpack_f2(add2_s(qy[h], -s1.y), ey[1][0], ey[1][1]);
                        unpack_f2(add2_s(qz[h], -s1.z), ez[1][0], ez[1][1]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            ex[0][e] = fmaxf(fabsf(ex[0][e]) - h0.x, 0.f); ey[0][e] = fmaxf(fabsf(ey[0][e]) - h0.y, 0.f);
                            ez[0][e] = fmaxf(fabsf(ez[0][e]) - h0.z, 0.f);
                            ex[1][e] = fmaxf(fabsf(ex[1][e]) - h1.x, 0.f); ey[1][e] = fmaxf(fabsf(ey[1][e]) - h1.y, 0.f);
                            ez[1][e] = fmaxf(fabsf(ez[1][e]) - h1.z, 0.f);
                        }
                        const uint64_t X0 = pack_f2(ex[0][0], ex[0][1]), Y0 = pack_f2(ey[0][0], ey[0][1]), Z0 = pack_f2(ez[0][0], ez[0][1]);
                        const uint64_t X1 = pack_f2(ex[1][0], ex[1][1]), Y1 = pack_f2(ey[1][0], ey[1][1]), Z1 = pack_f2(ez[1][0], ez[1][1]);
                        const uint64_t d20 = fma2(X0, X0, fma2(Y0, Y0, mul2(Z0, Z0)));
                        const uint64_t d21 = fma2(X1, X1, fma2(Y1, Y1, mul2(Z1, Z1)));
                        float da, db, la, lb;
                        unpack_f2(dlim2[h], la, lb);
                        unpack_f2(d20, da, db);
                        mine0 |= (da <= la) | (db <= lb);
                        unpack_f2(d21, da, db);
                        mine1 |= (da <= la) | (db <= lb);
                    }
                    mymask |= (mine0 ? 1u << l0 : 0u) | (mine1 ? 1u << l1 : 0u);
                }
#else
                ring[0][lane] = sp;
                __syncwarp();
                unsigned mymask = 0u, b = bal;
                while (b) {
                    const int l0 = __ffs(b) - 1;
                    b &= b - 1;
                    const int l1 = b ? __ffs(b) - 1 : l0;   // (a repeat of the first candidate when it was the last)
                    b &= b - 1;
                    const float4 s0 = ring[0][l0], s1 = ring[0][l1];
                    bool mine0 = false, mine1 = false;
#pragma unroll
                    for (int h = 0; h < Q / 2; ++h) {
                        // both queries of the pair at once: |q - s|^2 <= (d_ub(q) + r)^2 (1 + 1e-6), packed FP32
                        const uint64_t ax0 = add2_s(qx[h], -s0.x), ay0 = add2_s(qy[h], -s0.y), az0 = add2_s(qz[h], -s0.z);
                        const uint64_t ax1 = add2_s(qx[h], -s1.x), ay1 = add2_s(qy[h], -s1.y), az1 = add2_s(qz[h], -s1.z);
                        const uint64_t d20 = fma2(ax0, ax0, fma2(ay0, ay0, mul2(az0, az0)));
                        const uint64_t d21 = fma2(ax1, ax1, fma2(ay1, ay1, mul2(az1, az1)));
                        const uint64_t tt0 = add2_s(dq2[h], s0.w), tt1 = add2_s(dq2[h], s1.w);
                        const uint64_t lim0 = mul2(mul2(tt0, tt0), pack_f2(1.f + 1e-6f, 1.f + 1e-6f));
                        const uint64_t lim1 = mul2(mul2(tt1, tt1), pack_f2(1.f + 1e-6f, 1.f + 1e-6f));
                        float da, db, la, lb;
                        unpack_f2(d20, da, db);
                        unpack_f2(lim0, la, lb);
                        mine0 |= (da <= la) | (db <= lb);
                        unpack_f2(d21, da, db);
                        unpack_f2(lim1, la, lb);
                        mine1 |= (da <= la) | (db <= lb);
                    }
                    mymask |= (mine0 ? 1u << l0 : 0u) | (mine1 ? 1u << l1 : 0u);
                }
#endif
                const unsigned keep = __reduce_or_sync(0xffffffffu, mymask);
                if ((keep >> lane) & 1u) vl[nv + __popc(keep & ((1u << lane) - 1u))] = (unsigned short)(base + lane - pass);
                nv += __popc(keep);
#ifdef ORPCD_QUEUE_STATS
                if (lane == 0) { atomicAdd(c.q_stats + 632, (unsigned long long)__popc(bal)); atomicAdd(c.q_stats + 633, (unsigned long long)__popc(keep)); }
#endif
            }
        }
        __syncwarp();
        visited += nv;
        // 2) stream the listed chunks through the warp's ring (one 16-byte cp.async per lane per chunk)
        auto issue = [&](int k) {
            if (k < nv) cp_async16(&ring[k % CULL_RING][lane], tg + (size_t)(pass + vl[k]) * NN_CH + lane);
            cp_async_commit();
        };
#pragma unroll
        for (int k = 0; k < CULL_RING - 1; ++k) issue(k);
        for (int k = 0; k < nv; ++k) {
            issue(k + CULL_RING - 1);
            cp_async_wait<CULL_RING - 1>();
            __syncwarp();
            const float4 *tc = ring[k % CULL_RING];
            const int chunk = pass + vl[k];
            if constexpr (Q == 2 && SUB) {
                if (HALF) sweep_chunk<Q, SUB, true>(tc, chunk, par, qx, qy, qz, m1, c1, m2);
                else sweep_chunk<Q, SUB, false>(tc, chunk, 0, qx, qy, qz, m1, c1, m2);
            } else {
                sweep_chunk<Q, SUB, false>(tc, chunk, 0, qx, qy, qz, m1, c1, m2);
            }
            __syncwarp();   // everyone is done with this slot before it is refilled
        }
        cp_async_wait<0>();
        __syncwarp();
    }
    return visited;
}

template <int Q, int CULL_WARPS>
__global__ void __launch_bounds__(CULL_WARPS * 32) nn_sweep_cull_kernel(Ctx c, int round)
{
    __shared__ __align__(16) float4 ring[CULL_WARPS][CULL_RING][NN_CH];
    __shared__ unsigned short vlist[CULL_WARPS][CULL_LIST];
    const int u = blockIdx.z, split = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const UnitState &st = c.state[u];
    if (st.done && !c.ignore_done) return;
    const int wbase = (blockIdx.x * CULL_WARPS + warp) * (32 * Q);
    if (wbase >= c.n_src) return;
    const float4 *tg = c.tgt4 + (size_t)st.tgt_set * c.nt_pad;
    const float4 *sph = c.chunk_sph + (size_t)st.tgt_set * c.n_chunks;
    const float4 *qsrc = c.cur_qf + (size_t)u * c.n_src;
    float4 qv[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) qv[k] = qsrc[min(wbase + k * 32 + lane, c.n_src - 1)];   // tail lanes repeat the last query
    float m1[Q], m2[Q];
    int c1[Q];
    const int cb = split * c.chunks_per_split;
    const int ce = min(c.n_chunks, cb + c.chunks_per_split);
    const int visited = cull_sweep_core<Q, false>(c, tg, sph, c.super_sph + (size_t)st.tgt_set * c.n_super, c.n_super,
                                           (float)c.tgt_meta[st.tgt_set].rmax, cb, ce, qv, ring[warp], vlist[warp], m1,
                                           c1, m2);
    const size_t pbase = ((size_t)u * c.splits + split) * c.n_src;
#pragma unroll
    for (int k = 0; k < Q; ++k) {
        const int i = wbase + k * 32 + lane;
        if (i < c.n_src) {
            c.part_m1[pbase + i] = m1[k];
            c.part_c1[pbase + i] = c1[k];
            c.part_m2[pbase + i] = m2[k];
        }
    }
    if (lane == 0 && !c.ignore_done) {
        const int nq = min(32 * Q, c.n_src - wbase);
        atomicAdd(c.visit_count + round, (unsigned long long)visited * NN_CH * nq);
    }
}

// ----------------------------------------------------------------------------
// resolve: merge split partials, certify the best chunk, pick the exact neighbour
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(128) nn_resolve_kernel(Ctx c, int round)
{
    const int u = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const UnitState &st = c.state[u];
    if (st.done) return;  // block-uniform
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&c.round_active[round], 1);
    const bool active = i < c.n_src;
    const TgtMeta tm = c.tgt_meta[st.tgt_set];
    const size_t oi = (size_t)u * c.n_src + (active ? i : 0);
    float fx = 0.f, fy = 0.f, fz = 0.f, thr = -CUDART_INF_F;
    int c1 = 0;
    bool cross = false, none = false;
    if (active) {
        const float4 qf = c.cur_qf[oi];
        fx = qf.x; fy = qf.y; fz = qf.z;
        float m1 = CUDART_INF_F, m2 = CUDART_INF_F;
        for (int s = 0; s < c.splits; ++s) {
            const size_t pi = ((size_t)u * c.splits + s) * c.n_src + i;
            const float a1 = c.part_m1[pi], a2 = c.part_m2[pi];
            const int ac = c.part_c1[pi];
            m2 = fminf(fminf(m2, a2), fmaxf(m1, a1));
            c1 = a1 < m1 ? ac : c1;
            m1 = fminf(m1, a1);
        }
        // FP32 error bound of the filter (DESIGN.md "1-NN filter bound"):
        //   |(s + |qf|^2) - D| <= eps32 * (4 rho^2 + 2 d rho),  rho = |qf| + |pf|.
        const double qn = (double)fx * fx + (double)fy * fy + (double)fz * fz;
        const double qnorm = sqrt(qn);
        const double rho_max = qnorm + tm.rmax;
        const double e_glob = (double)F32_EPS * 6.0 * rho_max * rho_max * 1.001;
        const double d_ub = sqrt(fmax((double)m1 + qn, 0.0) + e_glob) * (1.0 + 1e-6);
        const double rho = 2.0 * qnorm + d_ub + 4.0 * (double)F32_EPS * rho_max;
        const double margin = 2.02 * (double)F32_EPS * (4.0 * rho * rho + 2.0 * d_ub * rho) * 1.001 + 1e-300;
        const double lim = (double)m1 + margin;
        cross = !((double)m2 > lim);       // another chunk may hold the true neighbour
        thr = __double2float_ru(lim);
        none = !(m1 < CUDART_INF_F);       // the sweep visited no chunk: nothing within the cut-off
        if (none) cross = true;            // (keeps this query out of the ballot pass)
    }
    // Warp-cooperative pass over each query's best chunk: lane t recomputes the filter
    // value of target (chunk, t) -- same FMA chain as the sweep, so bit-identical -- and
    // the ballot counts the targets within the certified margin of the minimum.
    const float4 *tg = c.tgt4 + (size_t)st.tgt_set * c.nt_pad;
    int my_cnt = 0, my_first = 0;
    unsigned rem = __ballot_sync(0xffffffffu, active && !cross);
    while (rem) {
        const int l = __ffs(rem) - 1;
        rem &= rem - 1;
        const int cc = __shfl_sync(0xffffffffu, c1, l);
        const float qx = __shfl_sync(0xffffffffu, fx, l);
        const float qy = __shfl_sync(0xffffffffu, fy, l);
        const float qz = __shfl_sync(0xffffffffu, fz, l);
        const float th = __shfl_sync(0xffffffffu, thr, l);
        const float4 t = tg[(size_t)cc * NN_CH + lane];
        const float sv = __fmaf_rn(qx, t.x, __fmaf_rn(qy, t.y, __fmaf_rn(qz, t.z, t.w)));
        const unsigned bal = __ballot_sync(0xffffffffu, sv <= th);
        if (lane == l) { my_cnt = __popc(bal); my_first = __ffs(bal) - 1; }
    }
    if (!active) return;
    if (none) {
        c.corr[oi] = -1;
        c.corr_d2[oi] = 0.0;
        return;
    }
    if (cross || my_cnt != 1) {
        // ambiguous at FP32 resolution: exact float64 scan by a whole warp (nn_fix_kernel)
        const int slot = atomicAdd(&c.worklist_count[round], 1);
        c.worklist[slot] = (int)oi;
        c.corr[oi] = c1;   // best chunk of the filter: where nn_fix_kernel starts its exact search
        return;
    }
    const int j = c.tperm[(size_t)st.tgt_set * c.nt_pad + c1 * NN_CH + my_first];   // sorted position -> index
    const double *p = c.cur_pts + oi * 3;
    const double *tp = c.tgt_pts + ((size_t)st.tgt_set * c.n_tgt + j) * 3;
    const double best = sqdist3_rn(p[0], p[1], p[2], tp[0], tp[1], tp[2]);
    const bool ok = best < c.r2;
    c.corr[oi] = ok ? j : -1;
    c.corr_d2[oi] = ok ? best : 0.0;
}

// Exact float64 search by a whole warp for one query (x, y, z; centred FP32 copy qf).  The best chunk c1 of
// the FP32 filter is scored first; every other 32-target chunk is scored only if its bounding sphere comes
// within that distance of the query (strict test with FP32 slack, so equal-distance candidates are never
// skipped), lane-parallel over chunks for the test and over the 32 targets for the scoring.  Ties -> lowest
// original index.  All lanes return the same (best, bj); bj = 0x7fffffff if the set is empty.
__device__ __forceinline__ void exact_search_warp(const Ctx &c, int tgt_set, double x, double y, double z, float qfx,
                                                  float qfy, float qfz, int c1, double &best, int &bj)
{
    const int lane = threadIdx.x & 31;
    const double *tp = c.tgt_pts + (size_t)tgt_set * c.n_tgt * 3;
    const int *perm = c.tperm + (size_t)tgt_set * c.nt_pad;
    const float4 *sph = c.chunk_sph + (size_t)tgt_set * c.n_chunks;
    const float4 *half = c.chunk_half + (size_t)tgt_set * c.n_chunks;
    best = CUDART_INF;
    bj = 0x7fffffff;
    auto score = [&](int chunk) {
        const int j = perm[(size_t)chunk * NN_CH + lane];
        if (j >= 0) {
            const double d = sqdist3_rn(x, y, z, tp[(size_t)j * 3], tp[(size_t)j * 3 + 1], tp[(size_t)j * 3 + 2]);
            if (d < best || (d == best && j < bj)) { best = d; bj = j; }
        }
    };
    score(c1);
    double wb = best;
    for (int o = 16; o > 0; o >>= 1) wb = fmin(wb, __shfl_xor_sync(0xffffffffu, wb, o));
    const float mag = fabsf(qfx) + fabsf(qfy) + fabsf(qfz) + (float)c.tgt_meta[tgt_set].rmax;
    const float thr = __double2float_ru(sqrt(wb)) * (1.f + 4e-6f) + 8e-6f * mag;   // inf if chunk c1 was empty
    const float4 *ssph = c.super_sph + (size_t)tgt_set * c.n_super;
    for (int gbase = 0; gbase < c.n_super; gbase += 32) {
        // level 1: groups of 32 chunks
        bool gneed = false;
        if (gbase + lane < c.n_super) {
            const float4 sp = ssph[gbase + lane];
            const float sx = sp.x - qfx, sy = sp.y - qfy, sz = sp.z - qfz;
            gneed = !(sqrtf(sx * sx + sy * sy + sz * sz) - sp.w > thr);
        }
        unsigned gbal = __ballot_sync(0xffffffffu, gneed);
        while (gbal) {
            const int gl = __ffs(gbal) - 1;
            gbal &= gbal - 1;
            const int base = (gbase + gl) * 32;
            const int chunk = base + lane;
            bool need = false;
            if (chunk < c.n_chunks && chunk != c1) {
                // the chunk's bounding BOX (centre + half extents) against the query: a third fewer chunks to score
                // than with the bounding sphere
                const float4 sp = sph[chunk], hf = half[chunk];
                const float ex = fmaxf(fabsf(sp.x - qfx) - hf.x, 0.f), ey = fmaxf(fabsf(sp.y - qfy) - hf.y, 0.f);
                const float ez = fmaxf(fabsf(sp.z - qfz) - hf.z, 0.f);
                need = !(sqrtf(ex * ex + ey * ey + ez * ez) > thr);
            }
            unsigned bal = __ballot_sync(0xffffffffu, need);
            while (bal) {
                // two chunks per step: their index and coordinate gathers are in flight together (each is two
                // dependent L2 round trips, and this search sits on the critical path of a unit's round)
                const int l0 = __ffs(bal) - 1;
                bal &= bal - 1;
                const int l1 = bal ? __ffs(bal) - 1 : -1;
                bal &= bal - 1;
                const int j0 = perm[(size_t)(base + l0) * NN_CH + lane];
                const int j1 = l1 >= 0 ? perm[(size_t)(base + l1) * NN_CH + lane] : -1;
                double a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
                if (j0 >= 0) { a0 = tp[(size_t)j0 * 3]; a1 = tp[(size_t)j0 * 3 + 1]; a2 = tp[(size_t)j0 * 3 + 2]; }
                if (j1 >= 0) { b0 = tp[(size_t)j1 * 3]; b1 = tp[(size_t)j1 * 3 + 1]; b2 = tp[(size_t)j1 * 3 + 2]; }
                if (j0 >= 0) {
                    const double d = sqdist3_rn(x, y, z, a0, a1, a2);
                    if (d < best || (d == best && j0 < bj)) { best = d; bj = j0; }
                }
                if (j1 >= 0) {
                    const double d = sqdist3_rn(x, y, z, b0, b1, b2);
                    if (d < best || (d == best && j1 < bj)) { best = d; bj = j1; }
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (od < best || (od == best && oj < bj)) { best = od; bj = oj; }
    }
}

// one warp per query of the work list written by nn_resolve_kernel
__global__ void __launch_bounds__(128) nn_fix_kernel(Ctx c, int round)
{
    const int count = c.worklist_count[round];
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int w = gw; w < count; w += nw) {
        const int item = c.worklist[w];
        const int u = item / c.n_src;
        const UnitState &st = c.state[u];
        const double *p = c.cur_pts + (size_t)item * 3;
        const float4 qf = c.cur_qf[item];
        const int c1 = c.corr[item];   // nn_resolve_kernel parked the filter's best chunk here
        double best;
        int bj;
        exact_search_warp(c, st.tgt_set, p[0], p[1], p[2], qf.x, qf.y, qf.z, c1, best, bj);
        if (lane == 0) {
            const bool ok = bj != 0x7fffffff && best < c.r2;
            c.corr[item] = ok ? bj : -1;
            c.corr_d2[item] = ok ? best : 0.0;
        }
    }
}

// ----------------------------------------------------------------------------
// dense helpers (float64)
// ----------------------------------------------------------------------------
#define JACOBI_ROT(app, aqq, apq, arp, arq, vp, vq)                                        \
    if (apq != 0.0) {                                                                      \
        const double theta = (aqq - app) / (2.0 * apq);                                    \
        const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0)); \
        const double cc = 1.0 / sqrt(tt * tt + 1.0), ss = tt * cc;                         \
        const double npp = app - tt * apq, nqq = aqq + tt * apq;                           \
        const double nrp = cc * arp - ss * arq, nrq = ss * arp + cc * arq;                 \
        app = npp; aqq = nqq; apq = 0.0; arp = nrp; arq = nrq;                             \
        _Pragma("unroll") for (int r = 0; r < 3; ++r) {                                    \
            const double xx = v[r][vp], yy = v[r][vq];                                     \
            v[r][vp] = cc * xx - ss * yy;                                                  \
            v[r][vq] = ss * xx + cc * yy;                                                  \
        }                                                                                  \
    }

// W = (M^-1)^(1/2) for symmetric positive definite M via cyclic Jacobi
// (restates Eigen's M.inverse().sqrt()).
__device__ __forceinline__ Sym3 inv_sqrt_sym3(const Sym3 &M)
{
    double a00 = M.xx, a01 = M.xy, a02 = M.xz, a11 = M.yy, a12 = M.yz, a22 = M.zz;
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = a01 * a01 + a02 * a02 + a12 * a12;
        const double dg = a00 * a00 + a11 * a11 + a22 * a22;
        if (off <= 1e-36 * dg) break;  // off-diagonals below 1e-18 |A|: no effect in float64
        JACOBI_ROT(a00, a11, a01, a02, a12, 0, 1)
        JACOBI_ROT(a00, a22, a02, a01, a12, 0, 2)
        JACOBI_ROT(a11, a22, a12, a01, a02, 1, 2)
    }
    const double s0 = 1.0 / sqrt(a00), s1 = 1.0 / sqrt(a11), s2 = 1.0 / sqrt(a22);
    Sym3 W;
    W.xx = v[0][0] * s0 * v[0][0] + v[0][1] * s1 * v[0][1] + v[0][2] * s2 * v[0][2];
    W.xy = v[0][0] * s0 * v[1][0] + v[0][1] * s1 * v[1][1] + v[0][2] * s2 * v[1][2];
    W.xz = v[0][0] * s0 * v[2][0] + v[0][1] * s1 * v[2][1] + v[0][2] * s2 * v[2][2];
    W.yy = v[1][0] * s0 * v[1][0] + v[1][1] * s1 * v[1][1] + v[1][2] * s2 * v[1][2];
    W.yz = v[1][0] * s0 * v[2][0] + v[1][1] * s1 * v[2][1] + v[1][2] * s2 * v[2][2];
    W.zz = v[2][0] * s0 * v[2][0] + v[2][1] * s1 * v[2][1] + v[2][2] * s2 * v[2][2];
    return W;
}

// x = A.ldlt().solve(b), 6x6 symmetric, diagonal pivoting (Eigen LDLT restated).
__device__ void ldlt_solve6(double *A, const double *b, double *x)
{
    int perm[6];
    for (int i = 0; i < 6; ++i) perm[i] = i;
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        double best = fabs(A[k * 6 + k]);
        for (int i = k + 1; i < 6; ++i)
            if (fabs(A[i * 6 + i]) > best) { best = fabs(A[i * 6 + i]); piv = i; }
        if (piv != k) {
            for (int j = 0; j < 6; ++j) { const double t = A[k * 6 + j]; A[k * 6 + j] = A[piv * 6 + j]; A[piv * 6 + j] = t; }
            for (int j = 0; j < 6; ++j) { const double t = A[j * 6 + k]; A[j * 6 + k] = A[j * 6 + piv]; A[j * 6 + piv] = t; }
            const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        const double d = A[k * 6 + k];
        if (d == 0.0) continue;
        double col[6];
        for (int i = k + 1; i < 6; ++i) col[i] = A[i * 6 + k];
        for (int i = k + 1; i < 6; ++i) {
            const double l = col[i] / d;
            for (int j = k + 1; j <= i; ++j) A[i * 6 + j] -= l * col[j];
            A[i * 6 + k] = l;
        }
        for (int i = k + 1; i < 6; ++i)
            for (int j = i + 1; j < 6; ++j) A[i * 6 + j] = A[j * 6 + i];
    }
    double y[6];
    for (int i = 0; i < 6; ++i) y[i] = b[perm[i]];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < i; ++j) y[i] -= A[i * 6 + j] * y[j];
    for (int i = 0; i < 6; ++i) y[i] = A[i * 6 + i] != 0.0 ? y[i] / A[i * 6 + i] : 0.0;
    for (int i = 5; i >= 0; --i)
        for (int j = i + 1; j < 6; ++j) y[i] -= A[j * 6 + i] * y[j];
    for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
}

// Eigen::umeyama(src, dst, with_scaling = false) from the one-pass sums
//   tot[0..2] = sum s, tot[3..5] = sum t, tot[6..14] = sum t s^T (row-major), n correspondences:
// sigma = sum t s^T / n - mt ms^T; right singular vectors from the Jacobi eigen-decomposition of
// sigma^T sigma; optimal proper rotation R = [u1 u2 u1xu2][v1 v2 v1xv2]^T; translation mt - R ms.
// (TransformationEstimationPointToPoint::ComputeTransformation, Aligner/Aligner.py:353-360.)
__device__ void umeyama_from_sums(const double *tot, double n, double *Um)
{
    for (int k = 0; k < 12; ++k) Um[k] = (k % 5 == 0) ? 1.0 : 0.0;
    if (!(n > 0)) return;
    const double ms[3] = {tot[0] / n, tot[1] / n, tot[2] / n}, mt[3] = {tot[3] / n, tot[4] / n, tot[5] / n};
    double S[3][3];
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) S[r][cc] = tot[6 + r * 3 + cc] / n - mt[r] * ms[cc];
    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
    for (int k = 0; k < 3; ++k) {
        a00 += S[k][0] * S[k][0]; a01 += S[k][0] * S[k][1]; a02 += S[k][0] * S[k][2];
        a11 += S[k][1] * S[k][1]; a12 += S[k][1] * S[k][2]; a22 += S[k][2] * S[k][2];
    }
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 16; ++sweep) {
        const double off = a01 * a01 + a02 * a02 + a12 * a12;
        if (off == 0.0) break;
        JACOBI_ROT(a00, a11, a01, a02, a12, 0, 1)
        JACOBI_ROT(a00, a22, a02, a01, a12, 0, 2)
        JACOBI_ROT(a11, a22, a12, a01, a02, 1, 2)
    }
    const double w[3] = {a00, a11, a22};
    int o0 = 0, o1 = 1, o2 = 2;
    if (w[o1] > w[o0]) { const int t = o0; o0 = o1; o1 = t; }
    if (w[o2] > w[o0]) { const int t = o0; o0 = o2; o2 = t; }
    if (w[o2] > w[o1]) { const int t = o1; o1 = o2; o2 = t; }
    const double v1[3] = {v[0][o0], v[1][o0], v[2][o0]}, v2[3] = {v[0][o1], v[1][o1], v[2][o1]};
    const double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
    double u1[3], u2[3];
    for (int r = 0; r < 3; ++r) {
        u1[r] = S[r][0] * v1[0] + S[r][1] * v1[1] + S[r][2] * v1[2];
        u2[r] = S[r][0] * v2[0] + S[r][1] * v2[1] + S[r][2] * v2[2];
    }
    const double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    if (!(n1 > 0)) return;
    for (int r = 0; r < 3; ++r) u1[r] /= n1;
    const double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
    for (int r = 0; r < 3; ++r) u2[r] -= d12 * u1[r];
    double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    if (!(n2 > 1e-300)) {
        double e[3] = {0, 0, 0};
        e[fabs(u1[0]) <= fabs(u1[1]) ? (fabs(u1[0]) <= fabs(u1[2]) ? 0 : 2) : (fabs(u1[1]) <= fabs(u1[2]) ? 1 : 2)] = 1.0;
        u2[0] = u1[1] * e[2] - u1[2] * e[1]; u2[1] = u1[2] * e[0] - u1[0] * e[2]; u2[2] = u1[0] * e[1] - u1[1] * e[0];
        n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    }
    for (int r = 0; r < 3; ++r) u2[r] /= n2;
    const double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
    for (int r = 0; r < 3; ++r) {
        for (int cc = 0; cc < 3; ++cc) Um[r * 4 + cc] = u1[r] * v1[cc] + u2[r] * v2[cc] + u3[r] * v3[cc];
        Um[r * 4 + 3] = mt[r] - (Um[r * 4] * ms[0] + Um[r * 4 + 1] * ms[1] + Um[r * 4 + 2] * ms[2]);
    }
}

// ----------------------------------------------------------------------------
// residual / Jacobian reduction + solve
// ----------------------------------------------------------------------------
// One correspondence (source point s with covariance data cs, target point t with ct, squared distance d2)
// added to the 29 running sums: acc[0..20] upper triangle of J^T J (row-major), acc[21..26] J^T r,
// acc[27] sum d2, acc[28] count.  EST: 0 Generalized-ICP with full covariances, 1 Generalized-ICP with plane
// axes, 2 evaluation only, 3 point-to-point (sums for Umeyama), 4 point-to-plane.
//
// EST 1 uses J^T J = Jr^T M^-1 Jr and J^T r = Jr^T M^-1 d (W = M^-1/2 is symmetric, Jr = [-skew(s) | I]) with the
// closed form  M^-1 = I/2 + g+ p p^T + g- m m^T,  p = a + sgn(a.b) b, m = a - sgn(a.b) b, g+- = kappa / (4 l+-),
// l+- = 2 - kappa (1 +- |a.b|)  (M = 2I - kappa (a a^T + b b^T) has eigenvectors p, m).  So
//   Jr^T M^-1 Jr = Jr^T Jr / 2 + g+ e+ e+^T + g- e- e-^T,   e+- = Jr^T (p | m) = (s x p, p) | (s x m, m),
// accumulated as TWICE these sums (finish_round halves them); the identity block of Jr^T Jr is added from the
// count there.  No square root, one division, DFMA issued explicitly.
template <int EST>
__device__ __forceinline__ void accumulate_pair(const Ctx &c, double (&acc)[32], double sx, double sy, double sz,
                                                double tx, double ty, double tz, const double *__restrict__ cs,
                                                const double *__restrict__ ct, double d2)
{
    acc[27] += d2;
    acc[28] += 1.0;
    if (EST == 2) return;
    const double dx = sx - tx, dy = sy - ty, dz = sz - tz;
    if (EST == 3) {
        acc[0] += sx; acc[1] += sy; acc[2] += sz;
        acc[3] += tx; acc[4] += ty; acc[5] += tz;
        acc[6] += tx * sx; acc[7] += tx * sy; acc[8] += tx * sz;
        acc[9] += ty * sx; acc[10] += ty * sy; acc[11] += ty * sz;
        acc[12] += tz * sx; acc[13] += tz * sy; acc[14] += tz * sz;
    } else if (EST == 4) {
        // point-to-plane: r = (vs - vt).nt, J = [vs x nt, nt]
        const double nx = ct[0], ny = ct[1], nz = ct[2];
        const double r = dx * nx + dy * ny + dz * nz;
        const double J[6] = {sy * nz - sz * ny, sz * nx - sx * nz, sx * ny - sy * nx, nx, ny, nz};
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = a; b < 6; ++b) acc[k++] += J[a] * J[b];
#pragma unroll
        for (int a = 0; a < 6; ++a) acc[21 + a] += J[a] * r;
    } else if (EST == 1) {
        const double ax = cs[0], ay = cs[1], az = cs[2], bx = ct[0], by = ct[1], bz = ct[2];
        const double dot = fma(ax, bx, fma(ay, by, az * bz));
        const double sg = dot >= 0 ? 1.0 : -1.0;
        const double ac = fmin(fabs(dot), 1.0);
        const double lp = fma(-c.kappa, 1.0 + ac, 2.0), lm = fma(-c.kappa, 1.0 - ac, 2.0);
        const double inv = (0.5 * c.kappa) / (lp * lm);
        const double gp = inv * lm, gm = inv * lp;                      // 2 g+, 2 g-
        const double px = fma(sg, bx, ax), py = fma(sg, by, ay), pz = fma(sg, bz, az);
        const double mx = fma(-sg, bx, ax), my = fma(-sg, by, ay), mz = fma(-sg, bz, az);
        const double e[6] = {fma(sy, pz, -(sz * py)), fma(sz, px, -(sx * pz)), fma(sx, py, -(sy * px)), px, py, pz};
        const double f[6] = {fma(sy, mz, -(sz * my)), fma(sz, mx, -(sx * mz)), fma(sx, my, -(sy * mx)), mx, my, mz};
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double ga = gp * e[a], ha = gm * f[a];
#pragma unroll
            for (int b = a; b < 6; ++b) { acc[k] = fma(ga, e[b], fma(ha, f[b], acc[k])); ++k; }
        }
        const double al = gp * fma(px, dx, fma(py, dy, pz * dz)), be = gm * fma(mx, dx, fma(my, dy, mz * dz));
#pragma unroll
        for (int a = 0; a < 6; ++a) acc[21 + a] = fma(al, e[a], fma(be, f[a], acc[21 + a]));
        // Jr^T Jr (upper triangle, identity block left to finish_round) and Jr^T d
        acc[0] = fma(sy, sy, fma(sz, sz, acc[0]));
        acc[1] = fma(-sx, sy, acc[1]);
        acc[2] = fma(-sx, sz, acc[2]);
        acc[6] = fma(sx, sx, fma(sz, sz, acc[6]));
        acc[7] = fma(-sy, sz, acc[7]);
        acc[11] = fma(sx, sx, fma(sy, sy, acc[11]));
        acc[4] -= sz; acc[5] += sy;      // (0,4) (0,5)
        acc[8] += sz; acc[10] -= sx;     // (1,3) (1,5)
        acc[12] -= sy; acc[13] += sx;    // (2,3) (2,4)
        acc[21] += fma(sy, dz, -(sz * dy));
        acc[22] += fma(sz, dx, -(sx * dz));
        acc[23] += fma(sx, dy, -(sy * dx));
        acc[24] += dx; acc[25] += dy; acc[26] += dz;
    } else {
        const Sym3 Cs = load_sym3(cs);
        const Sym3 Ct = load_sym3(ct);
        Sym3 M;
        M.xx = Ct.xx + Cs.xx; M.xy = Ct.xy + Cs.xy; M.xz = Ct.xz + Cs.xz;
        M.yy = Ct.yy + Cs.yy; M.yz = Ct.yz + Cs.yz; M.zz = Ct.zz + Cs.zz;
        const Sym3 W = inv_sqrt_sym3(M);
        const double Wm[3][3] = {{W.xx, W.xy, W.xz}, {W.xy, W.yy, W.yz}, {W.xz, W.yz, W.zz}};
        // J = W [ -skew(vs) | I ],  r = W d
        double J[3][6], rr[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            J[r][0] = Wm[r][1] * (-sz) + Wm[r][2] * sy;
            J[r][1] = Wm[r][0] * sz + Wm[r][2] * (-sx);
            J[r][2] = Wm[r][0] * (-sy) + Wm[r][1] * sx;
            J[r][3] = Wm[r][0]; J[r][4] = Wm[r][1]; J[r][5] = Wm[r][2];
            rr[r] = Wm[r][0] * dx + Wm[r][1] * dy + Wm[r][2] * dz;
        }
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = a; b < 6; ++b) acc[k++] += J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b];
#pragma unroll
        for (int a = 0; a < 6; ++a) acc[21 + a] += J[0][a] * rr[0] + J[1][a] * rr[1] + J[2][a] * rr[2];
    }
}

// End of a round for one unit, run by a single thread on the unit's 29 totals: Open3D's stopping rule, the
// 6x6 solve (or Umeyama), T <- U T, results for finished units.
// Returns true when the unit goes on to another round.
template <int EST>
__device__ bool finish_round(const Ctx &c, UnitState &st, int u, double *tot, int round)
{
    const double n = tot[28];
    if (EST == 1) {   // accumulate_pair<1> sums twice J^T J / J^T r without the identity block
        tot[15] += n; tot[18] += n; tot[20] += n;
        for (int k = 0; k < 27; ++k) tot[k] *= 0.5;
    }
    double fitness = 0.0, rmse = 0.0;
    if (n > 0) { fitness = n / (double)c.n_src; rmse = sqrt(tot[27] / n); }
    if (c.trace) {
        double *tr = c.trace + ((size_t)u * (c.max_iter + 1) + round) * 2;
        tr[0] = fitness; tr[1] = rmse;
    }
    // the unit's state was written by whichever warp closed its previous round, possibly on another SM of the same
    // (persistent) launch: read it through L2
    const bool converged = round >= 1 && fabs(__ldcg(&st.fitness) - fitness) < c.rel_fit &&
                           fabs(__ldcg(&st.rmse) - rmse) < c.rel_rmse;
    st.fitness = fitness;
    st.rmse = rmse;
    st.n_corr = (int)n;
    double Told[16];
    for (int k = 0; k < 16; ++k) Told[k] = __ldcg(&st.T[k]);
    if (converged || round >= c.max_iter) {
        st.done = 1;
        st.has_update = 0;
        st.status = (rmse == 0.0) ? 1 : 0;
        for (int k = 0; k < 16; ++k) c.out_T[(size_t)u * 16 + k] = Told[k];
        double *os = c.out_stats + (size_t)u * 4;
        os[0] = fitness; os[1] = rmse; os[2] = (double)__ldcg(&st.iters); os[3] = n;
        c.out_status[u] = st.status;
        if (atomicAdd(c.done_count, 1) == c.n_units - 1) {   // last unit of the batch: release the persistent workers
            __threadfence();
            atomicExch(c.q_ctrl + Q_DONE, 1u);
        }
        return false;
    }
    double Um[12];
    if (EST == 3) {
        umeyama_from_sums(tot, n, Um);
    } else if (n > 0) {
        double A[36], nb[6], x[6];
        int k = 0;
        for (int a = 0; a < 6; ++a)
            for (int b = a; b < 6; ++b) { A[a * 6 + b] = tot[k]; A[b * 6 + a] = tot[k]; ++k; }
        for (int a = 0; a < 6; ++a) nb[a] = -tot[21 + a];
        ldlt_solve6(A, nb, x);
        double sa, ca, sb, cb, sg, cg;
        sincos(x[0], &sa, &ca);
        sincos(x[1], &sb, &cb);
        sincos(x[2], &sg, &cg);
        Um[0] = cg * cb; Um[1] = cg * sb * sa - sg * ca; Um[2] = cg * sb * ca + sg * sa; Um[3] = x[3];
        Um[4] = sg * cb; Um[5] = sg * sb * sa + cg * ca; Um[6] = sg * sb * ca - cg * sa; Um[7] = x[4];
        Um[8] = -sb;     Um[9] = cb * sa;                Um[10] = cb * ca;               Um[11] = x[5];
    } else {
        for (int k = 0; k < 12; ++k) Um[k] = (k % 5 == 0) ? 1.0 : 0.0;
    }
    // T = U * T
    double Tn[16];
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 4; ++cc)
            Tn[r * 4 + cc] = Um[r * 4] * Told[cc] + Um[r * 4 + 1] * Told[4 + cc] + Um[r * 4 + 2] * Told[8 + cc] +
                             Um[r * 4 + 3] * Told[12 + cc];
    Tn[12] = Told[12]; Tn[13] = Told[13]; Tn[14] = Told[14]; Tn[15] = Told[15];
    for (int k = 0; k < 16; ++k) st.T[k] = Tn[k];
    for (int k = 0; k < 12; ++k) st.U[k] = Um[k];
    st.has_update = 1;
    st.iters = round + 1;
    return true;
}

// ----------------------------------------------------------------------------
// The same end of a round run by a whole WARP (the fused kernels): in a persistent launch the closing of a unit's
// round sits on the unit's critical path -- round r+1 cannot start before it -- and the single-thread version spends
// ~19 us in the 6x6 LDLT alone (36-entry arrays in local memory, every access a dependent L1 round trip), a third of
// a straggler's per-round latency.  Here the matrix lives in shared memory, the pivot search is a shuffle arg-max, the
// 15 updates of a step run on 15 lanes, the three sincos on three lanes and the rows of T <- U T on twelve.  Every
// entry sees exactly the operations of ldlt_solve6 / finish_round in the same order: bit-identical results.
// sm: >= 96 doubles of the warp's shared memory that do not overlap tot[0..31].
// ----------------------------------------------------------------------------
template <int EST>
__device__ __forceinline__ bool finish_round_warp(const Ctx &c, UnitState &st, int u, double *tot, double *sm, int round,
                                                  int lane)
{
    double *A = sm, *nb = sm + 36, *x = sm + 42, *Um = sm + 48, *sc = sm + 60, *Told = sm + 66;   // 36+6+6+12+6+16
    int *sperm = reinterpret_cast<int *>(sm + 82);
    int what = 0;        // 0: the unit is done; 1: solve the 6x6 system; 2: Um is ready (Umeyama / identity)
    if (lane < 16) Told[lane] = __ldcg(&st.T[lane]);
    __syncwarp();
    constexpr bool ASSEMBLED = EST == 1 && ORPCD_PLANE_MMA;    // plane_tiles_to_system halved the sums and wrote A, nb
    if (lane == 0) {
        const double n = tot[28];
        if (EST == 1 && !ASSEMBLED) {   // accumulate_pair<1> sums twice J^T J / J^T r without the identity block
            tot[15] += n; tot[18] += n; tot[20] += n;
            for (int k = 0; k < 27; ++k) tot[k] *= 0.5;
        }
        double fitness = 0.0, rmse = 0.0;
        if (n > 0) { fitness = n / (double)c.n_src; rmse = sqrt(tot[27] / n); }
        if (c.trace) {
            double *tr = c.trace + ((size_t)u * (c.max_iter + 1) + round) * 2;
            tr[0] = fitness; tr[1] = rmse;
        }
        const bool converged = round >= 1 && fabs(__ldcg(&st.fitness) - fitness) < c.rel_fit &&
                               fabs(__ldcg(&st.rmse) - rmse) < c.rel_rmse;
        st.fitness = fitness;
        st.rmse = rmse;
        st.n_corr = (int)n;
        if (converged || round >= c.max_iter) {
            st.done = 1;
            st.has_update = 0;
            st.status = (rmse == 0.0) ? 1 : 0;
            for (int k = 0; k < 16; ++k) c.out_T[(size_t)u * 16 + k] = Told[k];
            double *os = c.out_stats + (size_t)u * 4;
            os[0] = fitness; os[1] = rmse; os[2] = (double)__ldcg(&st.iters); os[3] = n;
            c.out_status[u] = st.status;
            if (atomicAdd(c.done_count, 1) == c.n_units - 1) {   // last unit of the batch: release the persistent workers
                __threadfence();
                atomicExch(c.q_ctrl + Q_DONE, 1u);
            }
            what = 0;
        } else if (EST == 3) {
            double U3[12];
            umeyama_from_sums(tot, n, U3);
            for (int k = 0; k < 12; ++k) Um[k] = U3[k];
            what = 2;
        } else if (n > 0) {
            if (!ASSEMBLED) {
                int k = 0;
                for (int a = 0; a < 6; ++a)
                    for (int b = a; b < 6; ++b) { A[a * 6 + b] = tot[k]; A[b * 6 + a] = tot[k]; ++k; }
                for (int a = 0; a < 6; ++a) nb[a] = -tot[21 + a];
            }
            what = 1;
        } else {
            for (int k = 0; k < 12; ++k) Um[k] = (k % 5 == 0) ? 1.0 : 0.0;
            what = 2;
        }
    }
    what = __shfl_sync(0xffffffffu, what, 0);
    if (what == 0) return false;
    __syncwarp();
    if (what == 1) {
        ldlt_solve6_warp(A, nb, x, sperm, lane);
        if (lane < 3) sincos(x[lane], &sc[lane], &sc[3 + lane]);       // sin alpha, beta, gamma | cos alpha, beta, gamma
        __syncwarp();
        if (lane == 0) {
            const double sa = sc[0], sb = sc[1], sg = sc[2], ca = sc[3], cb = sc[4], cg = sc[5];
            Um[0] = cg * cb; Um[1] = cg * sb * sa - sg * ca; Um[2] = cg * sb * ca + sg * sa; Um[3] = x[3];
            Um[4] = sg * cb; Um[5] = sg * sb * sa + cg * ca; Um[6] = sg * sb * ca - cg * sa; Um[7] = x[4];
            Um[8] = -sb;     Um[9] = cb * sa;                Um[10] = cb * ca;               Um[11] = x[5];
        }
    }
    __syncwarp();
    // T = U * T (twelve lanes, one entry each), the pending update for the points, the unit's round counter
    if (lane < 12) {
        const int r = lane >> 2, cc = lane & 3;
        st.T[lane] = Um[r * 4] * Told[cc] + Um[r * 4 + 1] * Told[4 + cc] + Um[r * 4 + 2] * Told[8 + cc] +
                     Um[r * 4 + 3] * Told[12 + cc];
        st.U[lane] = Um[lane];
    }
    if (lane == 0) {
        st.has_update = 1;
        st.iters = round + 1;
    }
    return true;
}

template <int COV_MODE>
__global__ void __launch_bounds__(RED_THREADS) reduce_kernel(Ctx c, int round)
{
    __shared__ double sm[RED_THREADS / 32][RED_STRIDE];
    __shared__ double tot[RED_STRIDE];
    __shared__ int is_last;
    const int u = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    UnitState &st = c.state[u];
    if (st.done) return;
    double acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.0;
    // RED_PER_THREAD correspondences per thread (block-strided, so loads stay coalesced) before the shuffles
#pragma unroll 1
    for (int rep = 0; rep < RED_PER_THREAD; ++rep) {
        const int i = (blockIdx.x * RED_PER_THREAD + rep) * RED_THREADS + tid;
        if (i >= c.n_src) break;
        const size_t oi = (size_t)u * c.n_src + i;
        const int j = c.corr[oi];
        if (j >= 0) {
            const double *ps = c.cur_pts + oi * 3;
            const size_t tj = (size_t)st.tgt_set * c.n_tgt + j;
            const double *pt = c.tgt_pts + tj * 3;
            const int cst = COV_MODE == 0 ? 6 : 3;
            accumulate_pair<COV_MODE>(c, acc, ps[0], ps[1], ps[2], pt[0], pt[1], pt[2], c.cur_cov + oi * cst,
                                      c.tgt_cov + tj * cst, c.corr_d2[oi]);
        }
    }
#pragma unroll
    for (int k = 0; k < RED_VALS; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < RED_VALS; ++k) sm[wid][k] = acc[k];
    }
    __syncthreads();
    if (tid < RED_VALS) {
        double s = sm[0][tid];
#pragma unroll
        for (int w = 1; w < RED_THREADS / 32; ++w) s += sm[w][tid];
        c.red_partial[((size_t)u * c.red_blocks + blockIdx.x) * RED_STRIDE + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&c.red_counter[(size_t)u * (c.red_groups + 1)], 1u);
        is_last = (prev == (unsigned)gridDim.x - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid < RED_VALS) {
        double s = 0.0;
        for (int b = 0; b < (int)gridDim.x; ++b)
            s += __ldcg(&c.red_partial[((size_t)u * c.red_blocks + b) * RED_STRIDE + tid]);
        tot[tid] = s;
    }
    __syncthreads();
    if (tid != 0) return;
    c.red_counter[(size_t)u * (c.red_groups + 1)] = 0;
    finish_round<COV_MODE>(c, st, u, tot, round);
}

// ----------------------------------------------------------------------------
// One Gauss-Newton round of every unit in ONE launch (the default path): each warp owns a patch of 32*Q
// spatially consecutive source points and runs, for them,
//   1. the pending rigid update (update_kernel's arithmetic) and the bound d_ub from the previous match,
//   2. the chunk-culled FP32 sweep (cull_sweep_core),
//   3. the certificate of nn_resolve_kernel with the warp-wide ballot over the best chunk,
//   4. the exact float64 search for the few uncertified queries (exact_search_warp),
//   5. the residual / Jacobian sums (accumulate_pair), a transposing butterfly that leaves total k in lane k,
//      one 256-byte partial per warp; the last warp of the unit to finish adds the partials in warp order
//      (fixed order: run-to-run deterministic) and closes the round (finish_round).
// The float64 phases of one warp overlap the FP32 sweep of the others on the same SM.
// ----------------------------------------------------------------------------
// after the call lane L holds the warp total of v[L]
__device__ __forceinline__ double warp_transpose_sum32(double (&v)[32], int lane)
{
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double send = up ? v[i] : v[i + half];
            const double keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

// D += A (8x4) B (4x8) in float64 on the tensor cores: lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4) + {0,1}]
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// record v (8 doubles: six vector entries, the residual projection, the weight) of the warp's record buffer.  The
// four 16-byte pieces of a record are stored at piece index (q ^ ((v >> 1) & 3)): eight lanes writing the same piece
// of eight consecutive records then hit eight different bank groups, and the reads of a tensor-core step (four
// consecutive records, all 32 doubles) stay conflict free.
__device__ __forceinline__ void store_record(double *rec, int v, double x0, double x1, double x2, double x3, double x4,
                                             double x5, double x6, double x7)
{
    double2 *r = reinterpret_cast<double2 *>(rec + v * 8);
    const int sw = (v >> 1) & 3;
    r[0 ^ sw] = make_double2(x0, x1);
    r[1 ^ sw] = make_double2(x2, x3);
    r[2 ^ sw] = make_double2(x4, x5);
    r[3 ^ sw] = make_double2(x6, x7);
}

// The totals of the two 8x8 tiles of the tensor-core path (EST 1) are
//   EF = sum_pairs  g+ (e|re)(e|re)^T + g- (f|rf)(f|rf)^T   (rows / cols 0..5 = J^T J part, col 6 = J^T r part),
//   S  = sum_pairs  z z^T,  z = (sx, sy, sz, 1, dx, dy, dz),
// t[63] (the unused EF(7,7) slot) carries sum d2; "twice the sums, identity block left out" as in accumulate_pair<1>.
// Their assembly by a whole warp, one entry per lane, straight into the system finish_round_warp solves: lane k
// (k < 29) computes sum k (the 21 + 6 + 2 sums of accumulate_pair), lanes 0..26 also the halved entry ("twice the sums, identity block
// left out": + n on the translation diagonal, x 0.5) and write it to the 6x6 matrix A (both triangles) or to the
// right-hand side nb = -J^T r.  sm = the buffer of finish_round_warp (A at 0, nb at 36).
__device__ __forceinline__ void plane_tiles_to_system(const double *t, double *o, double *sm, int lane)
{
    const double *S = t + 64;
    const int k = lane;
    double v = 0.0;
    int a = 0, b = 0;
    if (k < 21) {
        // (a, b) of upper-triangle entry k, row-major
        int r = k;
        a = 0;
        while (r >= 6 - a) { r -= 6 - a; ++a; }
        b = a + r;
        v = t[a * 8 + b];
        const double Sxx = S[0], Sxy = S[1], Sxz = S[2], Syy = S[9], Syz = S[10], Szz = S[18];
        const double sx = S[3], sy = S[11], sz = S[19];
        switch (k) {
        case 0: v += Syy + Szz; break;
        case 1: v -= Sxy; break;
        case 2: v -= Sxz; break;
        case 4: v -= sz; break;
        case 5: v += sy; break;
        case 6: v += Sxx + Szz; break;
        case 7: v -= Syz; break;
        case 8: v += sz; break;
        case 10: v -= sx; break;
        case 11: v += Sxx + Syy; break;
        case 12: v -= sy; break;
        case 13: v += sx; break;
        default: break;
        }
    } else if (k < 27) {
        a = k - 21;
        v = t[a * 8 + 6];
        switch (a) {
        case 0: v += S[8 + 6] - S[16 + 5]; break;     // sum (sy dz - sz dy)
        case 1: v += S[16 + 4] - S[6]; break;         // sum (sz dx - sx dz)
        case 2: v += S[5] - S[8 + 4]; break;          // sum (sx dy - sy dx)
        default: v += S[24 + 4 + (a - 3)]; break;     // sum d
        }
    } else if (k == 27) {
        v = t[63];
    } else if (k == 28) {
        v = S[24 + 3];
    }
    const double n = S[24 + 3];
    if (k < 27) {
        if (k == 15 || k == 18 || k == 20) v += n;
        v *= 0.5;
    }
    o[k] = v;                       // the sums as finish_round_warp's lane 0 reads them (already halved)
    if (k < 21) { sm[a * 6 + b] = v; sm[b * 6 + a] = v; }
    else if (k < 27) sm[36 + a] = -v;
}

// items per group of the two-level totals: the multiple of 8 next to sqrt(items), so that both levels read about
// the same number of partials (8 for up to 64 items, 16 up to 256, ..., 56 for the 3125 patches of a 200k-point cloud)
__host__ __device__ inline int red_group_size(int n_items)
{
    int g = 8;
    while (g * g < n_items) g += 8;
    return g;
}

// PERSIST: the caller is the persistent kernel -- one launch for all rounds, so data another SM wrote in an earlier
// round of the same launch (unit state, moving points / axes, correspondences) is read with ld.global.cg (L2, the
// coherence point; an L1 line could be stale), and the warp that closes a unit's round puts the patches of its next
// round on the work queue.
//
// Patch size.  Rounds below c.half_from_round work on 64-point patches (two queries per lane); from that round on a
// unit's items are 32-point patches (one query per lane, `half`): a unit cannot overlap its own rounds, so late in a
// batch -- when the units that run to the iteration cap are all that is left and workers wait on an empty ring -- the
// time is set by the latency of a unit's round, and a 32-point patch has half the sweep per warp and one pass
// through the per-query phases instead of two.  (Smaller patches cost more list building per query, which is why
// the loaded first rounds keep 64.)  The switch depends on the round index only, so results are deterministic and
// independent of which other units share the batch.
template <int Q, int EST, bool PERSIST>
__device__ __forceinline__ void round_body(const Ctx &c, UnitState &st, int u, int wi, int round, WarpSmem &wsm)
{
    float4 (*ring)[NN_CH] = wsm.ring;
    unsigned short *vl = wsm.vl;
    const int lane = threadIdx.x & 31;
    const bool half = round >= c.half_from_round;
    const int nk = half ? 1 : 2;                    // queries per lane
    const int wbase = wi * (half ? 32 : 64);
    if (wbase >= c.n_src) return;
    const int tgt_set = __ldcg(&st.tgt_set);   // (immutable after init_units_kernel; .cg keeps the state's lines out of L1)
    if (wi == 0 && lane == 0) atomicAdd(&c.round_active[round], 1);
    const TgtMeta tm = c.tgt_meta[tgt_set];
    const double *tpts = c.tgt_pts + (size_t)tgt_set * c.n_tgt * 3;
    const float4 *tg = c.tgt4 + (size_t)tgt_set * c.nt_pad;
    constexpr int CST = EST == 0 ? 6 : 3;

    // The two queries of a lane (k = 0, 1) go through phases 1, 3, 4 and 5 in ROLLED loops: the state of the query
    // in hand lives in the "a" variables, the other one's in "b", and they trade places at the end of each pass.
    // Unrolled, the straight-line code of a round was 38 KB per warp with 24 warps per SM in different phases: ncu
    // showed the SM's instruction cache at 87 % hits and the GPC-level instruction cache at 79 % of its request
    // peak over the whole launch -- the kernel was bound by instruction fetch, not by any execution pipe.
    static_assert(Q == 2, "round_body rolls its per-query phases over exactly two queries per lane");
#define ORPCD_SWAP(T, x, y) { const T t__ = x; x = y; y = t__; }

    // ---- 1. pending update, d_ub, centred FP32 copy
    float4 qa = make_float4(0.f, 0.f, 0.f, 0.f), qb = qa;
    {
        const bool has_update = __ldcg(&st.has_update) != 0;
        double U[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) U[k] = __ldcg(&st.U[k]);
#if ORPCD_PH1_HOIST
#pragma unroll 1
        for (int k = 0; k < nk; ++k) {
            const int i = wbase + k * 32 + lane;
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < c.n_src) {
                const size_t oi = (size_t)u * c.n_src + i;
                double *p = c.cur_pts + oi * 3;
                // every load of the pass is issued before its first store (the compiler cannot move a load above a
                // store that might alias it): point, axis and previous match together, then the matched target's point --
                // two L2 round trips per pass instead of four (point | axis | match | target point, each behind stores)
                const int j = round >= 1 ? __ldcg(c.corr + oi) : -1;
                double x = __ldcg(p), y = __ldcg(p + 1), z = __ldcg(p + 2);
                double ax = 0.0, ay = 0.0, az = 0.0;
                if (EST == 1) {
                    const double *ap = c.cur_cov + oi * 3;
                    ax = __ldcg(ap); ay = __ldcg(ap + 1); az = __ldcg(ap + 2);
                }
                double tx = 0.0, ty = 0.0, tz = 0.0;
                if (j >= 0) {
                    const double *tp = tpts + (size_t)j * 3;
                    tx = tp[0]; ty = tp[1]; tz = tp[2];
                }
                if (has_update) {
                    transform_point(U, x, y, z);
                    p[0] = x; p[1] = y; p[2] = z;
                    if (EST == 0) {
                        double *cp = c.cur_cov + oi * 6;
                        Sym3 C;
                        C.xx = __ldcg(cp); C.xy = __ldcg(cp + 1); C.xz = __ldcg(cp + 2);
                        C.yy = __ldcg(cp + 3); C.yz = __ldcg(cp + 4); C.zz = __ldcg(cp + 5);
                        store_sym3(cp, rotate_sym3(U, C));
                    } else if (EST == 1) {
                        double *ap = c.cur_cov + oi * 3;
                        ap[0] = U[0] * ax + U[1] * ay + U[2] * az;
                        ap[1] = U[4] * ax + U[5] * ay + U[6] * az;
                        ap[2] = U[8] * ax + U[9] * ay + U[10] * az;
                    }
                }
                float dub = c.dcap;
                if (j >= 0) {
                    const double dx = x - tx, dy = y - ty, dz = z - tz;
                    dub = fminf(dub, __double2float_ru(sqrt(dx * dx + dy * dy + dz * dz)) * (1.f + 1e-6f));
                }
                r = make_float4((float)(x - tm.cx), (float)(y - tm.cy), (float)(z - tm.cz), dub);
            }
            if (k == 0) qa = r;
            else qb = r;
        }
#else
#pragma unroll 1
        for (int k = 0; k < nk; ++k) {
            const int i = wbase + k * 32 + lane;
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < c.n_src) {
                const size_t oi = (size_t)u * c.n_src + i;
                double *p = c.cur_pts + oi * 3;
                double x = __ldcg(p), y = __ldcg(p + 1), z = __ldcg(p + 2);
                if (has_update) {
                    transform_point(U, x, y, z);
                    p[0] = x; p[1] = y; p[2] = z;
                    if (EST == 0) {
                        double *cp = c.cur_cov + oi * 6;
                        Sym3 C;
                        C.xx = __ldcg(cp); C.xy = __ldcg(cp + 1); C.xz = __ldcg(cp + 2);
                        C.yy = __ldcg(cp + 3); C.yz = __ldcg(cp + 4); C.zz = __ldcg(cp + 5);
                        store_sym3(cp, rotate_sym3(U, C));
                    } else if (EST == 1) {
                        double *ap = c.cur_cov + oi * 3;
                        const double ax = __ldcg(ap), ay = __ldcg(ap + 1), az = __ldcg(ap + 2);
                        ap[0] = U[0] * ax + U[1] * ay + U[2] * az;
                        ap[1] = U[4] * ax + U[5] * ay + U[6] * az;
                        ap[2] = U[8] * ax + U[9] * ay + U[10] * az;
                    }
                }
                float dub = c.dcap;
                if (round >= 1) {
                    const int j = __ldcg(c.corr + oi);
                    if (j >= 0) {
                        const double *tp = tpts + (size_t)j * 3;
                        const double dx = x - tp[0], dy = y - tp[1], dz = z - tp[2];
                        dub = fminf(dub, __double2float_ru(sqrt(dx * dx + dy * dy + dz * dz)) * (1.f + 1e-6f));
                    }
                }
                r = make_float4((float)(x - tm.cx), (float)(y - tm.cy), (float)(z - tm.cz), dub);
            }
            if (k == 0) qa = r;
            else qb = r;
        }
#endif
        // lanes past the end of the cloud carry the patch's first query (always a real one)
        const float fx = __shfl_sync(0xffffffffu, qa.x, 0), fy = __shfl_sync(0xffffffffu, qa.y, 0);
        const float fz = __shfl_sync(0xffffffffu, qa.z, 0), fw = __shfl_sync(0xffffffffu, qa.w, 0);
        if (wbase + lane >= c.n_src) qa = make_float4(fx, fy, fz, fw);
        if (half || wbase + 32 + lane >= c.n_src) qb = make_float4(fx, fy, fz, fw);
    }

    // ---- 2. sweep (ONE call site: the set-up and the visit list of cull_sweep_core exist once in the kernel)
    float m1a, m1b, m2a, m2b;
    int c1a, c1b;
    {
        float4 qv[Q] = {qa, qb};
        if (half) {
            // both half-warps take the pair (query l & 15, query (l & 15) + 16) and one parity of every chunk's targets
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int from = (lane & 15) + 16 * e;
                qv[e].x = __shfl_sync(0xffffffffu, qa.x, from); qv[e].y = __shfl_sync(0xffffffffu, qa.y, from);
                qv[e].z = __shfl_sync(0xffffffffu, qa.z, from); qv[e].w = __shfl_sync(0xffffffffu, qa.w, from);
            }
        }
        float m1[Q], m2[Q];
        int c1[Q];
        const int visited = cull_sweep_core<Q, true>(c, tg, c.chunk_sph + (size_t)tgt_set * c.n_chunks,
                                                     c.super_sph + (size_t)tgt_set * c.n_super, c.n_super, (float)tm.rmax, 0,
                                                     c.n_chunks, qv, ring, vl, m1, c1, m2, half);
        if (lane == 0)
            atomicAdd(c.visit_count + round, (unsigned long long)visited * NN_CH * min(half ? 32 : 64, c.n_src - wbase));
        if (!half) {
            m1a = m1[0]; m1b = m1[1]; m2a = m2[0]; m2b = m2[1]; c1a = c1[0]; c1b = c1[1];
        } else {
            // merge the two halves (each saw the targets of one parity); on an exact tie both keep parity 0's sub-chunk --
            // the certificate then sees m2 == m1 and sends the query to the exact search
            const int par = lane >> 4;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float om1 = __shfl_xor_sync(0xffffffffu, m1[e], 16), om2 = __shfl_xor_sync(0xffffffffu, m2[e], 16);
                const int oc1 = __shfl_xor_sync(0xffffffffu, c1[e], 16);
                const bool take = om1 < m1[e] || (om1 == m1[e] && par == 1);
                m2[e] = fminf(fminf(m2[e], om2), fmaxf(m1[e], om1));
                c1[e] = take ? oc1 : c1[e];
                m1[e] = fminf(m1[e], om1);
            }
            // lane l's own query is element l >> 4 of its pair
            m1a = par ? m1[1] : m1[0]; m2a = par ? m2[1] : m2[0]; c1a = par ? c1[1] : c1[0];
            m1b = CUDART_INF_F; m2b = CUDART_INF_F; c1b = 0;
        }
    }

    // ---- 3. certificate (nn_resolve_kernel's bound), then each lane scans the best 8-target sub-chunk of its own queries:
    //         the FP32 values are recomputed with the sweep's FMA chain (bit-identical), and the chunk decides
    //         the query iff exactly one target lies within the certified margin of the minimum
    int tja = -1, tjb = -1;   // matched target (original index); -1 none within the cut-off; -2 exact search needed
    const int *tperm = c.tperm + (size_t)tgt_set * c.nt_pad;
#pragma unroll 1
    for (int k = 0; k < nk; ++k) {
        const float fx = qa.x, fy = qa.y, fz = qa.z;
        const double qn = (double)fx * fx + (double)fy * fy + (double)fz * fz;
        // |q| and d_ub only enter the error BOUND below: any upper bound serves, so the two square roots are single
        // FP32 instructions rounded up (of arguments rounded up) instead of two float64 sqrt sequences per query
        const double qnorm = (double)__fsqrt_ru(__double2float_ru(qn));
        const double rho_max = qnorm + tm.rmax;
        const double e_glob = (double)F32_EPS * 6.0 * rho_max * rho_max * 1.001;
        const double d_ub = (double)__fsqrt_ru(__double2float_ru(fmax((double)m1a + qn, 0.0) + e_glob)) * (1.0 + 1e-6);
        const double rho = 2.0 * qnorm + d_ub + 4.0 * (double)F32_EPS * rho_max;
        const double margin = 2.02 * (double)F32_EPS * (4.0 * rho * rho + 2.0 * d_ub * rho) * 1.001 + 1e-300;
        const double lim = (double)m1a + margin;
        const bool none = !(m1a < CUDART_INF_F);          // no chunk visited: nothing within the cut-off
        const bool cross = !((double)m2a > lim);          // another chunk may hold the true neighbour
        const float thr = __double2float_ru(lim);
        // c1 = 8-target sub-chunk: 8 consecutive targets, or (32-point patches) 8 targets of one parity, stride 2
        const size_t sub_base = half ? (size_t)(c1a >> 2) * 32 + (size_t)(c1a & 1) * 16 + (size_t)((c1a >> 1) & 1)
                                     : (size_t)c1a * 8;
        const int sub_stride = half ? 2 : 1;
        const float4 *tc = tg + sub_base;
        int cnt = 0, first = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(tc + j * sub_stride);
            const float sv = __fmaf_rn(fx, t.x, __fmaf_rn(fy, t.y, __fmaf_rn(fz, t.z, t.w)));
            const bool in = sv <= thr;
            first = (in && cnt == 0) ? j : first;
            cnt += in ? 1 : 0;
        }
        int t = none ? -1 : ((cross || cnt != 1) ? -2 : tperm[sub_base + (size_t)first * sub_stride]);
        if (wbase + k * 32 + lane >= c.n_src) t = -1;
        tja = t;
        if (nk == 2) {     // trade places: after the two passes the "a" set is query 0 again
            ORPCD_SWAP(float4, qa, qb) ORPCD_SWAP(float, m1a, m1b) ORPCD_SWAP(float, m2a, m2b) ORPCD_SWAP(int, c1a, c1b)
            ORPCD_SWAP(int, tja, tjb)
        }
    }

    // ---- 4. exact search for what the FP32 filter could not decide
    double fda = -1.0, fdb = -1.0;
    int n_fix = 0;
#pragma unroll 1
    for (int k = 0; k < nk; ++k) {
        unsigned fb = __ballot_sync(0xffffffffu, tja == -2);
        n_fix += __popc(fb);
        while (fb) {
            const int l = __ffs(fb) - 1;
            fb &= fb - 1;
            const double *p = c.cur_pts + ((size_t)u * c.n_src + wbase + k * 32 + l) * 3;
            const float qfx = __shfl_sync(0xffffffffu, qa.x, l), qfy = __shfl_sync(0xffffffffu, qa.y, l);
            const float qfz = __shfl_sync(0xffffffffu, qa.z, l);
            double best;
            int bj;
            exact_search_warp(c, tgt_set, __ldcg(p), __ldcg(p + 1), __ldcg(p + 2), qfx, qfy, qfz,
                              __shfl_sync(0xffffffffu, c1a, l) >> 2, best, bj);
            if (lane == l) { tja = bj == 0x7fffffff ? -1 : bj; fda = best; }
        }
        if (nk == 2) {
            ORPCD_SWAP(int, tja, tjb) ORPCD_SWAP(double, fda, fdb) ORPCD_SWAP(float4, qa, qb) ORPCD_SWAP(int, c1a, c1b)
        }
    }
    if (lane == 0 && n_fix) atomicAdd(&c.worklist_count[round], n_fix);

    // ---- 5. sums over this warp's correspondences
    const int n_warps = half ? c.n_patches2 : c.n_patches;     // work items of this unit's round
    double *tot = reinterpret_cast<double *>(&wsm);   // the sweep's ring and visit list are free now
    if constexpr (EST == 1 && ORPCD_PLANE_MMA) {
        // Plane-to-plane Generalized-ICP on the float64 tensor cores.  Per correspondence three weighted 8-vectors
        //   (e | p.d ; 2g+)   (f | m.d ; 2g-)   (sx, sy, sz, 1, dx, dy, dz ; 1)
        // go through the warp's record buffer (see accumulate_pair<1> for e, f, g+-); sum_v w v v^T over the 32 pairs
        // of a pass is 24 mma.m8n8k4.f64 steps (A = w v, B = v: lane l reads entry l/4 of record 4 s + l%4):
        // 16 on the (e, f) records into one 8x8 tile, then 8 on the (s, d) records into a second one -- 27 + 16
        // needed sums for 4 registers of accumulators.
        double *rec = reinterpret_cast<double *>(&wsm);
        double ef0 = 0.0, ef1 = 0.0, sd0 = 0.0, sd1 = 0.0, d2sum = 0.0;
        const int jv = lane & 3, cc = lane >> 2;
        // offsets (in doubles) of this lane's entry and of the weight inside record 4 s + jv, for even / odd s
        const int sw_e = jv >> 1, sw_o = 2 + (jv >> 1);
        const int x_e = jv * 8 + ((((cc >> 1) ^ sw_e) << 1) | (cc & 1)), w_e = jv * 8 + (((3 ^ sw_e) << 1) | 1);
        const int x_o = jv * 8 + ((((cc >> 1) ^ sw_o) << 1) | (cc & 1)), w_o = jv * 8 + (((3 ^ sw_o) << 1) | 1);
#pragma unroll 1
        for (int k = 0; k < nk; ++k) {
            const int i = wbase + k * 32 + lane;
            bool have = false;
            double sx = 0, sy = 0, sz = 0, dx = 0, dy = 0, dz = 0;
            if (i < c.n_src) {
                const size_t oi = (size_t)u * c.n_src + i;
                int j = tja;
                if (j >= 0) {
                    const double *ps = c.cur_pts + oi * 3;
                    const double *pt = tpts + (size_t)j * 3;
                    sx = __ldcg(ps); sy = __ldcg(ps + 1); sz = __ldcg(ps + 2);
                    const double tx = pt[0], ty = pt[1], tz = pt[2];
#if ORPCD_PH5_HOIST
                    // both axes are requested with the points (a match beyond the cut-off is the rare case)
                    const double *as = c.cur_cov + oi * 3;   // moving source axis: written this round, read through L2
                    const double *bt = c.tgt_cov + ((size_t)tgt_set * c.n_tgt + j) * 3;
                    const double ax = __ldcg(as), ay = __ldcg(as + 1), az = __ldcg(as + 2);
                    const double bx = bt[0], by = bt[1], bz = bt[2];
#endif
                    const double d2 = fda >= 0.0 ? fda : sqdist3_rn(sx, sy, sz, tx, ty, tz);
                    if (d2 < c.r2) {
                        have = true;
                        d2sum += d2;
                        dx = sx - tx; dy = sy - ty; dz = sz - tz;
#if !ORPCD_PH5_HOIST
                        const double *as = c.cur_cov + oi * 3;   // moving source axis: written this round, read through L2
                        const double *bt = c.tgt_cov + ((size_t)tgt_set * c.n_tgt + j) * 3;
                        const double ax = __ldcg(as), ay = __ldcg(as + 1), az = __ldcg(as + 2);
                        const double bx = bt[0], by = bt[1], bz = bt[2];
#endif
                        const double dot = fma(ax, bx, fma(ay, by, az * bz));
                        const double sg = dot >= 0 ? 1.0 : -1.0;
                        const double ac = fmin(fabs(dot), 1.0);
                        const double lp = fma(-c.kappa, 1.0 + ac, 2.0), lm = fma(-c.kappa, 1.0 - ac, 2.0);
                        const double inv = (0.5 * c.kappa) / (lp * lm);
                        {
                            const double px = fma(sg, bx, ax), py = fma(sg, by, ay), pz = fma(sg, bz, az);
                            store_record(rec, lane, fma(sy, pz, -(sz * py)), fma(sz, px, -(sx * pz)), fma(sx, py, -(sy * px)),
                                         px, py, pz, fma(px, dx, fma(py, dy, pz * dz)), inv * lm);       // weight 2 g+
                        }
                        {
                            const double mx = fma(-sg, bx, ax), my = fma(-sg, by, ay), mz = fma(-sg, bz, az);
                            store_record(rec, 32 + lane, fma(sy, mz, -(sz * my)), fma(sz, mx, -(sx * mz)),
                                         fma(sx, my, -(sy * mx)), mx, my, mz, fma(mx, dx, fma(my, dy, mz * dz)),
                                         inv * lp);                                                      // weight 2 g-
                        }
                    } else {
                        j = -1;
                    }
                }
                c.corr[oi] = j;
            }
            if (!have) {
                store_record(rec, lane, 0, 0, 0, 0, 0, 0, 0, 0);
                store_record(rec, 32 + lane, 0, 0, 0, 0, 0, 0, 0, 0);
            }
            __syncwarp();
#pragma unroll 2
            for (int sstep = 0; sstep < 16; ++sstep) {
                const double *r4 = rec + sstep * 32;
                const double x = r4[(sstep & 1) ? x_o : x_e], w = r4[(sstep & 1) ? w_o : w_e];
                dmma884(ef0, ef1, w * x, x);
            }
            __syncwarp();
            // the (s, d) records of the same 32 pairs reuse the first half of the buffer
            if (have) store_record(rec, lane, sx, sy, sz, 1.0, dx, dy, dz, 1.0);
            else store_record(rec, lane, 0, 0, 0, 0, 0, 0, 0, 0);
            __syncwarp();
#pragma unroll 2
            for (int sstep = 0; sstep < 8; ++sstep) {
                const double *r4 = rec + sstep * 32;
                const double x = r4[(sstep & 1) ? x_o : x_e], w = r4[(sstep & 1) ? w_o : w_e];
                dmma884(sd0, sd1, w * x, x);
            }
            __syncwarp();
            if (nk == 2) { ORPCD_SWAP(int, tja, tjb) ORPCD_SWAP(double, fda, fdb) }
        }
        d2sum = warp_sum(d2sum);
        if (lane == 31) ef1 = d2sum;      // EF(7,7) is not a needed sum: its slot carries sum d2
        double *part = c.red_partial + ((size_t)u * c.red_blocks + wi) * RED_STRIDE_MMA;
        __stcg(reinterpret_cast<double2 *>(part) + lane, make_double2(ef0, ef1));
        __stcg(reinterpret_cast<double2 *>(part + 64) + lane, make_double2(sd0, sd1));
    } else {
        double acc[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = 0.0;
#pragma unroll 1
        for (int k = 0; k < nk; ++k) {
            const int i = wbase + k * 32 + lane;
            if (i < c.n_src) {
                const size_t oi = (size_t)u * c.n_src + i;
                int j = tja;
                if (j >= 0) {
                    const double *ps = c.cur_pts + oi * 3;
                    const double *pt = tpts + (size_t)j * 3;
                    const double sx = __ldcg(ps), sy = __ldcg(ps + 1), sz = __ldcg(ps + 2), tx = pt[0], ty = pt[1], tz = pt[2];
                    const double d2 = fda >= 0.0 ? fda : sqdist3_rn(sx, sy, sz, tx, ty, tz);
                    if (d2 < c.r2) {
                        double cs[CST];   // the moving source's covariance data: written this round, read through L2
                        if (EST <= 1) {
#pragma unroll
                            for (int e = 0; e < CST; ++e) cs[e] = __ldcg(c.cur_cov + oi * CST + e);
                        }
                        accumulate_pair<EST>(c, acc, sx, sy, sz, tx, ty, tz, cs,
                                             c.tgt_cov + ((size_t)tgt_set * c.n_tgt + j) * CST, d2);
                    } else {
                        j = -1;
                    }
                }
                c.corr[oi] = j;
            }
            if (nk == 2) { ORPCD_SWAP(int, tja, tjb) ORPCD_SWAP(double, fda, fdb) }
        }
        const double mine = warp_transpose_sum32(acc, lane);
        __stcg(c.red_partial + ((size_t)u * c.red_blocks + wi) * RED_STRIDE_MMA + lane, mine);
    }
#undef ORPCD_SWAP
    // Two-level, fixed-order totals.  The items of a round are grouped (red_group_size: 16 for the 79 / 157 patches
    // of a 5000-point cloud); the last warp to arrive in a group adds the group's partials in item order into a group
    // partial, and the last group to close adds the group partials in group order.  The warp that closes the round
    // therefore waits for two or three batches of loads instead of ten (a single warp reading 79 x 1 KB took ~10 us,
    // a sixth of a straggler's round); the order of the additions is fixed: results stay run-to-run deterministic.
    constexpr bool MMA = EST == 1 && ORPCD_PLANE_MMA;
    const int gsz = red_group_size(n_warps);
    const int n_groups = (n_warps + gsz - 1) / gsz;
    const int grp = wi / gsz;
    unsigned *cnt = c.red_counter + (size_t)u * (c.red_groups + 1);
    double *part_u = c.red_partial + (size_t)u * c.red_blocks * RED_STRIDE_MMA;
    double *gpart_u = part_u + (size_t)c.n_patches2 * RED_STRIDE_MMA;  // group partials follow the item partials
    fence_release_side();
    __syncwarp();
    int last = 0;
    if (lane == 0) last = atom_add_release(cnt + 1 + grp, 1u) == (unsigned)min(gsz, n_warps - grp * gsz) - 1u;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    fence_acquire_side();
    {
        const int b0 = grp * gsz, b1 = min(n_warps, b0 + gsz);
        double *gp = gpart_u + (size_t)grp * RED_STRIDE_MMA;
        if constexpr (MMA) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 8
            for (int b = b0; b < b1; ++b) {
                const double2 *pp = reinterpret_cast<const double2 *>(part_u + (size_t)b * RED_STRIDE_MMA) + lane * 2;
                const double2 v0 = __ldcg(pp), v1 = __ldcg(pp + 1);
                a0 += v0.x; a1 += v0.y; a2 += v1.x; a3 += v1.y;
            }
            __stcg(reinterpret_cast<double2 *>(gp) + lane * 2, make_double2(a0, a1));
            __stcg(reinterpret_cast<double2 *>(gp) + lane * 2 + 1, make_double2(a2, a3));
        } else {
            double a0 = 0.0;
#pragma unroll 8
            for (int b = b0; b < b1; ++b) a0 += __ldcg(part_u + (size_t)b * RED_STRIDE_MMA + lane);
            __stcg(gp + lane, a0);
        }
    }
    fence_release_side();
    __syncwarp();
    last = 0;
    if (lane == 0) {
        cnt[1 + grp] = 0u;                                   // ready for the unit's next round
        last = atom_add_release(cnt, 1u) == (unsigned)n_groups - 1u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
#ifdef ORPCD_QUEUE_STATS
    unsigned long long tc0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tc0));
#endif
    fence_acquire_side();
#ifdef ORPCD_QUEUE_STATS
    {   // span of this unit's round: from its publication (stamp kept in the unused work list) to its last arrival
        unsigned long long *pub = reinterpret_cast<unsigned long long *>(c.worklist) + u;
        const unsigned long long t0 = *reinterpret_cast<volatile unsigned long long *>(c.q_stats);
        const unsigned long long tpub = __ldcg(pub);
        if (lane == 0 && tpub != 0ull) {
            const unsigned b = (unsigned)min((unsigned long long)63, (tc0 - t0) / 250000ull);
            atomicAdd(c.q_stats + 448 + b, tc0 - tpub);
        }
    }
#endif
    // unit totals: the group partials added in group order
    if constexpr (MMA) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 8
        for (int b = 0; b < n_groups; ++b) {
            const double2 *pp = reinterpret_cast<const double2 *>(gpart_u + (size_t)b * RED_STRIDE_MMA) + lane * 2;
            const double2 v0 = __ldcg(pp), v1 = __ldcg(pp + 1);
            a0 += v0.x; a1 += v0.y; a2 += v1.x; a3 += v1.y;
        }
        double *t128 = tot + 32;
        t128[lane * 4] = a0; t128[lane * 4 + 1] = a1; t128[lane * 4 + 2] = a2; t128[lane * 4 + 3] = a3;
        __syncwarp();
        plane_tiles_to_system(t128, tot, tot + 160, lane);   // the 29 sums AND the halved 6x6 system, one entry per lane
    } else {
        double s0 = 0.0;
        for (int b = 0; b < n_groups; ++b) s0 += __ldcg(gpart_u + (size_t)b * RED_STRIDE_MMA + lane);
        tot[lane] = s0;
    }
    __syncwarp();
#ifdef ORPCD_QUEUE_STATS
    unsigned long long tc_tot;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tc_tot));
#endif
    if (lane == 0) cnt[0] = 0u;
    const bool go_on = finish_round_warp<EST>(c, st, u, tot, tot + 160, round, lane);
#ifdef ORPCD_QUEUE_STATS
    unsigned long long tc_fin;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tc_fin));
    if (lane == 0) {
        const unsigned long long t0 = *reinterpret_cast<volatile unsigned long long *>(c.q_stats);
        const unsigned b = (unsigned)min((unsigned long long)63, (tc_fin - t0) / 250000ull);
        atomicAdd(c.q_stats + 512 + b, tc_tot - tc0);      // totals
        atomicAdd(c.q_stats + 576 + b, tc_fin - tc_tot);   // stopping rule, solve, T
    }
#endif
    fence_release_side();  // the unit's new state is visible before its next round is published
    __syncwarp();
    if (!PERSIST) return;
#ifdef ORPCD_QUEUE_STATS
    if (!go_on && lane == 0) {   // finish stamp of the unit (ns since the first worker started), tagged with bit 63
        unsigned long long tf;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tf));
        reinterpret_cast<unsigned long long *>(c.worklist)[u] =
            (tf - *reinterpret_cast<volatile unsigned long long *>(c.q_stats)) | (1ull << 63);
    }
#endif
    if (!go_on) return;
    // publish the work items of the unit's next round: consecutive tickets of the ring.  The ticket atomic is the
    // release: the slot writes below depend on its result.
    const int n_next = (round + 1 >= c.half_from_round) ? c.n_patches2 : c.n_patches;
    unsigned base = 0;
    if (lane == 0) base = atom_add_release(c.q_ctrl + Q_TAIL, (unsigned)n_next);
    base = __shfl_sync(0xffffffffu, base, 0);
    // the ring holds at most one round per unit, so a slot still occupied belongs to a consumer that has taken its
    // ticket and is about to clear it.  Four compare-and-swaps per lane are issued together (their round trips
    // overlap); the rare occupied slot is retried afterwards.
    for (int k0 = lane; k0 < n_next; k0 += 128) {
        unsigned old[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = k0 + 32 * e;
            old[e] = k < n_next ? atomicCAS(c.q_slots + ((base + (unsigned)k) & c.q_mask), 0u,
                                            (unsigned)(u * c.n_patches2 + k) + 1u) : 0u;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = k0 + 32 * e;
            if (old[e] != 0u)
                while (atomicCAS(c.q_slots + ((base + (unsigned)k) & c.q_mask), 0u, (unsigned)(u * c.n_patches2 + k) + 1u) != 0u)
                    __nanosleep(32);
        }
    }
#ifdef ORPCD_QUEUE_STATS
    __syncwarp();
    if (lane == 0) {
        unsigned long long tc1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tc1));
        reinterpret_cast<unsigned long long *>(c.worklist)[u] = tc1;     // publication stamp of the next round
        const unsigned long long t0 = *reinterpret_cast<volatile unsigned long long *>(c.q_stats);
        const unsigned b = (unsigned)min((unsigned long long)63, (tc1 - t0) / 250000ull);
        atomicAdd(c.q_stats + 320 + b, tc1 - tc0);              // closing ns (totals, solve, publish)
        atomicAdd(c.q_stats + 384 + b, 1ull);                   // closings
    }
#endif
}

// 64 source points per warp (Q = 2): the smaller the patch, the tighter its sphere and the fewer chunks it
// visits, and the more warps a unit spreads over -- late rounds, when few units are still running, are bound by
// the latency of one warp's patch, not by throughput (measured: 128-point patches 16.3 ms, 64-point 13.8 ms
// of round kernels per 180-unit batch).
#ifndef ROUND_MIN_BLOCKS
#define ROUND_MIN_BLOCKS 6
#endif
#ifndef ROUND_MIN_BLOCKS_PLANE
// EST 1 (tensor-core Jacobian sums).  Measured on the 180-unit bench batch, persistent kernel, ms per batch:
// 8 blocks (64 registers, 436 B spills) 14.9, 7 (72) 13.2, 6 (80) 12.0, 5 (96 registers, 136 B spills) 11.7,
// 4 (128) 11.7 -- the kernel is bound by the shared-memory / L1 data pipe (58 % of peak over the whole launch, the
// target-chunk LDS.128 broadcasts of the sweep) and by its L1 hit rate, not by resident warps: spill traffic and
// a smaller L1 carve-out cost more than extra warps hide.
#define ROUND_MIN_BLOCKS_PLANE 5
#endif
#define ROUND_BLOCKS(EST) ((EST) == 1 ? ROUND_MIN_BLOCKS_PLANE : ROUND_MIN_BLOCKS)

#ifndef ROUND_W
#define ROUND_W 4            // warps per block (independent workers)
#endif
#ifndef ROUND_Q
#define ROUND_Q 2
#endif

template <int EST>
__global__ void __launch_bounds__(ROUND_W * 32, ROUND_BLOCKS(EST)) gicp_round_kernel(Ctx c, int round)
{
    __shared__ WarpSmem wsm[ROUND_W];     // 4 KB per warp
    const int u = blockIdx.z, warp = threadIdx.x >> 5;
    UnitState &st = c.state[u];
    if (st.done) return;
    round_body<ROUND_Q, EST, false>(c, st, u, blockIdx.x * ROUND_W + warp, round, wsm[warp]);
}

// The whole registration of every unit in ONE launch (the default path).  The grid is sized to what is resident at
// once; every warp is a worker that takes (unit, patch) items from a ring in global memory -- a ticket from the head
// counter, then the slot the ticket names -- and runs round_body on them.  Units advance independently: the warp that
// closes a unit's round (finish_round: stopping rule, 6x6 solve, T <- U T) publishes the patches of that unit's next
// round, so there is no lock step between units, no launch per round and no partly filled last wave: late in a batch
// the workers simply find the stragglers' patches in the ring.  The warp that finishes the last unit raises the done
// flag, which is what idle workers poll while they wait on an empty slot.
template <int EST>
__global__ void __launch_bounds__(ROUND_W * 32, ROUND_BLOCKS(EST)) gicp_persistent_kernel(Ctx c)
{
    __shared__ WarpSmem wsm[ROUND_W];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        unsigned item = 0u;
        if (lane == 0) {
#ifdef ORPCD_QUEUE_STATS
            unsigned long long tw0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tw0));
            if (c.q_stats[0] == 0ull) atomicCAS(c.q_stats, 0ull, tw0);
#endif
            const unsigned t = atomicAdd(c.q_ctrl + Q_HEAD, 1u);
            unsigned *slot = c.q_slots + (t & c.q_mask);
            for (;;) {
                if (*reinterpret_cast<volatile unsigned *>(slot) != 0u) {
                    item = atomicExch(slot, 0u);      // claim; 0 if the previous owner of this slot took it first
                    if (item != 0u) break;
                }
                if (*reinterpret_cast<volatile unsigned *>(c.q_ctrl + Q_DONE) != 0u) break;
                __nanosleep(ORPCD_SPIN_NS);      // an idle worker must not take issue slots from the busy warps of its SM
            }
#ifdef ORPCD_QUEUE_STATS
            unsigned long long tw1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tw1));
            const unsigned long long t0 = *reinterpret_cast<volatile unsigned long long *>(c.q_stats);
            const unsigned b = (unsigned)min((unsigned long long)63, (tw1 - t0) / 250000ull);
            atomicAdd(c.q_stats + 64 + b, tw1 - tw0);
            if (item) atomicAdd(c.q_stats + 128 + b, 1ull);
#endif
        }
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item == 0u) return;                       // every unit is done
        fence_acquire_side();                         // what the publisher wrote before the item is visible from here on
        const int u = (int)((item - 1u) / (unsigned)c.n_patches2), wi = (int)((item - 1u) % (unsigned)c.n_patches2);
        UnitState &st = c.state[u];
#ifdef ORPCD_QUEUE_STATS
        unsigned long long tp0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tp0));
#endif
        round_body<ROUND_Q, EST, true>(c, st, u, wi, __ldcg(&st.iters), wsm[warp]);
        __syncwarp();
#ifdef ORPCD_QUEUE_STATS
        if (lane == 0) {
            unsigned long long tp1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tp1));
            const unsigned long long t0 = *reinterpret_cast<volatile unsigned long long *>(c.q_stats);
            const unsigned b = (unsigned)min((unsigned long long)63, (tp1 - t0) / 250000ull);
            atomicAdd(c.q_stats + 192 + b, tp1 - tp0);          // busy ns (patch incl. a closing if this warp did one)
            atomicMax(c.q_stats + 256 + b, tp1 - tp0);          // longest item
        }
#endif
    }
}

// ring <- the patches of round 0 of every unit, unit-major (units of one target set are neighbours: the first
// workers share target chunks in L2 / L1)
__global__ void queue_init_kernel(Ctx c)
{
    const unsigned per_unit = (unsigned)(0 >= c.half_from_round ? c.n_patches2 : c.n_patches);   // items of round 0
    const unsigned cap = c.q_mask + 1u, n_items = (unsigned)c.n_units * per_unit;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x)
        c.q_slots[i] = i < n_items ? (i / per_unit) * (unsigned)c.n_patches2 + (i % per_unit) + 1u : 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c.q_ctrl[Q_HEAD] = 0u;
        c.q_ctrl[Q_TAIL] = n_items;
        c.q_ctrl[Q_DONE] = 0u;
    }
#ifdef ORPCD_QUEUE_STATS
    if (blockIdx.x == 0) for (int k = threadIdx.x; k < 640; k += blockDim.x) c.q_stats[k] = 0ull;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)c.n_units; i += gridDim.x * blockDim.x)
        reinterpret_cast<unsigned long long *>(c.worklist)[i] = 0ull;
#endif
}

// correspondences of unit u in the caller's source order (the unit arrays are spatially sorted)
__global__ void __launch_bounds__(256) unpermute_corr_kernel(Ctx c, int *__restrict__ out)
{
    const int u = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= c.n_src) return;
    const int i = c.sperm[(size_t)c.state[u].src_set * c.n_src + p];
    out[(size_t)u * c.n_src + i] = c.corr[(size_t)u * c.n_src + p];
}

// ----------------------------------------------------------------------------
// workspace layout
// ----------------------------------------------------------------------------
struct Layout {
    size_t off_state, off_pts, off_cov, off_qf, off_corr, off_d2, off_m1, off_c1, off_m2, off_tgt4, off_meta,
        off_partial, off_counter, off_worklist, off_wcount, off_ract, off_done, off_visit, off_sperm, off_tperm,
        off_sph, off_half, off_ssph, off_sort, off_bbox, off_qslots, off_qctrl, off_qstats, total;
    int nt_pad, n_chunks, splits, chunks_per_split, red_blocks, q_tiles, nn_q, nn_tj, cov_stride, smem_pad, nn_mode;
    int cull;
    int fused;       // one gicp_round_kernel launch per round instead of update / sweep / resolve / fix / reduce
    int persistent;  // ... and all rounds in one launch of gicp_persistent_kernel (the default)
    int n_patches;   // work items per unit and round of the fused kernels (64-point patches)
    int n_patches2;  // ... of their late rounds (32-point patches)
    int red_groups;  // groups of the two-level totals
    unsigned q_cap;  // work-queue ring capacity (power of two)
    int red_grid;    // blocks per unit of reduce_kernel (red_blocks is the partial-sum stride, >= patches per unit)
};

bool make_layout(const orpcd_gicp_params *p, Layout &L)
{
    if (!p || p->n_units <= 0 || p->n_src <= 0 || p->n_tgt <= 0 || p->n_src_sets <= 0 || p->n_tgt_sets <= 0 ||
        p->max_iterations < 0 || !(p->max_correspondence_distance > 0) || p->n_units > 65535 ||
        (p->cov_mode != 0 && p->cov_mode != 1))
        return false;
    if ((long long)p->n_units * p->n_src > 0x7fffffffLL) return false;
    if (p->cov_mode == 1 && !(p->gicp_epsilon > 0 && p->gicp_epsilon < 1)) return false;
    L.cov_stride = p->cov_mode == 0 ? 6 : 3;
    L.nt_pad = (p->n_tgt + NN_CH - 1) / NN_CH * NN_CH;
    L.n_chunks = L.nt_pad / NN_CH;
    L.nn_q = (p->sweep_q == 4 || p->sweep_q == 8) ? p->sweep_q : (p->n_src >= 2048 ? 8 : 4);
    L.smem_pad = p->sweep_smem_pad > 0 && p->sweep_smem_pad <= 190000 ? p->sweep_smem_pad : 0;
    L.nn_tj = (p->sweep_tj == 2 || p->sweep_tj == 4 || p->sweep_tj == 8) ? p->sweep_tj : 4;
    L.nn_mode = (p->sweep_mode >= 1 && p->sweep_mode <= 3) ? p->sweep_mode - 1 : 0;
    L.cull = !(p->sweep_mode >= 1 && p->sweep_mode <= 3);   // 0 (auto), 4, 5: culled sweep; 1..3: dense variants
    if (L.cull && !(p->sweep_q == 4 || p->sweep_q == 8)) L.nn_q = 4;   // smaller patches cull better
    L.fused = L.cull && p->sweep_mode != 4;  // 4: culled sweep with the five separate kernels
    L.persistent = L.fused && p->sweep_mode != 5;   // 5: fused rounds, one launch per round (A/B runs)
    L.q_tiles = (p->n_src + NN_THREADS * L.nn_q - 1) / (NN_THREADS * L.nn_q);
    int s = p->target_splits;
    if (s <= 0) {
        // dense sweep: enough blocks for ~12 waves of 148 SMs x 4 resident blocks, >= 8 chunks per split;
        // culled sweep: each warp already skips most chunks, 2 waves are enough
        const long long blocks = (long long)p->n_units * L.q_tiles;
        const long long want = L.cull ? 1480 : 7104;
        s = (int)((want + blocks - 1) / blocks);
        if (s > 8) s = 8;
    }
    if (L.cull) s = 1;   // the culled sweep walks whole groups of 32 chunks: one split
    const int s_max = L.n_chunks / 8 > 1 ? L.n_chunks / 8 : 1;
    if (s > s_max) s = s_max;
    if (s < 1) s = 1;
    L.chunks_per_split = (L.n_chunks + s - 1) / s;
    L.splits = (L.n_chunks + L.chunks_per_split - 1) / L.chunks_per_split;
    L.red_grid = (p->n_src + RED_THREADS * RED_PER_THREAD - 1) / (RED_THREADS * RED_PER_THREAD);
    const int patches = (p->n_src + 32 * ROUND_Q - 1) / (32 * ROUND_Q);   // fused rounds: one partial per patch
    const int patches2 = (p->n_src + 31) / 32;                            // ... or per 32-point patch (late rounds)
    {
        const int g1 = (patches + red_group_size(patches) - 1) / red_group_size(patches);
        const int g2 = (patches2 + red_group_size(patches2) - 1) / red_group_size(patches2);
        L.red_groups = g1 > g2 ? g1 : g2;
    }
    L.red_blocks = (L.red_grid > patches2 ? L.red_grid : patches2) + L.red_groups;   // item partials, then group partials
    L.n_patches = patches;
    L.n_patches2 = patches2;
    {
        const unsigned long long items = (unsigned long long)p->n_units * (unsigned long long)patches2;
        if (items >= 0x7fffffffull) return false;
        unsigned cap = 1024;
        while ((unsigned long long)cap < items) cap <<= 1;
        L.q_cap = cap;
    }
    if (L.fused) { L.chunks_per_split = L.n_chunks; L.splits = 1; }
    const size_t U = p->n_units, ns = p->n_src;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += orpcd_align_up(bytes, 256); return r; };
    L.off_state = take(U * sizeof(UnitState));
    L.off_pts = take(U * ns * 3 * sizeof(double));
    L.off_cov = take(U * ns * L.cov_stride * sizeof(double));
    L.off_qf = take(U * ns * sizeof(float4));
    L.off_corr = take(U * ns * sizeof(int));
    L.off_d2 = take(U * ns * sizeof(double));
    L.off_m1 = take(U * L.splits * ns * sizeof(float));
    L.off_c1 = take(U * L.splits * ns * sizeof(int));
    L.off_m2 = take(U * L.splits * ns * sizeof(float));
    L.off_tgt4 = take((size_t)p->n_tgt_sets * L.nt_pad * sizeof(float4));
    L.off_meta = take((size_t)p->n_tgt_sets * sizeof(TgtMeta));
    L.off_partial = take(U * L.red_blocks * RED_STRIDE_MMA * sizeof(double));
    L.off_counter = take(U * (size_t)(L.red_groups + 1) * sizeof(unsigned));
    L.off_worklist = take(U * ns * sizeof(int));
    L.off_wcount = take((size_t)(p->max_iterations + 2) * sizeof(int));
    L.off_ract = take((size_t)(p->max_iterations + 2) * sizeof(int));
    L.off_done = take(sizeof(int));
    L.off_visit = take((size_t)(p->max_iterations + 2) * sizeof(unsigned long long));
    L.off_sperm = take((size_t)p->n_src_sets * ns * sizeof(int));
    L.off_tperm = take((size_t)p->n_tgt_sets * L.nt_pad * sizeof(int));
    L.off_sph = take((size_t)p->n_tgt_sets * L.n_chunks * sizeof(float4));
    L.off_half = take((size_t)p->n_tgt_sets * L.n_chunks * sizeof(float4));
    L.off_ssph = take((size_t)p->n_tgt_sets * ((L.n_chunks + 31) / 32) * sizeof(float4));
    const size_t sort_s = (size_t)p->n_src_sets * sort_pad(p->n_src), sort_t = (size_t)p->n_tgt_sets * sort_pad(p->n_tgt);
    L.off_sort = take((sort_s > sort_t ? sort_s : sort_t) * sizeof(unsigned long long));
    L.off_bbox = take((size_t)(p->n_src_sets > p->n_tgt_sets ? p->n_src_sets : p->n_tgt_sets) * 6 * sizeof(double));
    L.off_qslots = take((size_t)L.q_cap * sizeof(unsigned));
    L.off_qctrl = take(128 * sizeof(unsigned));
    L.off_qstats = take(10 * 64 * sizeof(unsigned long long));
    L.total = o;
    return true;
}

void fill_ctx(const orpcd_gicp_params *p, const Layout &L, void *ws, Ctx &c)
{
    unsigned char *w = (unsigned char *)ws;
    c.n_units = p->n_units; c.n_src = p->n_src; c.n_tgt = p->n_tgt;
    c.nt_pad = L.nt_pad; c.n_chunks = L.n_chunks; c.splits = L.splits; c.chunks_per_split = L.chunks_per_split;
    c.max_iter = p->max_iterations; c.red_blocks = L.red_blocks;
    c.cov_mode = p->cov_mode; c.cov_stride = L.cov_stride; c.ignore_done = 0;
    c.r2 = p->max_correspondence_distance * p->max_correspondence_distance;
    c.rel_fit = p->relative_fitness; c.rel_rmse = p->relative_rmse;
    c.kappa = p->cov_mode == 1 ? 1.0 - p->gicp_epsilon : 0.0;
    c.state = (UnitState *)(w + L.off_state);
    c.cur_pts = (double *)(w + L.off_pts);
    c.cur_cov = (double *)(w + L.off_cov);
    c.cur_qf = (float4 *)(w + L.off_qf);
    c.corr = (int *)(w + L.off_corr);
    c.corr_d2 = (double *)(w + L.off_d2);
    c.part_m1 = (float *)(w + L.off_m1);
    c.part_c1 = (int *)(w + L.off_c1);
    c.part_m2 = (float *)(w + L.off_m2);
    c.tgt4 = (float4 *)(w + L.off_tgt4);
    c.tgt_meta = (TgtMeta *)(w + L.off_meta);
    c.red_partial = (double *)(w + L.off_partial);
    c.red_counter = (unsigned *)(w + L.off_counter);
    c.worklist = (int *)(w + L.off_worklist);
    c.worklist_count = (int *)(w + L.off_wcount);
    c.done_count = (int *)(w + L.off_done);
    c.round_active = (int *)(w + L.off_ract);
    c.visit_count = (unsigned long long *)(w + L.off_visit);
    c.sperm = (int *)(w + L.off_sperm);
    c.tperm = (int *)(w + L.off_tperm);
    c.chunk_sph = (float4 *)(w + L.off_sph);
    c.chunk_half = (float4 *)(w + L.off_half);
    c.super_sph = (float4 *)(w + L.off_ssph);
    c.n_super = (L.n_chunks + 31) / 32;
    c.sortbuf = (unsigned long long *)(w + L.off_sort);
    c.bbox = (double *)(w + L.off_bbox);
    c.q_slots = (unsigned *)(w + L.off_qslots);
    c.q_ctrl = (unsigned *)(w + L.off_qctrl);
    c.q_stats = (unsigned long long *)(w + L.off_qstats);
    c.q_mask = L.q_cap - 1u;
    c.n_patches = L.n_patches;
    c.n_patches2 = L.n_patches2;
    c.red_groups = L.red_groups;
    // default: 32-point patches from round 32 on for clouds of at most 16384 points (a unit-round of a larger cloud
    // has hundreds of patches and fills the GPU by itself: measured +8 % time at 200k points with the switch on)
    c.half_from_round = p->half_patch_from_round > 0 ? p->half_patch_from_round
                        : ((p->half_patch_from_round < 0 || p->n_src > 16384) ? 0x7fffffff : ORPCD_HALF_FROM_ROUND);
    c.dcap = (float)(p->max_correspondence_distance * (1.0 + 1e-5)) * (1.f + 1e-6f) + 1e-30f;
    c.tgt_pts = nullptr; c.tgt_cov = nullptr; c.out_T = nullptr; c.out_stats = nullptr; c.trace = nullptr;
    c.out_status = nullptr;
}

template <int Q, int TJ, int MODE>
void launch_sweep_m(const Ctx &c, const dim3 &g, int pad, int round, cudaStream_t st)
{
    // `pad` bytes of unused dynamic shared memory cap the resident blocks per SM (tuning)
    if (pad > 0)
        cudaFuncSetAttribute(nn_sweep_kernel<Q, TJ, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
    nn_sweep_kernel<Q, TJ, MODE><<<g, NN_THREADS, pad, st>>>(c, round);
}

template <int Q, int TJ>
void launch_sweep_t(const Ctx &c, const dim3 &g, int pad, int mode, int round, cudaStream_t st)
{
    if (mode == 0) launch_sweep_m<Q, TJ, 0>(c, g, pad, round, st);
    else if (mode == 2) launch_sweep_m<Q, TJ, 2>(c, g, pad, round, st);
    else launch_sweep_m<Q, TJ, 1>(c, g, pad, round, st);
}

void launch_sweep(const Ctx &c, const Layout &L, int n_units, int round, cudaStream_t st)
{
    const dim3 g(L.q_tiles, L.splits, n_units);
    const int pad = L.smem_pad;
    if (L.cull) {
        const int W = L.nn_tj == 2 ? 1 : (L.nn_tj == 8 ? 4 : 2);   // tuning: sweep_tj selects warps per block here
        const dim3 gc((c.n_src + W * 32 * L.nn_q - 1) / (W * 32 * L.nn_q), L.splits, n_units);
        if (L.nn_q == 8) {
            if (W == 1) nn_sweep_cull_kernel<8, 1><<<gc, 32, 0, st>>>(c, round);
            else if (W == 4) nn_sweep_cull_kernel<8, 4><<<gc, 128, 0, st>>>(c, round);
            else nn_sweep_cull_kernel<8, 2><<<gc, 64, 0, st>>>(c, round);
        } else {
            if (W == 1) nn_sweep_cull_kernel<4, 1><<<gc, 32, 0, st>>>(c, round);
            else if (W == 4) nn_sweep_cull_kernel<4, 4><<<gc, 128, 0, st>>>(c, round);
            else nn_sweep_cull_kernel<4, 2><<<gc, 64, 0, st>>>(c, round);
        }
        return;
    }
    if (L.nn_q == 8) {
        if (L.nn_tj == 2) launch_sweep_t<8, 2>(c, g, pad, L.nn_mode, round, st);
        else if (L.nn_tj == 8) launch_sweep_t<8, 8>(c, g, pad, L.nn_mode, round, st);
        else launch_sweep_t<8, 4>(c, g, pad, L.nn_mode, round, st);
    } else {
        if (L.nn_tj == 2) launch_sweep_t<4, 2>(c, g, pad, L.nn_mode, round, st);
        else if (L.nn_tj == 8) launch_sweep_t<4, 8>(c, g, pad, L.nn_mode, round, st);
        else launch_sweep_t<4, 4>(c, g, pad, L.nn_mode, round, st);
    }
}

template <int EST>
void launch_round_e(const Ctx &c, int n_units, int round, cudaStream_t st)
{
    const int items = round >= c.half_from_round ? c.n_patches2 : c.n_patches;
    const dim3 g((items + ROUND_W - 1) / ROUND_W, 1, n_units);
    gicp_round_kernel<EST><<<g, ROUND_W * 32, 0, st>>>(c, round);
}

// all rounds of all units in one launch: grid = what is resident at once, capped at the round-0 work
template <int EST>
int launch_persistent_e(const Ctx &c, cudaStream_t st)
{
    static int blocks_per_sm = 0, n_sm = 0;   // per template instance; one device per process (one process per GPU)
    if (blocks_per_sm == 0) {
        int dev = 0, b = 0, sm = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return -1;
        if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, gicp_persistent_kernel<EST>, ROUND_W * 32, 0) != cudaSuccess)
            return -1;
        if (b < 1 || sm < 1) return -1;
        n_sm = sm;
        blocks_per_sm = b;
    }
    const long long items = (long long)c.n_units * (0 >= c.half_from_round ? c.n_patches2 : c.n_patches);
    long long blocks = (long long)n_sm * blocks_per_sm;        // every block is resident: workers may spin on the ring
    const long long need = (items + ROUND_W - 1) / ROUND_W;
    if (blocks > need) blocks = need;
    queue_init_kernel<<<(unsigned)((c.q_mask + 1u + 255u) / 256u > 1184u ? 1184u : (c.q_mask + 1u + 255u) / 256u), 256, 0, st>>>(c);
    gicp_persistent_kernel<EST><<<(unsigned)blocks, ROUND_W * 32, 0, st>>>(c);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_persistent(const Ctx &c, int est, cudaStream_t st)
{
    if (est == 0) return launch_persistent_e<0>(c, st);
    if (est == 1) return launch_persistent_e<1>(c, st);
    if (est == 2) return launch_persistent_e<2>(c, st);
    if (est == 3) return launch_persistent_e<3>(c, st);
    return launch_persistent_e<4>(c, st);
}

// one fused round; est: 0 / 1 Generalized-ICP, 2 evaluation, 3 point-to-point, 4 point-to-plane
void launch_round(const Ctx &c, int est, int n_units, int round, cudaStream_t st)
{
    if (est == 0) launch_round_e<0>(c, n_units, round, st);
    else if (est == 1) launch_round_e<1>(c, n_units, round, st);
    else if (est == 2) launch_round_e<2>(c, n_units, round, st);
    else if (est == 3) launch_round_e<3>(c, n_units, round, st);
    else launch_round_e<4>(c, n_units, round, st);
}

// Morton permutations of all source and target sets, the sorted FP32 target tables and the unit state
int prepare_batch(const Ctx &c, const Layout &L, const orpcd_gicp_params *p, const double *d_src_points,
                  const double *d_src_cov, const int *d_unit_src, const int *d_unit_tgt, const double *d_init_T,
                  cudaStream_t st, int *launches)
{
    // rigid copies of one cloud keep its spatial coherence under the same point order: one sort serves them all
    const int sort_sets = p->src_rigid_copies ? 1 : p->n_src_sets;
    if (spatial_perm(d_src_points, p->n_src, sort_sets, p->n_src, c.bbox, c.sortbuf, c.sperm, st) != 0) return -1;
    if (sort_sets < p->n_src_sets)
        replicate_perm_kernel<<<dim3((p->n_src + 255) / 256, p->n_src_sets - 1), 256, 0, st>>>(c.sperm, p->n_src);
    if (spatial_perm(c.tgt_pts, p->n_tgt, p->n_tgt_sets, L.nt_pad, c.bbox, c.sortbuf, c.tperm, st) != 0) return -1;
    const int gx = (p->n_src + 255) / 256;
    init_units_kernel<<<dim3(gx > 64 ? 64 : gx, p->n_units), 256, 0, st>>>(c, d_unit_src, d_unit_tgt, d_src_points,
                                                                          d_src_cov, d_init_T);
    prep_targets_kernel<<<p->n_tgt_sets, 1024, 0, st>>>(c);
    if (launches) {
        int per_sort = 0;
        for (int n : {p->n_src, p->n_tgt}) {
            const int pad2 = sort_pad(n);
            per_sort += 4;   // bbox, morton, local sort, perm
            for (int k = 2 * SORT_TILE; k <= pad2; k <<= 1) {
                for (int j = k >> 1; j >= SORT_TILE; j >>= 1) ++per_sort;
                ++per_sort;
            }
        }
        *launches += per_sort + 2;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace

ORPCD_API size_t orpcd_gicp_batch_workspace_bytes(const orpcd_gicp_params *p)
{
    Layout L;
    if (!make_layout(p, L)) return 0;
    return L.total;
}

namespace {

// profiling-mode resources, released on every return path
struct ProfEvents {
    cudaEvent_t *ev = nullptr;
    int n = 0;
    bool init(int count)
    {
        ev = (cudaEvent_t *)malloc(sizeof(cudaEvent_t) * (size_t)count);
        if (!ev) return false;
        for (n = 0; n < count; ++n)
            if (cudaEventCreate(&ev[n]) != cudaSuccess) return false;
        return true;
    }
    ~ProfEvents()
    {
        for (int k = 0; k < n; ++k) cudaEventDestroy(ev[k]);
        free(ev);
    }
};
struct HostBuf {
    void *p = nullptr;
    bool alloc(size_t bytes) { p = malloc(bytes); return p != nullptr; }
    ~HostBuf() { free(p); }
};

// est_mode: 0 / 1 = Generalized-ICP (covariances / plane axes), 3 = point-to-point, 4 = point-to-plane
// (d_tgt_cov = target normals).  d_init_T: optional initial transforms [n_units][16].
int run_registration(const orpcd_gicp_params *p, int est_mode, const double *d_src_points,
                     const double *d_src_cov, const double *d_tgt_points, const double *d_tgt_cov,
                     const int *d_unit_src, const int *d_unit_tgt, const double *d_init_T, void *d_workspace,
                     size_t workspace_bytes, double *d_out_T, double *d_out_stats, int *d_out_status,
                     double *d_trace, double *h_profile, orpcd_stream_t stream)
{
    Layout L;
    if (!make_layout(p, L)) return ORPCD_ERR_INVALID;
    if (!d_src_points || !d_tgt_points || !d_unit_src || !d_unit_tgt || !d_out_T || !d_out_stats || !d_out_status)
        return ORPCD_ERR_INVALID;
    if ((est_mode <= 1 && (!d_src_cov || !d_tgt_cov)) || (est_mode == 4 && !d_tgt_cov)) return ORPCD_ERR_INVALID;
    if (!d_workspace || workspace_bytes < L.total || ((uintptr_t)d_workspace & 255)) return ORPCD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    Ctx c;
    fill_ctx(p, L, d_workspace, c);
    c.tgt_pts = d_tgt_points; c.tgt_cov = d_tgt_cov;
    c.out_T = d_out_T; c.out_stats = d_out_stats; c.out_status = d_out_status; c.trace = d_trace;
    c.cov_mode = est_mode;
    int launches = 0;

    if (prepare_batch(c, L, p, d_src_points, est_mode <= 1 ? d_src_cov : nullptr, d_unit_src, d_unit_tgt, d_init_T, st,
                      &launches) != 0)
        return ORPCD_ERR_CUDA;
    const dim3 g_upd((p->n_src + 255) / 256, p->n_units);
    const dim3 g_res((p->n_src + 127) / 128, p->n_units);
    const dim3 g_red(L.red_grid, p->n_units);
    const int fix_blocks = 148 * 4;
    // profiling mode (a diagnostic: it synchronises the stream and reads counters back before returning): CUDA
    // events around every launch of the dominant kernel, on the caller's stream.  Everything it owns is released
    // on every return path.
    const bool prof = p->profile != 0 && h_profile != nullptr;
    const int n_rounds_max = p->max_iterations + 1;
    ProfEvents pe;
    if (prof && !pe.init(L.persistent ? 2 : 2 * n_rounds_max)) return ORPCD_ERR_CUDA;
    int rounds = 0;
    if (L.persistent) {
        if (prof) cudaEventRecord(pe.ev[0], st);
        if (launch_persistent(c, est_mode, st) != 0) return ORPCD_ERR_CUDA;
        if (prof) cudaEventRecord(pe.ev[1], st);
        launches += 2;                       // queue_init_kernel + gicp_persistent_kernel
        rounds = n_rounds_max;               // upper bound; profiling mode replaces it by the rounds that ran
    } else {
        for (int round = 0; round <= p->max_iterations; ++round) {
            if (L.fused) {
                if (prof) cudaEventRecord(pe.ev[2 * round], st);
                launch_round(c, est_mode, p->n_units, round, st);
                if (prof) cudaEventRecord(pe.ev[2 * round + 1], st);
                launches += 1;
            } else {
                update_kernel<<<g_upd, 256, 0, st>>>(c, round);
                if (prof) cudaEventRecord(pe.ev[2 * round], st);
                launch_sweep(c, L, p->n_units, round, st);
                if (prof) cudaEventRecord(pe.ev[2 * round + 1], st);
                nn_resolve_kernel<<<g_res, 128, 0, st>>>(c, round);
                nn_fix_kernel<<<fix_blocks, 128, 0, st>>>(c, round);
                if (est_mode == 0) reduce_kernel<0><<<g_red, RED_THREADS, 0, st>>>(c, round);
                else if (est_mode == 1) reduce_kernel<1><<<g_red, RED_THREADS, 0, st>>>(c, round);
                else if (est_mode == 3) reduce_kernel<3><<<g_red, RED_THREADS, 0, st>>>(c, round);
                else reduce_kernel<4><<<g_red, RED_THREADS, 0, st>>>(c, round);
                launches += 5;
            }
            ORPCD_CHECK_LAUNCH();
            rounds = round + 1;
            if (p->poll_interval > 0 && round < p->max_iterations && (round + 1) % p->poll_interval == 0) {
                int done = 0;
                if (cudaMemcpyAsync(&done, c.done_count, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess)
                    return ORPCD_ERR_CUDA;
                if (cudaStreamSynchronize(st) != cudaSuccess) return ORPCD_ERR_CUDA;
                if (done >= p->n_units) break;
            }
        }
    }
    if (h_profile) {
        for (int k = 0; k < 8; ++k) h_profile[k] = 0.0;
        h_profile[0] = (double)launches;
        h_profile[3] = (double)rounds;
        h_profile[5] = (double)L.splits;
    }
    if (prof) {
        const int nr = p->max_iterations + 2;
        HostBuf hb;
        if (!hb.alloc((sizeof(int) * 2 + sizeof(unsigned long long)) * (size_t)nr)) return ORPCD_ERR_INVALID;
        unsigned long long *hv = (unsigned long long *)hb.p;
        int *h_active = (int *)(hv + nr), *h_fixed = h_active + nr;
        if (cudaMemcpyAsync(hv, c.visit_count, sizeof(unsigned long long) * nr, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(h_active, c.round_active, sizeof(int) * nr, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(h_fixed, c.worklist_count, sizeof(int) * nr, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            return ORPCD_ERR_CUDA;
        double kernel_ms = 0.0, unit_rounds = 0.0, fixed = 0.0, visited = 0.0;
        int launches_with_work = 0, rounds_run = 0;
        for (int r = 0; r < n_rounds_max; ++r) {
            if (h_active[r] <= 0) continue;
            rounds_run = r + 1;
            unit_rounds += h_active[r];
            fixed += h_fixed[r];
            visited += (double)hv[r];
            if (!L.persistent && r < rounds) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, pe.ev[2 * r], pe.ev[2 * r + 1]);
                kernel_ms += ms;
                ++launches_with_work;
            }
        }
        if (L.persistent) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, pe.ev[0], pe.ev[1]);   // includes the ring set-up kernel (microseconds)
            kernel_ms = ms;
            launches_with_work = 1;
            rounds = rounds_run;
            h_profile[3] = (double)rounds;
        }
#ifdef ORPCD_QUEUE_STATS
        {
            unsigned long long hq[640];
            cudaMemcpy(hq, c.q_stats, sizeof(hq), cudaMemcpyDeviceToHost);
            {
                std::vector<unsigned long long> fin(p->n_units);
                std::vector<int> it(p->n_units);
                cudaMemcpy(fin.data(), c.worklist, sizeof(unsigned long long) * p->n_units, cudaMemcpyDeviceToHost);
                std::vector<UnitState> hs(p->n_units);
                cudaMemcpy(hs.data(), c.state, sizeof(UnitState) * p->n_units, cudaMemcpyDeviceToHost);
                std::vector<int> ord(p->n_units);
                for (int i = 0; i < p->n_units; ++i) ord[i] = i;
                std::sort(ord.begin(), ord.end(), [&](int a, int b) { return (fin[a] & ~(1ull << 63)) > (fin[b] & ~(1ull << 63)); });
                fprintf(stderr, "last units to finish (unit: ms, iterations):");
                for (int i = 0; i < std::min(p->n_units, 24); ++i)
                    fprintf(stderr, " %d: %.2f, %d;", ord[i], (fin[ord[i]] & ~(1ull << 63)) * 1e-6, hs[ord[i]].iters);
                fprintf(stderr, "\n");
            }
            fprintf(stderr, "visit lists: %llu box survivors, %llu listed\n", hq[632], hq[633]);
            for (int b = 0; b < 64; ++b)
                if (hq[128 + b] || hq[64 + b])
                    fprintf(stderr, "t %5.2f ms  items %7llu  idle %5.1f %%  item mean %6.1f us max %6.1f us  closings %5llu mean %5.1f us\n",
                            b * 0.25, hq[128 + b], 100.0 * hq[64 + b] * 1e-6 / (0.25 * 148 * 20),
                            hq[128 + b] ? hq[192 + b] * 1e-3 / hq[128 + b] : 0.0, hq[256 + b] * 1e-3, hq[384 + b],
                            hq[384 + b] ? hq[320 + b] * 1e-3 / hq[384 + b] : 0.0),
                    fprintf(stderr, "            round span mean %6.1f us   closing: totals %5.1f us, solve %5.1f us, publish %5.1f us\n",
                            hq[384 + b] ? hq[448 + b] * 1e-3 / hq[384 + b] : 0.0, hq[384 + b] ? hq[512 + b] * 1e-3 / hq[384 + b] : 0.0,
                            hq[384 + b] ? hq[576 + b] * 1e-3 / hq[384 + b] : 0.0,
                            hq[384 + b] ? (hq[320 + b] - hq[512 + b] - hq[576 + b]) * 1e-3 / hq[384 + b] : 0.0);
        }
#endif
        if (getenv("ORPCD_DEBUG_ROUNDS")) {
            for (int r = 0; r < rounds; ++r)
                fprintf(stderr, "round %3d active %4d visited %.3f worklist %d\n", r, h_active[r],
                        h_active[r] > 0 ? (double)hv[r] / ((double)h_active[r] * p->n_src * p->n_tgt) : 0.0, h_fixed[r]);
        }
        h_profile[1] = kernel_ms;                       // total time inside the dominant kernel (round / sweep)
        h_profile[2] = visited;                         // (query, target) pairs the sweep evaluated
        h_profile[4] = fixed;                           // queries sent to the exact float64 search
        h_profile[6] = (double)launches_with_work;      // launches of the dominant kernel that had work
        h_profile[7] = unit_rounds;                     // brute force would be unit_rounds * n_src * n_tgt pairs
    }
    return ORPCD_OK;
}

}  // namespace

ORPCD_API int orpcd_gicp_batch(const orpcd_gicp_params *p, const double *d_src_points,
                               const double *d_src_cov, const double *d_tgt_points,
                               const double *d_tgt_cov, const int *d_unit_src, const int *d_unit_tgt,
                               void *d_workspace, size_t workspace_bytes, double *d_out_T,
                               double *d_out_stats, int *d_out_status, double *d_trace, double *h_profile,
                               orpcd_stream_t stream)
{
    if (!p || (p->cov_mode != 0 && p->cov_mode != 1)) return ORPCD_ERR_INVALID;
    return run_registration(p, p->cov_mode, d_src_points, d_src_cov, d_tgt_points, d_tgt_cov, d_unit_src, d_unit_tgt,
                            nullptr, d_workspace, workspace_bytes, d_out_T, d_out_stats, d_out_status, d_trace,
                            h_profile, stream);
}

ORPCD_API int orpcd_icp_batch(const orpcd_gicp_params *p_in, int point_to_plane, const double *d_src_points,
                              const double *d_tgt_points, const double *d_tgt_normals, const int *d_unit_src,
                              const int *d_unit_tgt, const double *d_init_T, void *d_workspace,
                              size_t workspace_bytes, double *d_out_T, double *d_out_stats, int *d_out_status,
                              orpcd_stream_t stream)
{
    if (!p_in || (point_to_plane && !d_tgt_normals)) return ORPCD_ERR_INVALID;
    orpcd_gicp_params q = *p_in;
    q.cov_mode = 1;                      // workspace layout of the plane-axis path (3 doubles per point)
    if (!(q.gicp_epsilon > 0 && q.gicp_epsilon < 1)) q.gicp_epsilon = 1e-3;
    q.profile = 0;
    return run_registration(&q, point_to_plane ? 4 : 3, d_src_points, nullptr, d_tgt_points, d_tgt_normals, d_unit_src,
                            d_unit_tgt, d_init_T, d_workspace, workspace_bytes, d_out_T, d_out_stats, d_out_status,
                            nullptr, nullptr, stream);
}

ORPCD_API int orpcd_evaluate_batch(const orpcd_gicp_params *p_in, const double *d_src_points,
                                   const double *d_tgt_points, const int *d_unit_src, const int *d_unit_tgt,
                                   const double *d_T, void *d_workspace, size_t workspace_bytes,
                                   double *d_out_stats, int *d_out_status, int *d_out_corr, orpcd_stream_t stream)
{
    if (!p_in) return ORPCD_ERR_INVALID;
    orpcd_gicp_params q = *p_in;
    q.cov_mode = 1;
    if (!(q.gicp_epsilon > 0 && q.gicp_epsilon < 1)) q.gicp_epsilon = 1e-3;
    Layout L;
    if (!make_layout(&q, L)) return ORPCD_ERR_INVALID;
    if (!d_src_points || !d_tgt_points || !d_unit_src || !d_unit_tgt || !d_out_stats || !d_out_status)
        return ORPCD_ERR_INVALID;
    if (!d_workspace || workspace_bytes < L.total || ((uintptr_t)d_workspace & 255)) return ORPCD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    Ctx c;
    fill_ctx(&q, L, d_workspace, c);
    c.cov_mode = 2;       // no covariances: Evaluate only
    c.max_iter = 0;       // round 0 is the evaluation; the unit is done right after it
    c.tgt_pts = d_tgt_points; c.tgt_cov = nullptr;
    // out_T is required by the reduction's epilogue: park it in the (unused) covariance slice
    c.out_T = c.cur_cov; c.out_stats = d_out_stats; c.out_status = d_out_status; c.trace = nullptr;
    if ((size_t)q.n_units * 16 > (size_t)q.n_units * q.n_src * 3) return ORPCD_ERR_INVALID;
    if (prepare_batch(c, L, &q, d_src_points, nullptr, d_unit_src, d_unit_tgt, d_T, st, nullptr) != 0)
        return ORPCD_ERR_CUDA;
    if (L.persistent) {
        if (launch_persistent(c, 2, st) != 0) return ORPCD_ERR_CUDA;
    } else if (L.fused) {
        launch_round(c, 2, q.n_units, 0, st);
    } else {
        update_kernel<<<dim3((q.n_src + 255) / 256, q.n_units), 256, 0, st>>>(c, 0);
        launch_sweep(c, L, q.n_units, 0, st);
        nn_resolve_kernel<<<dim3((q.n_src + 127) / 128, q.n_units), 128, 0, st>>>(c, 0);
        nn_fix_kernel<<<148 * 4, 128, 0, st>>>(c, 0);
        reduce_kernel<2><<<dim3(L.red_grid, q.n_units), RED_THREADS, 0, st>>>(c, 0);
    }
    if (d_out_corr)   // back to the caller's source order
        unpermute_corr_kernel<<<dim3((q.n_src + 255) / 256, q.n_units), 256, 0, st>>>(c, d_out_corr);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API int orpcd_gicp_sweep_only(const orpcd_gicp_params *p, void *d_workspace, size_t workspace_bytes,
                                    orpcd_stream_t stream)
{
    Layout L;
    if (!make_layout(p, L)) return ORPCD_ERR_INVALID;
    if (!d_workspace || workspace_bytes < L.total || ((uintptr_t)d_workspace & 255)) return ORPCD_ERR_WORKSPACE;
    Ctx c;
    fill_ctx(p, L, d_workspace, c);
    c.ignore_done = 1;
    launch_sweep(c, L, p->n_units, 0, (cudaStream_t)stream);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}
