// Fast Global Registration (Zhou, Park, Koltun) as Open3D 0.18.0 runs it, batched over
// the Aligner's (multi-start x scale-candidate) units, for sm_100a.
//
// Replaces open3d.pipelines.registration.registration_fgr_based_on_feature_matching
// as called from Optimizer/fastGlobalOptimizer.py:166-174.  Algorithm: SURVEY.md
// Appendix A.8 (FastGlobalRegistration.cpp: NormalizePointCloud, InitialMatching,
// AdvancedMatching, OptimizePairwiseRegistration, GetTransformationOriginalScale).
//
// Feature matching -- the one dense contraction on this path (N x M x 33) -- runs on the
// 5th-generation tensor cores:
//   feat_pack_kernel   float64 features -> two TF32 planes (hi + lo split, K padded 33->40)
//                      written in global memory as ready-made shared-memory images of
//                      128-row UMMA operand tiles (K-major, no swizzle) + |f|^2 per row.
//   feat_match_kernel  one CTA per 128 query rows: operand tiles arrive by TMA bulk copies
//                      (cp.async.bulk + mbarrier complete_tx, 3 stages), one elected thread
//                      issues tcgen05.mma kind::tf32 (3 products hi*hi + hi*lo + lo*hi per
//                      K step: FP32-class accuracy) into two 128x128 FP32 accumulators in
//                      TMEM, and the four warps read them back with tcgen05.ld to fold
//                      |d|^2 - 2 q.d into per-row (best 32-column chunk min, chunk id,
//                      second-best chunk min) while the next tile's MMAs run.
//   feat_resolve_kernel certifies with an error bound that the best chunk holds the exact
//                      nearest neighbour and re-scores its 32 candidates in float64 in
//                      nanoflann's summation order (ties -> lowest index); uncertified rows
//                      get an exact float64 scan.  Matched indices are therefore exact.
// The rest (mutual check, tuple test, graduated non-convexity solve) is small float64 work,
// one block per unit.
#include "common.cuh"

namespace {

// ----------------------------------------------------------------------------
// operand tile image
// ----------------------------------------------------------------------------
constexpr int FM_DIM = 33;
constexpr int FM_K = 40;                               // padded to a multiple of the tf32 MMA K (8)
constexpr int FM_ROWS = 128;                           // UMMA M = N = 128
constexpr int FM_LBO = FM_ROWS * 16;                   // 2048 B between 16-byte K chunks
constexpr int FM_SBO = 128;                            // 128 B between 8-row groups
constexpr int FM_PLANE_BYTES = FM_ROWS * FM_K * 4;     // 20480
constexpr int FM_NORM_OFF = 2 * FM_PLANE_BYTES;        // hi plane, lo plane, then 128 norms
constexpr int FM_TILE_BYTES = FM_NORM_OFF + FM_ROWS * 4;   // 41472
constexpr int FM_STAGES = 3;
// After the operand tiles of a set comes a feature-major float64 copy for the exact re-score, one block of
// [33][32] doubles per 32-row chunk, so that 32 lanes scoring the 32 rows of a chunk read 256 contiguous bytes
// per feature.
constexpr int FM_CHUNK_F64_BYTES = 33 * 32 * 8;         // 8448
constexpr int FM_SET_BYTES_PER_TILE = FM_TILE_BYTES + 4 * FM_CHUNK_F64_BYTES;   // 75264
constexpr int FM_ATILES = 2;                          // query tiles per CTA: every B tile brought in is used twice
constexpr int FM_EPI_WARPS = FM_ATILES * 4;           // warps 0-3 own the rows of A tile 0, warps 4-7 those of A tile 1
constexpr int FM_THREADS = (FM_EPI_WARPS + 1) * 32;   // + one warp that only drives TMA and tcgen05.mma
constexpr int FM_SMEM_BYTES = (FM_ATILES + FM_STAGES) * FM_TILE_BYTES + 128 /*barriers*/ + 128 /*alignment*/;
constexpr float FM_PAD_NORM = 3.0e38f;

__device__ __forceinline__ int fm_image_offset(int r, int c)
{
    return (c >> 2) * FM_LBO + (r >> 3) * FM_SBO + (r & 7) * 16 + (c & 3) * 4;
}

__device__ __forceinline__ float tf32_rna(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// grid (tiles, sets), 128 threads: thread r packs row r of the tile.
// chunks recorded per query row by the tensor-core filter (smallest 32-column chunk minima, ascending, with their ids;
// one more minimum is kept as the bound for everything not recorded).  With 3, 1.4 % of the FPFH rows exhausted their
// chunks without a certificate, and their exact scans of the whole set (1.3 MB each) were 60 % of
// feat_resolve_kernel's traffic.  Measured on the 180-unit FGR batch (ms per batch / rows needing the exact scan):
// 3: 19.4 / 25759, 4: 19.1 / 14524, 5: 18.7 / 9014, 6: 18.0 / 6008, 8: 18.0 / 3013, 10: 17.4 / 1703, 12: 17.9 / 1009.
#ifndef FM_TOP
#define FM_TOP 8
#endif

__global__ void __launch_bounds__(FM_ROWS)
feat_pack_kernel(const double *__restrict__ feat, int n, int tiles, const double *__restrict__ center,
                 unsigned char *__restrict__ packed, float *__restrict__ set_maxnorm)
{
    const int set = blockIdx.y, tile = blockIdx.x, r = threadIdx.x;
    const int row = tile * FM_ROWS + r;
    unsigned char *set_base = packed + (size_t)set * tiles * FM_SET_BYTES_PER_TILE;
    unsigned char *img = set_base + (size_t)tile * FM_TILE_BYTES;
    double *tr = reinterpret_cast<double *>(set_base + (size_t)tiles * FM_TILE_BYTES +
                                            (size_t)(tile * 4 + (r >> 5)) * FM_CHUNK_F64_BYTES) + (r & 31);
    float norm = FM_PAD_NORM;
    double nd = 0.0;
    const double *f = feat + ((size_t)set * n + (row < n ? row : 0)) * FM_DIM;
    for (int c = 0; c < FM_K; ++c) {
        float hi = 0.f, lo = 0.f;
        if (row < n && c < FM_DIM) {
            // distances are translation invariant: the filter works on features centred on `center`
            // (smaller norms -> a tighter error bound); the exact re-score uses the original values
            const double v = center ? f[c] - center[c] : f[c];
            nd += v * v;
            const float x = (float)v;
            hi = tf32_rna(x);
            lo = tf32_rna(x - hi);
        }
        if (c < FM_DIM) tr[c * 32] = row < n ? f[c] : 0.0;
        const int o = fm_image_offset(r, c);
        *reinterpret_cast<float *>(img + o) = hi;
        *reinterpret_cast<float *>(img + FM_PLANE_BYTES + o) = lo;
    }
    if (row < n) {
        norm = __double2float_rn(nd);
        atomicMax(reinterpret_cast<int *>(set_maxnorm + set), __float_as_int(norm));   // norms are >= 0
    }
    *reinterpret_cast<float *>(img + FM_NORM_OFF + r * 4) = norm;
}

// mean feature of one set (fixed summation order): 33 x 31 threads, thread (k, part) sums rows part, part+31, ...
__global__ void __launch_bounds__(1023) feat_center_kernel(const double *__restrict__ feat, int n, double *__restrict__ center)
{
    __shared__ double part[31][FM_DIM];
    const int k = threadIdx.x % FM_DIM, p = threadIdx.x / FM_DIM;
    double s = 0.0;
    for (int r = p; r < n; r += 31) s += feat[(size_t)r * FM_DIM + k];
    part[p][k] = s;
    __syncthreads();
    if (p == 0) {
        double t = 0.0;
        for (int q = 0; q < 31; ++q) t += part[q][k];
        center[k] = t / (double)n;
    }
}

// ----------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA bulk copy, tcgen05
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
// TMA engine, 1-D bulk copy global -> shared, completion counted on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, no swizzle, canonical ((8,n),2):((16 B, SBO), LBO)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(FM_LBO >> 4) << 16;
    d |= (uint64_t)(FM_SBO >> 4) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = 128
constexpr uint32_t FM_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(FM_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float min3f(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// issue only: tmem_wait_ld() before the registers are read
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------
// tensor-core filter: per query row (best chunk min, chunk id, second-best chunk min) of
//   v(row, col) = |d_col|^2 - 2 q_row . d_col          (= dist^2 - |q_row|^2)
// grid (pairs of query tiles, pairs of sets); 288 threads: warps 0-3 own the 128 rows of the CTA's first query tile,
// warps 4-7 those of the second, warp 8 only drives the TMA copies and the MMAs (warp-specialised: the epilogue of
// tile e overlaps the issue and execution of tile e+1); every database tile staged in shared memory feeds the MMAs
// of both query tiles.
// ----------------------------------------------------------------------------
template <bool DEBUG>
__global__ void __launch_bounds__(FM_THREADS, 1)
feat_match_kernel(const unsigned char *__restrict__ packed_q, int q_tiles, const unsigned char *__restrict__ packed_db,
                  int db_tiles, const int *__restrict__ pair_q, const int *__restrict__ pair_db, int nq,
                  float *__restrict__ out_m, int *__restrict__ out_c, size_t out_stride, float *__restrict__ dbg_v)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    unsigned char *sA = smem;                                  // FM_ATILES query tiles
    unsigned char *sB = smem + FM_ATILES * FM_TILE_BYTES;      // FM_STAGES database tiles
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (FM_ATILES + FM_STAGES) * FM_TILE_BYTES);
    // bars[0]: A landed; bars[1..3]: B stage landed; bars[4..5]: accumulators complete; bars[6] low word: TMEM base;
    // bars[7..8]: accumulator buffer released by the epilogue warps
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int atile = warp >> 2, quarter = warp & 3;           // TMEM lanes 32*quarter.. belong to warps quarter, quarter+4
    const int pair = blockIdx.y, qt0 = blockIdx.x * FM_ATILES;
    const int n_at = min(FM_ATILES, q_tiles - qt0);            // query tiles this CTA really has
    const unsigned char *gA = packed_q + (size_t)pair_q[pair] * q_tiles * FM_SET_BYTES_PER_TILE + (size_t)qt0 * FM_TILE_BYTES;
    const unsigned char *gB = packed_db + (size_t)pair_db[pair] * db_tiles * FM_SET_BYTES_PER_TILE;
    const uint32_t bar_a = smem_u32(bars), bar_full0 = smem_u32(bars + 1), bar_mma0 = smem_u32(bars + 4);
    const uint32_t bar_free0 = smem_u32(bars + 7);   // bars[7..8]: accumulator buffer (and the stage of its tile) free again

    if (tid == 0) {
        mbar_init(bar_a, 1);
        for (int s = 0; s < FM_STAGES; ++s) mbar_init(bar_full0 + 8 * s, 1);
        mbar_init(bar_mma0, 1);
        mbar_init(bar_mma0 + 8, 1);
        mbar_init(bar_free0, FM_EPI_WARPS);
        mbar_init(bar_free0 + 8, FM_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // all 512 TMEM columns: (2 buffers) x (2 query tiles) 128 x 128 FP32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // the FM_TOP smallest chunk minima (ascending, with chunk ids) and the next one as the bound for the rest
    float mm[FM_TOP + 1];
    int cc[FM_TOP];
#pragma unroll
    for (int k = 0; k <= FM_TOP; ++k) mm[k] = CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < FM_TOP; ++k) cc[k] = 0;
    const int row = (qt0 + atile) * FM_ROWS + quarter * 32 + (tid & 31);
    const bool have_tile = warp < FM_EPI_WARPS && atile < n_at;

    if (warp == FM_EPI_WARPS) {
        // ---- producer warp: one thread brings the tiles in with TMA bulk copies and issues the MMAs
        if ((tid & 31) == 0) {
            mbar_expect_tx(bar_a, (uint32_t)n_at * FM_TILE_BYTES);
            tma_bulk_g2s(smem_u32(sA), gA, (uint32_t)n_at * FM_TILE_BYTES, bar_a);   // the tiles of a set are contiguous
            for (int s = 0; s < FM_STAGES && s < db_tiles; ++s) {
                mbar_expect_tx(bar_full0 + 8 * s, FM_TILE_BYTES);
                tma_bulk_g2s(smem_u32(sB + (size_t)s * FM_TILE_BYTES), gB + (size_t)s * FM_TILE_BYTES, FM_TILE_BYTES,
                             bar_full0 + 8 * s);
            }
            mbar_wait(bar_a, 0);
            for (int t = 0; t < db_tiles; ++t) {
                if (t >= 2) {
                    // the epilogue of tile t-2 is done: its accumulators (buffer t&1) and its stage are free
                    mbar_wait(bar_free0 + 8 * (t & 1), ((t - 2) >> 1) & 1);
                    const int nt = t + 1;   // stage (t - 2) % 3 == (t + 1) % 3
                    if (nt < db_tiles) {
                        const int ns = nt % FM_STAGES;
                        mbar_expect_tx(bar_full0 + 8 * ns, FM_TILE_BYTES);
                        tma_bulk_g2s(smem_u32(sB + (size_t)ns * FM_TILE_BYTES), gB + (size_t)nt * FM_TILE_BYTES,
                                     FM_TILE_BYTES, bar_full0 + 8 * ns);
                    }
                }
                const int s = t % FM_STAGES;
                mbar_wait(bar_full0 + 8 * s, (t / FM_STAGES) & 1);
                tc_fence_after();
                const uint32_t b_hi = smem_u32(sB + (size_t)s * FM_TILE_BYTES), b_lo = b_hi + FM_PLANE_BYTES;
                for (int at = 0; at < n_at; ++at) {
                    const uint32_t a_hi = smem_u32(sA + (size_t)at * FM_TILE_BYTES), a_lo = a_hi + FM_PLANE_BYTES;
                    const uint32_t acc = tmem_base + (uint32_t)((t & 1) * FM_ATILES + at) * FM_ROWS;
#pragma unroll
                    for (int k = 0; k < FM_K / 8; ++k) {
                        const uint32_t ko = (uint32_t)k * 2 * FM_LBO;   // two 16-byte K chunks per MMA (K = 8 tf32)
                        umma_tf32(acc, umma_desc(a_hi + ko), umma_desc(b_hi + ko), k > 0);
                        umma_tf32(acc, umma_desc(a_hi + ko), umma_desc(b_lo + ko), 1);
                        umma_tf32(acc, umma_desc(a_lo + ko), umma_desc(b_hi + ko), 1);
                    }
                }
                umma_commit(bar_mma0 + 8 * (t & 1));   // implies tcgen05.fence::before_thread_sync
            }
        }
    } else {
        // ---- epilogue warps: tile e while the tensor core works on tile e+1
        for (int e = 0; e < db_tiles; ++e) {
            const int s = e % FM_STAGES;
            mbar_wait(bar_full0 + 8 * s, (e / FM_STAGES) & 1);   // tile e's norms (TMA-written) visible to this thread
            mbar_wait(bar_mma0 + 8 * (e & 1), (e >> 1) & 1);
            tc_fence_after();
            if (have_tile) {
                const uint32_t dn_addr = smem_u32(sB + (size_t)s * FM_TILE_BYTES + FM_NORM_OFF);
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                       (uint32_t)((e & 1) * FM_ATILES + atile) * FM_ROWS;
                uint32_t v[4][32];   // this row's 128 columns: four loads in flight, one wait
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) tmem_ld32(taddr + ch * 32, v[ch]);
                tmem_wait_ld();
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    float ma = CUDART_INF_F, mb = CUDART_INF_F;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float d0, d1, d2, d3;   // |d|^2 of four columns: one broadcast 16-byte shared load
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(d0), "=f"(d1), "=f"(d2), "=f"(d3) : "r"(dn_addr + (uint32_t)(ch * 32 + j) * 4));
                        const float v0 = __fmaf_rn(-2.f, __uint_as_float(v[ch][j]), d0);
                        const float v1 = __fmaf_rn(-2.f, __uint_as_float(v[ch][j + 1]), d1);
                        const float v2 = __fmaf_rn(-2.f, __uint_as_float(v[ch][j + 2]), d2);
                        const float v3 = __fmaf_rn(-2.f, __uint_as_float(v[ch][j + 3]), d3);
                        ma = min3f(ma, v0, v1);
                        mb = min3f(mb, v2, v3);
                        if (DEBUG && row < nq) {
                            float *o = dbg_v + ((size_t)pair * nq + row) * ((size_t)db_tiles * FM_ROWS) + (size_t)e * FM_ROWS +
                                       ch * 32 + j;
                            o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3;
                        }
                    }
                    float x = fminf(ma, mb);
                    int cx = e * 4 + ch;
                    if (x < mm[FM_TOP - 1]) {   // sorted insertion; strict compares keep the lowest chunk id first among equals
#pragma unroll
                        for (int k = 0; k < FM_TOP; ++k)
                            if (x < mm[k]) { const float tt = mm[k]; mm[k] = x; x = tt; const int tc = cc[k]; cc[k] = cx; cx = tc; }
                    }
                    mm[FM_TOP] = fminf(mm[FM_TOP], x);
                }
            }
            // this warp is done with accumulator buffer e&1 and with stage s (the norms)
            tc_fence_before();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(bar_free0 + 8 * (e & 1));
        }
    }
    if (have_tile && row < nq) {
        const size_t o = (size_t)pair * nq + row;
#pragma unroll
        for (int k = 0; k <= FM_TOP; ++k) out_m[o + (size_t)k * out_stride] = mm[k];
#pragma unroll
        for (int k = 0; k < FM_TOP; ++k) out_c[o + (size_t)k * out_stride] = cc[k];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// nanoflann L2_Adaptor::evalMetric for dim 33: groups of four summed left to right, then the tail;
// b is row (lane) of a feature-major chunk block: feature k at b[k * 32].
__device__ __forceinline__ double sqdist33_fm(const double *__restrict__ a, const double *__restrict__ b)
{
    double result = 0.0;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const double d0 = a[i] - b[i * 32], d1 = a[i + 1] - b[(i + 1) * 32], d2 = a[i + 2] - b[(i + 2) * 32],
                     d3 = a[i + 3] - b[(i + 3) * 32];
        result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    const double d0 = a[32] - b[32 * 32];
    result += d0 * d0;
    return result;
}

// One warp per query row.  The chunks with the smallest filter minima are re-scored exactly (float64,
// nanoflann's summation order) in ascending order of their minima until the next minimum provably lies
// above the best exact distance found so far; a row that exhausts the FM_TOP recorded chunks without
// reaching that certificate gets the exact scan of the whole set.  Ties -> lowest index.
// stats[0] += rows that needed the exact scan, stats[1] += chunks re-scored.

#ifndef FR_MIN_BLOCKS
#define FR_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(128, FR_MIN_BLOCKS)
feat_resolve_kernel(const double *__restrict__ feat_q, int nq, const unsigned char *__restrict__ packed_db,
                    int db_tiles, int ndb, const double *__restrict__ center, const int *__restrict__ pair_q, const int *__restrict__ pair_db,
                    const float *__restrict__ db_maxnorm, const float *__restrict__ in_m,
                    const int *__restrict__ in_c, size_t in_stride, int force_exact,
                    const unsigned char *__restrict__ row_mask,
                    int *__restrict__ out_nn, double *__restrict__ out_d2, int *__restrict__ stats)
{
    const int pair = blockIdx.y, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= nq) return;
    if (row_mask && !row_mask[(size_t)pair * nq + row]) {     // the caller does not need this row (see the header)
        if (lane == 0) {
            out_nn[(size_t)pair * nq + row] = -1;
            if (out_d2) out_d2[(size_t)pair * nq + row] = CUDART_INF;
        }
        return;
    }
    const double *q = feat_q + ((size_t)pair_q[pair] * nq + row) * FM_DIM;
    // feature-major float64 chunk blocks of the database set; lane l scores row (chunk * 32 + l)
    const double *db = reinterpret_cast<const double *>(packed_db + (size_t)pair_db[pair] * db_tiles * FM_SET_BYTES_PER_TILE +
                                                        (size_t)db_tiles * FM_TILE_BYTES) + lane;
    constexpr int CH = FM_CHUNK_F64_BYTES / 8;
    // the query row lives in shared memory (broadcast reads): 66 registers less per thread, 2.4x the resident warps
    __shared__ double qsh[4][FM_DIM + 3];
    double *qv = qsh[threadIdx.x >> 5];
    qv[lane] = q[lane];
    if (lane == 0) qv[32] = q[32];
    __syncwarp();
    // |q - center|^2: what the filter value is offset by.  One term per lane and a butterfly (it only enters bounds
    // that carry slack for its rounding; every lane summing all 33 terms was a third of the kernel's instructions).
    double qn;
    {
        const double c0 = center ? qv[lane] - center[lane] : qv[lane];
        qn = c0 * c0;
        if (lane == 0) { const double c1 = center ? qv[32] - center[32] : qv[32]; qn += c1 * c1; }
        for (int s = 16; s > 0; s >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, s);
    }
    const size_t o = (size_t)pair * nq + row;
    // filter error bound (DESIGN.md "feature filter bound"): |v - (d^2 - |q|^2)| <= E,
    //   E = 2^-15 |q| Dmax + 2^-23 Dmax^2 on the centred features; float64 rounding of the centring and of
    //   |q|^2 (2^-52 relative) is folded into the last term
    const double dmax = sqrt((double)db_maxnorm[pair_db[pair]]);
    const double E = (ldexp(sqrt(qn) * dmax, -15) + ldexp(dmax * dmax, -23)) * 1.0001 + ldexp(qn + dmax * dmax, -40);
    // (best, bj): the best exact (distance, index) so far, the same on every lane -- one arg-min butterfly per scored
    // chunk, which also serves the next certificate test and the result
    double best = CUDART_INF;
    int bj = 0x7fffffff;
    bool certified = false;
    int scored = 0;
    if (!force_exact) {
        // the recorded minima and chunk ids of this row: lane k holds entry k (one load each instead of one per step)
        const float my_m = lane <= FM_TOP ? in_m[o + (size_t)lane * in_stride] : 0.f;
        const int my_c = lane < FM_TOP ? in_c[o + (size_t)lane * in_stride] : 0;
#pragma unroll 1
        for (int k = 0; k < FM_TOP; ++k) {
            // chunk k holds nothing better than in_m[k] - E; everything not recorded nothing better than in_m[k..]
            const double lb = (double)__shfl_sync(0xffffffffu, my_m, k) - E + qn;
            if (lb > best) { certified = true; break; }
            const int ck = __shfl_sync(0xffffffffu, my_c, k);
            const int j = ck * 32 + lane;
            double d = CUDART_INF;
            int dj = 0x7fffffff;
            if (j < ndb) { d = sqdist33_fm(qv, db + (size_t)ck * CH); dj = j; }
            if (best < d || (best == d && bj < dj)) { d = best; dj = bj; }     // carry the running best along
            for (int s = 16; s > 0; s >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, d, s);
                const int oj = __shfl_xor_sync(0xffffffffu, dj, s);
                if (od < d || (od == d && oj < dj)) { d = od; dj = oj; }
            }
            best = d; bj = dj;
            ++scored;
        }
        if (!certified) certified = (double)__shfl_sync(0xffffffffu, my_m, FM_TOP) - E + qn > best;
    }
    if (!certified) {
        best = CUDART_INF;
        bj = 0x7fffffff;
        for (int j = lane; j < ndb; j += 32) {
            const double d = sqdist33_fm(qv, db + (size_t)(j >> 5) * CH);
            if (d < best) { best = d; bj = j; }
        }
        if (lane == 0 && stats) atomicAdd(stats, 1);
        for (int s = 16; s > 0; s >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, best, s);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, s);
            if (od < best || (od == best && oj < bj)) { best = od; bj = oj; }
        }
    }
    if (lane == 0) {
        out_nn[o] = bj == 0x7fffffff ? -1 : bj;
        if (out_d2) out_d2[o] = best;
        if (stats && scored) atomicAdd(stats + 1, scored);
    }
}

// ----------------------------------------------------------------------------
// per-unit float64 stages (one block per unit)
// ----------------------------------------------------------------------------
constexpr int FG_THREADS = 256;

struct FgrUnit {
    double mean_s[3], mean_t[3];
    double scale;        // scale_global
    int n_init, n_pairs, trials, swapped;
    int tt_done, tt_cnt;  // tuple test: quota reached or trials exhausted; accepted tuples so far
};

struct FgrCtx {
    int n_units, n_src, n_tgt, n_feat_src, n_feat_tgt, cap;
    int iteration_number, decrease_mu, max_tuple, tuple_test;
    double division_factor, tuple_scale, max_corr_dist;
    const double *src_pts, *tgt_pts;
    const int *unit_src, *unit_tgt;
    const int *ij, *ji;                 // [U][n_feat_src], [U][n_feat_tgt]
    const unsigned long long *keys;     // [U]
    FgrUnit *unit;
    int2 *corres_init, *corres;         // [U][cap]
    double *qcur, *pcur;                // [U][cap][3] each
    unsigned *tt_flags;                 // [U][tt_epoch / 32] accept bits of the tuple-test epoch in hand
    double *tt_pairs;                   // [U][cap][8] (source xyz, target xyz, 0, 0) normalised, of the cross-checked pairs
    int tt_epoch;                       // trials of a unit evaluated per epoch (multiple of 1024)
    double *out_T;
    int *out_info;
};

__device__ __forceinline__ double block_sum(double v, double *sh)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < FG_THREADS / 32; ++w) s += sh[w];
    return s;
}
__device__ __forceinline__ double block_max(double v, double *sh)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < FG_THREADS / 32; ++w) s = fmax(s, sh[w]);
    return s;
}
// exclusive block scan of 0/1 flags; returns this thread's offset, total in `total`
__device__ __forceinline__ int block_scan_flag(int flag, int *sh, int &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int within = __popc(bal & ((1u << lane) - 1u));
    __syncthreads();
    if (lane == 0) sh[wid] = __popc(bal);
    __syncthreads();
    int base = 0, tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        if (w < wid) base += sh[w];
        tot += sh[w];
    }
    total = tot;
    return base + within;
}

// exclusive block scan of small per-thread counts; returns this thread's offset, total in `total`
__device__ __forceinline__ int block_scan_count(int cnt, int *sh, int &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    __syncthreads();
    if (lane == 31) sh[wid] = incl;
    __syncthreads();
    int base = 0, tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        if (w < wid) base += sh[w];
        tot += sh[w];
    }
    total = tot;
    return base + incl - cnt;
}

__device__ __forceinline__ void norm_point(const double *p, int i, const double *mean, double scale, double o[3])
{
    o[0] = (p[(size_t)i * 3] - mean[0]) / scale;
    o[1] = (p[(size_t)i * 3 + 1] - mean[1]) / scale;
    o[2] = (p[(size_t)i * 3 + 2] - mean[2]) / scale;
}

// NormalizePointCloud (means, global scale) + InitialMatching's cross check, per unit.
__global__ void __launch_bounds__(FG_THREADS) fgr_prepare_kernel(FgrCtx c)
{
    __shared__ double shd[FG_THREADS / 32];
    __shared__ int shi[FG_THREADS / 32];
    __shared__ int s_cnt;
    const int u = blockIdx.x, tid = threadIdx.x;
    const double *sp = c.src_pts + (size_t)c.unit_src[u] * c.n_src * 3;
    const double *tp = c.tgt_pts + (size_t)c.unit_tgt[u] * c.n_tgt * 3;
    double mean[2][3];
    for (int w = 0; w < 2; ++w) {
        const double *p = w == 0 ? sp : tp;
        const int n = w == 0 ? c.n_src : c.n_tgt;
        for (int k = 0; k < 3; ++k) {
            double a = 0.0;
            for (int i = tid; i < n; i += FG_THREADS) a += p[(size_t)i * 3 + k];
            mean[w][k] = block_sum(a, shd) / (double)n;
        }
    }
    double mx = 0.0;
    for (int w = 0; w < 2; ++w) {
        const double *p = w == 0 ? sp : tp;
        const int n = w == 0 ? c.n_src : c.n_tgt;
        for (int i = tid; i < n; i += FG_THREADS) {
            const double x = p[(size_t)i * 3] - mean[w][0], y = p[(size_t)i * 3 + 1] - mean[w][1],
                         z = p[(size_t)i * 3 + 2] - mean[w][2];
            mx = fmax(mx, sqrt(x * x + y * y + z * z));
        }
    }
    const double scale = block_max(mx, shd);
    // cross check in ascending source-feature index (std::map order)
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    const int *ij = c.ij + (size_t)u * c.n_feat_src;
    const int *ji = c.ji + (size_t)u * c.n_feat_tgt;
    int2 *out = c.corres_init + (size_t)u * c.cap;
    for (int base = 0; base < c.n_feat_src; base += FG_THREADS) {
        const int i = base + tid;
        int flag = 0, j = -1;
        if (i < c.n_feat_src) {
            j = ij[i];
            flag = j >= 0 && j < c.n_feat_tgt && ji[j] == i && i < c.n_src && j < c.n_tgt;
        }
        int total;
        const int off = block_scan_flag(flag, shi, total);
        const int cnt0 = s_cnt;
        if (flag) out[cnt0 + off] = make_int2(i, j);
        __syncthreads();
        if (tid == 0) s_cnt = cnt0 + total;
        __syncthreads();
    }
    // normalised end points of the cross-checked pairs, once per pair: the tuple test draws ~100 x n_init triples of
    // them (norm_point in every trial was 18 float64 divisions per trial)
    {
        const int n_init = s_cnt;
        // one 64-byte record per pair (two sectors, three 16-byte loads): the trials gather them at random and the
        // tuple test is bound by the L1's request rate
        double2 *rec = reinterpret_cast<double2 *>(c.tt_pairs + (size_t)u * c.cap * 8);
        for (int i = tid; i < n_init; i += FG_THREADS) {
            const int2 pr = out[i];
            double a[3], b[3];
            norm_point(sp, pr.x, mean[0], scale, a);
            norm_point(tp, pr.y, mean[1], scale, b);
            rec[(size_t)i * 4] = make_double2(a[0], a[1]);
            rec[(size_t)i * 4 + 1] = make_double2(a[2], b[0]);
            rec[(size_t)i * 4 + 2] = make_double2(b[1], b[2]);
            rec[(size_t)i * 4 + 3] = make_double2(0.0, 0.0);
        }
    }
    if (tid == 0) {
        FgrUnit &fu = c.unit[u];
        for (int k = 0; k < 3; ++k) { fu.mean_s[k] = mean[0][k]; fu.mean_t[k] = mean[1][k]; }
        fu.scale = scale;
        fu.n_init = s_cnt;
        fu.n_pairs = 0; fu.trials = 0;
        fu.swapped = c.n_tgt > c.n_src ? 1 : 0;
        fu.tt_done = 0; fu.tt_cnt = 0;
    }
}

// Philox4x32-10, counter (t, 0, 0, 0): the project's documented tuple-test stream
__device__ __forceinline__ void philox4x32_10(unsigned long long counter, unsigned long long key, uint32_t out[4])
{
    uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32), c2 = 0u, c3 = 0u;
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r > 0) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}


// AdvancedMatching's tuple test.  Open3D runs up to 100 x n_init trials one after the other and stops at
// maximum_tuple_count accepted triples.  Trials are independent given the counter-based stream (trial t draws
// philox4x32-10(t, key)), so the test is split in two:
//   fgr_tuple_eval_kernel  evaluates an EPOCH of trials of every unit across the whole GPU (grid: 1024-trial blocks x
//                          units; one accept bit per trial, a ballot word per warp);
//   fgr_tuple_pick_kernel  (one block per unit) walks the epoch's accept bits in trial order, re-derives the
//                          accepted triples and appends them until the quota is reached -- the sequential
//                          semantics, including the count of trials the sequential loop would have run.
// Units whose quota is full (or whose trials are exhausted) skip the remaining epochs.  (Round 1 ran one 1024-thread
// block per unit over all its trials: 2.06 ms per 180-unit batch, two waves of 148 + 32 blocks, 18 float64 divisions
// and 6 square roots per trial.)
// The predicate `li * ts < lj && lj < li / ts` on lengths is decided on the SQUARED lengths whenever they are not
// within 1e-9 (relative) of the boundary -- the roundings of the square roots, the product and the quotient move the
// exact predicate by a few 1e-16 -- and evaluated literally otherwise: identical decisions, no sqrt or division on
// the common path.
constexpr int TT_BLOCK_TRIALS = 1024;

__device__ __forceinline__ double sqlen3(const double *a, const double *b)
{
    const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return x * x + y * y + z * z;       // dist3's sum before the square root
}
// one side of the tuple test: li * ts < lj && lj < li / ts, with Li = li^2, Lj = lj^2 (as dist3 would form them)
__device__ __forceinline__ bool tuple_edge_ok(double Li, double Lj, double ts, double ts2_hi, double ts2_lo)
{
    // li ts < lj  <=>  Li ts^2 < Lj ;  lj < li / ts  <=>  Lj ts^2 < Li   (real arithmetic, all lengths >= 0)
    const double a_hi = Li * ts2_hi, a_lo = Li * ts2_lo, b_hi = Lj * ts2_hi, b_lo = Lj * ts2_lo;
    const bool regular = Li > 1e-200 && Lj > 1e-200;   // (zero lengths, denormals and NaNs take the literal test)
    if (regular && a_hi < Lj && b_hi < Li) return true;        // clear of both boundaries
    if (regular && (a_lo > Lj || b_lo > Li)) return false;     // clearly violates one
    const double li = sqrt(Li), lj = sqrt(Lj);      // near a boundary (or not finite): the literal test
    return (li * ts < lj) && (lj < li / ts);
}

__global__ void __launch_bounds__(256) fgr_tuple_eval_kernel(FgrCtx c, long long epoch_base)
{
    const int u = blockIdx.y, tid = threadIdx.x;
    const FgrUnit &fu = c.unit[u];
    if (fu.tt_done) return;
    const int ncorr = fu.n_init;
    const long long n_trials = (long long)ncorr * 100;
    const long long block_base = epoch_base + (long long)blockIdx.x * TT_BLOCK_TRIALS;
    if (block_base >= n_trials) return;
    const double2 *rec = reinterpret_cast<const double2 *>(c.tt_pairs + (size_t)u * c.cap * 8);   // normalised end points of init pair i
    const double ts = c.tuple_scale;
    const double ts2_hi = ts * ts * (1.0 + 1e-9), ts2_lo = ts * ts * (1.0 - 1e-9);
    const bool swapped = fu.swapped != 0;
    const unsigned long long key = c.keys[u];
    unsigned *words = c.tt_flags + (size_t)u * (c.tt_epoch / 32);
#pragma unroll 2
    for (int e = 0; e < TT_BLOCK_TRIALS / 256; ++e) {
        const long long t = block_base + e * 256 + tid;
        bool flag = false;
        if (t < n_trials) {
            uint32_t r[4];
            philox4x32_10((unsigned long long)t, key, r);
            const int i0 = (int)(((unsigned long long)r[0] * (unsigned long long)ncorr) >> 32);
            const int i1 = (int)(((unsigned long long)r[1] * (unsigned long long)ncorr) >> 32);
            const int i2 = (int)(((unsigned long long)r[2] * (unsigned long long)ncorr) >> 32);
            double a0[3], a1[3], a2[3], b0[3], b1[3], b2[3];
            {
                const double2 u0 = __ldg(rec + (size_t)i0 * 4), u1 = __ldg(rec + (size_t)i0 * 4 + 1), u2 = __ldg(rec + (size_t)i0 * 4 + 2);
                const double2 v0 = __ldg(rec + (size_t)i1 * 4), v1 = __ldg(rec + (size_t)i1 * 4 + 1), v2 = __ldg(rec + (size_t)i1 * 4 + 2);
                const double2 w0 = __ldg(rec + (size_t)i2 * 4), w1 = __ldg(rec + (size_t)i2 * 4 + 1), w2 = __ldg(rec + (size_t)i2 * 4 + 2);
                a0[0] = u0.x; a0[1] = u0.y; a0[2] = u1.x; b0[0] = u1.y; b0[1] = u2.x; b0[2] = u2.y;
                a1[0] = v0.x; a1[1] = v0.y; a1[2] = v1.x; b1[0] = v1.y; b1[1] = v2.x; b1[2] = v2.y;
                a2[0] = w0.x; a2[1] = w0.y; a2[2] = w1.x; b2[0] = w1.y; b2[1] = w2.x; b2[2] = w2.y;
            }
            const double A0 = sqlen3(a0, a1), A1 = sqlen3(a1, a2), A2 = sqlen3(a2, a0);
            const double B0 = sqlen3(b0, b1), B1 = sqlen3(b1, b2), B2 = sqlen3(b2, b0);
            // fi = the larger cloud: li on the source unless the target is larger (swap)
            flag = swapped ? (tuple_edge_ok(B0, A0, ts, ts2_hi, ts2_lo) && tuple_edge_ok(B1, A1, ts, ts2_hi, ts2_lo) &&
                              tuple_edge_ok(B2, A2, ts, ts2_hi, ts2_lo))
                           : (tuple_edge_ok(A0, B0, ts, ts2_hi, ts2_lo) && tuple_edge_ok(A1, B1, ts, ts2_hi, ts2_lo) &&
                              tuple_edge_ok(A2, B2, ts, ts2_hi, ts2_lo));
        }
        const unsigned w = __ballot_sync(0xffffffffu, flag);
        if ((tid & 31) == 0) words[(t - epoch_base) >> 5] = w;
    }
}

__global__ void __launch_bounds__(FG_THREADS) fgr_tuple_pick_kernel(FgrCtx c, long long epoch_base)
{
    __shared__ int shi[FG_THREADS / 32];
    __shared__ int s_cnt;
    const int u = blockIdx.x, tid = threadIdx.x;
    FgrUnit &fu = c.unit[u];
    if (fu.tt_done) return;
    const int ncorr = fu.n_init;
    const int2 *in = c.corres_init + (size_t)u * c.cap;
    int2 *out = c.corres + (size_t)u * c.cap;
    if (!c.tuple_test) {
        for (int i = tid; i < ncorr; i += FG_THREADS) out[i] = in[i];
        __syncthreads();
        if (tid == 0) { fu.n_pairs = ncorr; fu.trials = 0; fu.tt_done = 1; }
        return;
    }
    const long long n_trials = (long long)ncorr * 100;
    const long long epoch_end = epoch_base + c.tt_epoch < n_trials ? epoch_base + c.tt_epoch : n_trials;
    const int n_words = epoch_end > epoch_base ? (int)((epoch_end - epoch_base + 31) >> 5) : 0;
    const unsigned *words = c.tt_flags + (size_t)u * (c.tt_epoch / 32);
    const unsigned long long key = c.keys[u];
    if (tid == 0) s_cnt = fu.tt_cnt;
    __syncthreads();
    bool full = false;
    for (int wb = 0; wb < n_words; wb += FG_THREADS) {
        const int wi = wb + tid;
        unsigned w = wi < n_words ? words[wi] : 0u;
        int total;
        const int off = block_scan_count(__popc(w), shi, total);
        const int cnt0 = s_cnt;
        int pos = cnt0 + off;                 // trial order: word-major, then the bits of a word
        while (w && pos < c.max_tuple) {
            const int bit = __ffs(w) - 1;
            w &= w - 1;
            const long long t = epoch_base + (long long)wi * 32 + bit;
            uint32_t r[4];
            philox4x32_10((unsigned long long)t, key, r);
            int2 *o = out + (size_t)pos * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                o[k] = in[(int)(((unsigned long long)r[k] * (unsigned long long)ncorr) >> 32)];
            // trials the sequential loop would have run: up to and including the one that fills the quota
            if (pos == c.max_tuple - 1) fu.trials = (int)(t + 1);
            ++pos;
        }
        __syncthreads();
        if (cnt0 + total >= c.max_tuple) { full = true; break; }
        if (tid == 0) s_cnt = cnt0 + total;
        __syncthreads();
    }
    if (tid == 0) {
        if (full) {
            fu.tt_cnt = c.max_tuple; fu.n_pairs = c.max_tuple * 3; fu.tt_done = 1;
        } else {
            fu.tt_cnt = s_cnt;
            if (epoch_end >= n_trials) { fu.n_pairs = s_cnt * 3; fu.trials = (int)n_trials; fu.tt_done = 1; }
        }
    }
}

// OptimizePairwiseRegistration (graduated non-convexity, Geman-McClure weights) and
// GetTransformationOriginalScale(...).inverse(), one block per unit.
__global__ void __launch_bounds__(FG_THREADS) fgr_gnc_kernel(FgrCtx c)
{
    __shared__ double sh_part[FG_THREADS / 32][28];
    __shared__ double sh_tot[28];
    __shared__ double sh_delta[12];
    __shared__ double sh_A[36], sh_x[6], sh_sc[6];
    __shared__ int sh_perm[6];
    const int u = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const FgrUnit fu = c.unit[u];
    const int ncorr = fu.n_pairs;
    const int2 *cor = c.corres + (size_t)u * c.cap;
    double *qc = c.qcur + (size_t)u * c.cap * 3;
    double *pc = c.pcur + (size_t)u * c.cap * 3;   // normalised source point of every pair (constant over the iterations)
    const double *sp = c.src_pts + (size_t)c.unit_src[u] * c.n_src * 3;
    const double *tp = c.tgt_pts + (size_t)c.unit_tgt[u] * c.n_tgt * 3;
    double trans[16];
    for (int k = 0; k < 16; ++k) trans[k] = (k % 5 == 0) ? 1.0 : 0.0;
    if (ncorr >= 10) {
        for (int i = tid; i < ncorr; i += FG_THREADS) {
            double q[3];
            norm_point(tp, cor[i].y, fu.mean_t, fu.scale, q);
            qc[(size_t)i * 3] = q[0]; qc[(size_t)i * 3 + 1] = q[1]; qc[(size_t)i * 3 + 2] = q[2];
            norm_point(sp, cor[i].x, fu.mean_s, fu.scale, q);
            pc[(size_t)i * 3] = q[0]; pc[(size_t)i * 3 + 1] = q[1]; pc[(size_t)i * 3 + 2] = q[2];
        }
        __syncthreads();
        double par = 1.0;   // scale_start (use_absolute_scale = false)
        for (int itr = 0; itr < c.iteration_number; ++itr) {
            double acc[27];
#pragma unroll
            for (int k = 0; k < 27; ++k) acc[k] = 0.0;
#pragma unroll 4
            for (int i = tid; i < ncorr; i += FG_THREADS) {
                const double p[3] = {pc[(size_t)i * 3], pc[(size_t)i * 3 + 1], pc[(size_t)i * 3 + 2]};
                const double qx = qc[(size_t)i * 3], qy = qc[(size_t)i * 3 + 1], qz = qc[(size_t)i * 3 + 2];
                const double rx = p[0] - qx, ry = p[1] - qy, rz = p[2] - qz;
                const double temp = par / ((rx * rx + ry * ry + rz * rz) + par);
                const double s = temp * temp;
                // rows of J: (0, -qz, qy, -1, 0, 0), (qz, 0, -qx, 0, -1, 0), (-qy, qx, 0, 0, 0, -1).  Every entry of
                // J^T J / J^T r receives `J[r][a] * J[r][b] * s` of rows 0, 1, 2 in that order; a term with a zero
                // factor is an exact (signed) zero and leaves the sum unchanged, so only the 18 + 9 non-zero terms
                // are formed -- same values, same order, a third of the float64 instructions (with -fmad=false the
                // compiler may not drop `0.0 * x` itself: the kernel spent 240 multiplies and adds per pair here).
                const double nqz = -qz, nqx = -qx, nqy = -qy;
                // upper-triangle index of (a, b): k(a, b) = a * 6 - a * (a - 1) / 2 + (b - a)
                // row 0: non-zeros at 1 (-qz), 2 (qy), 3 (-1)
                acc[6] += nqz * nqz * s;  acc[7] += nqz * qy * s;   acc[8] += nqz * -1.0 * s;
                acc[11] += qy * qy * s;   acc[12] += qy * -1.0 * s; acc[15] += -1.0 * -1.0 * s;
                acc[22] += nqz * rx * s;  acc[23] += qy * rx * s;   acc[24] += -1.0 * rx * s;
                // row 1: non-zeros at 0 (qz), 2 (-qx), 4 (-1)
                acc[0] += qz * qz * s;    acc[2] += qz * nqx * s;   acc[4] += qz * -1.0 * s;
                acc[11] += nqx * nqx * s; acc[13] += nqx * -1.0 * s; acc[18] += -1.0 * -1.0 * s;
                acc[21] += qz * ry * s;   acc[23] += nqx * ry * s;  acc[25] += -1.0 * ry * s;
                // row 2: non-zeros at 0 (-qy), 1 (qx), 5 (-1)
                acc[0] += nqy * nqy * s;  acc[1] += nqy * qx * s;   acc[5] += nqy * -1.0 * s;
                acc[6] += qx * qx * s;    acc[10] += qx * -1.0 * s; acc[20] += -1.0 * -1.0 * s;
                acc[21] += nqy * rz * s;  acc[22] += qx * rz * s;   acc[26] += -1.0 * rz * s;
            }
#pragma unroll
            for (int k = 0; k < 27; ++k) acc[k] = warp_sum(acc[k]);
            __syncthreads();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 27; ++k) sh_part[wid][k] = acc[k];
            }
            __syncthreads();
            if (tid < 27) {   // 27 threads add the warps' partials (same order as one thread would)
                double s = 0.0;
                for (int w = 0; w < FG_THREADS / 32; ++w) s += sh_part[w][tid];
                sh_tot[tid] = s;
            }
            __syncthreads();
            if (wid == 0) {      // warp 0 solves (-J^T J) x = J^T r cooperatively (ldlt_solve6_warp, common.cuh)
                if (lane == 0) {
                    int k = 0;
                    for (int a = 0; a < 6; ++a)
                        for (int b = a; b < 6; ++b) { sh_A[a * 6 + b] = -sh_tot[k]; sh_A[b * 6 + a] = -sh_tot[k]; ++k; }
                }
                __syncwarp();
                ldlt_solve6_warp(sh_A, sh_tot + 21, sh_x, sh_perm, lane);
                if (lane < 3) sincos(sh_x[lane], &sh_sc[lane], &sh_sc[3 + lane]);
                __syncwarp();
                if (lane == 0) {
                    const double sa = sh_sc[0], sb = sh_sc[1], sg = sh_sc[2], ca = sh_sc[3], cb = sh_sc[4], cg = sh_sc[5];
                    const double *x = sh_x;
                    sh_delta[0] = cg * cb; sh_delta[1] = cg * sb * sa - sg * ca; sh_delta[2] = cg * sb * ca + sg * sa; sh_delta[3] = x[3];
                    sh_delta[4] = sg * cb; sh_delta[5] = sg * sb * sa + cg * ca; sh_delta[6] = sg * sb * ca - cg * sa; sh_delta[7] = x[4];
                    sh_delta[8] = -sb;     sh_delta[9] = cb * sa;                sh_delta[10] = cb * ca;               sh_delta[11] = x[5];
                }
            }
            __syncthreads();
            double D[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) D[k] = sh_delta[k];
            // trans = delta * trans  (every thread keeps its own copy; only thread 0's is used)
            double Tn[12];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    Tn[r * 4 + cc] = D[r * 4] * trans[cc] + D[r * 4 + 1] * trans[4 + cc] + D[r * 4 + 2] * trans[8 + cc] +
                                     D[r * 4 + 3] * trans[12 + cc];
#pragma unroll
            for (int k = 0; k < 12; ++k) trans[k] = Tn[k];
            for (int i = tid; i < ncorr; i += FG_THREADS) {
                double x = qc[(size_t)i * 3], y = qc[(size_t)i * 3 + 1], z = qc[(size_t)i * 3 + 2];
                transform_point(D, x, y, z);
                qc[(size_t)i * 3] = x; qc[(size_t)i * 3 + 1] = y; qc[(size_t)i * 3 + 2] = z;
            }
            if (c.decrease_mu && itr % 4 == 0 && par > c.max_corr_dist) par /= c.division_factor;
        }
    }
    if (tid != 0) return;
    // t' = -R mean_t + t * scale + mean_s ; result = [R | t']^-1 = [R^T | -R^T t']
    double tt[3];
    for (int r = 0; r < 3; ++r)
        tt[r] = -(trans[r * 4] * fu.mean_t[0] + trans[r * 4 + 1] * fu.mean_t[1] + trans[r * 4 + 2] * fu.mean_t[2]) +
                trans[r * 4 + 3] * fu.scale + fu.mean_s[r];
    double *T = c.out_T + (size_t)u * 16;
    for (int r = 0; r < 3; ++r) {
        for (int cc = 0; cc < 3; ++cc) T[r * 4 + cc] = trans[cc * 4 + r];
        T[r * 4 + 3] = -(trans[0 * 4 + r] * tt[0] + trans[1 * 4 + r] * tt[1] + trans[2 * 4 + r] * tt[2]);
    }
    T[12] = T[13] = T[14] = 0.0; T[15] = 1.0;
    int *info = c.out_info + (size_t)u * 4;
    info[0] = fu.n_init; info[1] = fu.n_pairs; info[2] = fu.trials; info[3] = fu.swapped;
}

struct FgrLayout {
    size_t off_unit, off_init, off_cor, off_q, off_p, off_flags, off_pairs, total;
    int cap, tt_epoch, tt_epochs;
};

bool fgr_layout(const orpcd_fgr_params *p, FgrLayout &L)
{
    if (!p || p->n_units <= 0 || p->n_src <= 0 || p->n_tgt <= 0 || p->n_feat_src <= 0 || p->n_feat_tgt <= 0 ||
        p->iteration_number < 0 || p->max_tuple_count <= 0 || !(p->tuple_scale > 0) || !(p->division_factor > 0))
        return false;
    const long long tup = 3LL * p->max_tuple_count;
    const long long cap = tup > p->n_feat_src ? tup : p->n_feat_src;
    if (cap > 0x3fffffff) return false;
    L.cap = (int)cap;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += orpcd_align_up(bytes, 256); return r; };
    const size_t U = p->n_units;
    L.off_unit = take(U * sizeof(FgrUnit));
    L.off_init = take(U * L.cap * sizeof(int2));
    L.off_cor = take(U * L.cap * sizeof(int2));
    L.off_q = take(U * L.cap * 3 * sizeof(double));
    L.off_p = take(U * L.cap * 3 * sizeof(double));
    // tuple-test epochs: at most 100 x n_feat_src trials per unit, an eighth of them per epoch (a unit usually
    // fills its quota late: 139k of 167k possible trials on the 5000-point bench clouds), 64k .. 1M trials
    const long long max_trials = 100LL * p->n_feat_src;
    long long ep = 65536;
    while (ep * 8 < max_trials && ep < (1LL << 20)) ep *= 2;
    L.tt_epoch = (int)ep;
    L.tt_epochs = p->tuple_test ? (int)((max_trials + ep - 1) / ep) : 1;
    L.off_flags = take(U * (size_t)(ep / 32) * sizeof(unsigned));
    L.off_pairs = take(U * L.cap * 8 * sizeof(double));
    L.total = o;
    return true;
}

}  // namespace

ORPCD_API size_t orpcd_feature_pack_bytes(int n)
{
    if (n <= 0) return 0;
    return (size_t)((n + FM_ROWS - 1) / FM_ROWS) * FM_SET_BYTES_PER_TILE;
}

ORPCD_API int orpcd_feature_center(const double *d_features, int n, double *d_center, orpcd_stream_t stream)
{
    if (!d_features || !d_center || n <= 0) return ORPCD_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    feat_center_kernel<<<1, 1023, 0, st>>>(d_features, n, d_center);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API int orpcd_feature_pack(const double *d_features, int n, int n_sets, const double *d_center, void *d_packed,
                                 float *d_set_maxnorm, orpcd_stream_t stream)
{
    if (!d_features || !d_packed || !d_set_maxnorm || n <= 0 || n_sets <= 0) return ORPCD_ERR_INVALID;
    if ((uintptr_t)d_packed & 127) return ORPCD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(d_set_maxnorm, 0, sizeof(float) * n_sets, st) != cudaSuccess) return ORPCD_ERR_CUDA;
    const int tiles = (n + FM_ROWS - 1) / FM_ROWS;
    feat_pack_kernel<<<dim3(tiles, n_sets), FM_ROWS, 0, st>>>(d_features, n, tiles, d_center, (unsigned char *)d_packed,
                                                             d_set_maxnorm);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API size_t orpcd_feature_match_workspace_bytes(int n_query, int n_pairs)
{
    if (n_query <= 0 || n_pairs <= 0) return 0;
    return (2 * FM_TOP + 1) * orpcd_align_up((size_t)n_query * n_pairs * 4, 256);   // FM_TOP + 1 chunk minima, FM_TOP chunk ids per row
}

ORPCD_API int orpcd_feature_match(const double *d_feat_q, const void *d_packed_q, int n_query,
                                  const double *d_feat_db, const void *d_packed_db, const float *d_db_maxnorm,
                                  int n_db, const double *d_center, const int *d_pair_q, const int *d_pair_db,
                                  int n_pairs, int mode,
                                  void *d_workspace, size_t workspace_bytes, const unsigned char *d_row_mask,
                                  int *d_nn, double *d_nn_d2,
                                  int *d_stats, float *d_debug_v, orpcd_stream_t stream)
{
    if (!d_feat_q || !d_packed_q || !d_feat_db || !d_packed_db || !d_db_maxnorm || !d_pair_q || !d_pair_db || !d_nn)
        return ORPCD_ERR_INVALID;
    if (n_query <= 0 || n_db <= 0 || n_pairs <= 0 || n_pairs > 65535 || mode < 0 || mode > 1) return ORPCD_ERR_INVALID;
    if (!d_workspace || workspace_bytes < orpcd_feature_match_workspace_bytes(n_query, n_pairs) ||
        ((uintptr_t)d_workspace & 255))
        return ORPCD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t slice = orpcd_align_up((size_t)n_query * n_pairs * 4, 256);
    float *fm = (float *)d_workspace;                                                // [FM_TOP + 1][pairs * n_query]
    int *fc = (int *)((unsigned char *)d_workspace + (size_t)(FM_TOP + 1) * slice);  // [FM_TOP][pairs * n_query]
    const size_t stride = slice / 4;
    const int q_tiles = (n_query + FM_ROWS - 1) / FM_ROWS, db_tiles = (n_db + FM_ROWS - 1) / FM_ROWS;
    if (mode == 0) {
        auto kern = d_debug_v ? feat_match_kernel<true> : feat_match_kernel<false>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FM_SMEM_BYTES) != cudaSuccess)
            return ORPCD_ERR_CUDA;
        kern<<<dim3((q_tiles + FM_ATILES - 1) / FM_ATILES, n_pairs), FM_THREADS, FM_SMEM_BYTES, st>>>(
            (const unsigned char *)d_packed_q, q_tiles, (const unsigned char *)d_packed_db, db_tiles, d_pair_q,
            d_pair_db, n_query, fm, fc, stride, d_debug_v);
        ORPCD_CHECK_LAUNCH();
    }
    if (d_stats && cudaMemsetAsync(d_stats, 0, 2 * sizeof(int), st) != cudaSuccess) return ORPCD_ERR_CUDA;
    feat_resolve_kernel<<<dim3((n_query + 3) / 4, n_pairs), 128, 0, st>>>(d_feat_q, n_query,
                                                                          (const unsigned char *)d_packed_db, db_tiles,
                                                                          n_db, d_center, d_pair_q,
                                                                          d_pair_db, d_db_maxnorm, fm, fc, stride,
                                                                          mode == 1, d_row_mask, d_nn, d_nn_d2, d_stats);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

namespace {

// tuple_key(seed, source, target) of Optimizer/fastGlobalOptimizer.py: a hash of the seed, the two cloud sizes and
// the bit patterns of the first and last point of each cloud (one thread per unit).
__global__ void fgr_tuple_keys_kernel(const double *__restrict__ src, const double *__restrict__ tgt, int n_src,
                                      int n_tgt, const int *__restrict__ unit_src, const int *__restrict__ unit_tgt,
                                      int n_units, unsigned long long seed, unsigned long long *__restrict__ keys)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const double *s = src + (size_t)unit_src[u] * n_src * 3, *t = tgt + (size_t)unit_tgt[u] * n_tgt * 3;
    unsigned long long h = seed * 0x9E3779B97F4A7C15ull + 0x243F6A8885A308D3ull;
    auto mix = [&](unsigned long long w) {
        h = (h ^ w) * 0xBF58476D1CE4E5B9ull;
        h ^= h >> 31;
    };
    mix((unsigned long long)n_src);
    mix((unsigned long long)n_tgt);
    const double *ends[4] = {s, s + (size_t)(n_src - 1) * 3, t, t + (size_t)(n_tgt - 1) * 3};
    for (int e = 0; e < 4; ++e)
        for (int k = 0; k < 3; ++k) mix((unsigned long long)__double_as_longlong(ends[e][k]));
    keys[u] = h;
}

}  // namespace

ORPCD_API int orpcd_fgr_tuple_keys(const double *d_src_points, int n_src, const double *d_tgt_points, int n_tgt,
                                   const int *d_unit_src, const int *d_unit_tgt, int n_units,
                                   unsigned long long seed, unsigned long long *d_keys, orpcd_stream_t stream)
{
    if (!d_src_points || !d_tgt_points || !d_unit_src || !d_unit_tgt || !d_keys || n_units <= 0 || n_src <= 0 ||
        n_tgt <= 0)
        return ORPCD_ERR_INVALID;
    fgr_tuple_keys_kernel<<<(n_units + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        d_src_points, d_tgt_points, n_src, n_tgt, d_unit_src, d_unit_tgt, n_units, seed, d_keys);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API size_t orpcd_fgr_batch_workspace_bytes(const orpcd_fgr_params *p)
{
    FgrLayout L;
    if (!fgr_layout(p, L)) return 0;
    return L.total;
}

ORPCD_API int orpcd_fgr_batch(const orpcd_fgr_params *p, const double *d_src_points, const double *d_tgt_points,
                              const int *d_unit_src, const int *d_unit_tgt, const int *d_ij, const int *d_ji,
                              const unsigned long long *d_tuple_keys, void *d_workspace, size_t workspace_bytes,
                              double *d_out_T, int *d_out_info, orpcd_stream_t stream)
{
    FgrLayout L;
    if (!fgr_layout(p, L)) return ORPCD_ERR_INVALID;
    if (!d_src_points || !d_tgt_points || !d_unit_src || !d_unit_tgt || !d_ij || !d_ji || !d_tuple_keys || !d_out_T ||
        !d_out_info)
        return ORPCD_ERR_INVALID;
    if (!d_workspace || workspace_bytes < L.total || ((uintptr_t)d_workspace & 255)) return ORPCD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *w = (unsigned char *)d_workspace;
    FgrCtx c;
    c.n_units = p->n_units; c.n_src = p->n_src; c.n_tgt = p->n_tgt;
    c.n_feat_src = p->n_feat_src; c.n_feat_tgt = p->n_feat_tgt; c.cap = L.cap;
    c.iteration_number = p->iteration_number; c.decrease_mu = p->decrease_mu;
    c.max_tuple = p->max_tuple_count; c.tuple_test = p->tuple_test;
    c.division_factor = p->division_factor; c.tuple_scale = p->tuple_scale; c.max_corr_dist = p->max_correspondence_distance;
    c.src_pts = d_src_points; c.tgt_pts = d_tgt_points; c.unit_src = d_unit_src; c.unit_tgt = d_unit_tgt;
    c.ij = d_ij; c.ji = d_ji; c.keys = d_tuple_keys;
    c.unit = (FgrUnit *)(w + L.off_unit);
    c.corres_init = (int2 *)(w + L.off_init);
    c.corres = (int2 *)(w + L.off_cor);
    c.qcur = (double *)(w + L.off_q);
    c.pcur = (double *)(w + L.off_p);
    c.out_T = d_out_T; c.out_info = d_out_info;
    c.tt_flags = (unsigned *)(w + L.off_flags);
    c.tt_pairs = (double *)(w + L.off_pairs);
    c.tt_epoch = L.tt_epoch;
    fgr_prepare_kernel<<<p->n_units, FG_THREADS, 0, st>>>(c);
    for (int e = 0; e < L.tt_epochs; ++e) {
        const long long base = (long long)e * L.tt_epoch;
        if (p->tuple_test)
            fgr_tuple_eval_kernel<<<dim3(L.tt_epoch / TT_BLOCK_TRIALS, p->n_units), 256, 0, st>>>(c, base);
        fgr_tuple_pick_kernel<<<p->n_units, FG_THREADS, 0, st>>>(c, base);
    }
    fgr_gnc_kernel<<<p->n_units, FG_THREADS, 0, st>>>(c);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}
