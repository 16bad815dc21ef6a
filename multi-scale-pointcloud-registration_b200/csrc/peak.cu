// FP32 FMA-pipe micro-benchmarks: the roofline denominator for the 1-NN sweep
// (BASELINE.md section 4 asks for a measured FFMA peak from the same run).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) ffma_burn_kernel(int iters, float *__restrict__ sink)
{
    float a[8];
    const float x = 1.0f + 1e-7f * threadIdx.x, y = 1e-3f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = (float)k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = __fmaf_rn(a[k], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) ffma2_burn_kernel(int iters, float *__restrict__ sink)
{
    uint64_t a[8];
    const float x = 1.0f + 1e-7f * threadIdx.x, y = 1e-3f;
    uint64_t xx, yy;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("mov.b64 %0, {%1, %1};" : "=l"(yy) : "f"(y));
#pragma unroll
    for (int k = 0; k < 8; ++k) asm("mov.b64 %0, {%1, %2};" : "=l"(a[k]) : "f"((float)k), "f"((float)k + 0.5f));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[k]) : "l"(xx), "l"(yy));
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[k]));
        s += lo + hi;
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_burn_kernel(int iters, double *__restrict__ sink)
{
    double a[8];
    const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-3;
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = (double)k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = __fma_rn(a[k], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Issue-slot probe: 8 FFMA2 chains interleaved with `alu_per_fma` independent
// FMNMX ops per FFMA2 (ALU pipe), to see whether FFMA2 costs one or two issue slots.
template <int ALU>
__global__ void __launch_bounds__(256) mix_burn_kernel(int iters, float *__restrict__ sink)
{
    uint64_t a[8];
    float m[8];
    const float x = 1.0f + 1e-7f * threadIdx.x, y = 1e-3f;
    uint64_t xx, yy;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("mov.b64 %0, {%1, %1};" : "=l"(yy) : "f"(y));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        asm("mov.b64 %0, {%1, %2};" : "=l"(a[k]) : "f"((float)k), "f"((float)k + 0.5f));
        m[k] = 1e30f - k;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[k]) : "l"(xx), "l"(yy));
#pragma unroll
                for (int q = 0; q < ALU; ++q)
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[(k + q) & 7]) : "f"(x), "f"(y));
            }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[k]));
        s += lo + hi + m[k];
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

ORPCD_API int orpcd_mix_burn(int blocks, int iters, int alu_per_fma, float *d_sink, orpcd_stream_t stream)
{
    if (blocks <= 0 || iters <= 0 || !d_sink) return ORPCD_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (alu_per_fma == 0) mix_burn_kernel<0><<<blocks, 256, 0, st>>>(iters, d_sink);
    else if (alu_per_fma == 1) mix_burn_kernel<1><<<blocks, 256, 0, st>>>(iters, d_sink);
    else if (alu_per_fma == 2) mix_burn_kernel<2><<<blocks, 256, 0, st>>>(iters, d_sink);
    else return ORPCD_ERR_INVALID;
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API int orpcd_dfma_burn(int blocks, int iters, double *d_sink, orpcd_stream_t stream)
{
    if (blocks <= 0 || iters <= 0 || !d_sink) return ORPCD_ERR_INVALID;
    dfma_burn_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, d_sink);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API int orpcd_ffma_burn(int blocks, int iters, float *d_sink, orpcd_stream_t stream)
{
    if (blocks <= 0 || iters <= 0 || !d_sink) return ORPCD_ERR_INVALID;
    ffma_burn_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, d_sink);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API int orpcd_ffma2_burn(int blocks, int iters, float *d_sink, orpcd_stream_t stream)
{
    if (blocks <= 0 || iters <= 0 || !d_sink) return ORPCD_ERR_INVALID;
    ffma2_burn_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, d_sink);
    ORPCD_CHECK_LAUNCH();
    return ORPCD_OK;
}

ORPCD_API int orpcd_version(void) { return 100; }

ORPCD_API int orpcd_compiled_arch(void)
{
    return 100;  // built with -gencode arch=compute_100a,code=sm_100a only
}

ORPCD_API const char *orpcd_error_string(int code)
{
    switch (code) {
        case ORPCD_OK: return "ok";
        case ORPCD_ERR_INVALID: return "invalid argument";
        case ORPCD_ERR_WORKSPACE: return "workspace too small or misaligned";
        case ORPCD_ERR_CUDA: return "CUDA runtime/launch failure";
        default: return "unknown error";
    }
}
