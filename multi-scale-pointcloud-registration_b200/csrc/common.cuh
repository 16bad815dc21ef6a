// Shared device helpers for the sm_100a registration kernels.
// All float64 arithmetic that must agree with a non-contracting CPU build is
// written with the *_rn intrinsics (the library is also compiled -fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/orpcd_b200.h"

#define ORPCD_API extern "C" __attribute__((visibility("default")))

#define ORPCD_CHECK_LAUNCH()                                \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return ORPCD_ERR_CUDA;      \
    } while (0)

static inline size_t orpcd_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ((dx*dx) + dy*dy) + dz*dz, nanoflann L2 order for dim 3.
__device__ __forceinline__ double sqdist3_rn(double qx, double qy, double qz, double px,
                                             double py, double pz)
{
    const double dx = __dsub_rn(qx, px), dy = __dsub_rn(qy, py), dz = __dsub_rn(qz, pz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// symmetric 3x3 stored as (xx, xy, xz, yy, yz, zz)
struct Sym3 {
    double xx, xy, xz, yy, yz, zz;
};

__device__ __forceinline__ Sym3 load_sym3(const double *p)
{
    const double2 a = *reinterpret_cast<const double2 *>(p);
    const double2 b = *reinterpret_cast<const double2 *>(p + 2);
    const double2 c = *reinterpret_cast<const double2 *>(p + 4);
    Sym3 s;
    s.xx = a.x; s.xy = a.y; s.xz = b.x; s.yy = b.y; s.yz = c.x; s.zz = c.y;
    return s;
}

__device__ __forceinline__ void store_sym3(double *p, const Sym3 &s)
{
    *reinterpret_cast<double2 *>(p) = make_double2(s.xx, s.xy);
    *reinterpret_cast<double2 *>(p + 2) = make_double2(s.xz, s.yy);
    *reinterpret_cast<double2 *>(p + 4) = make_double2(s.yz, s.zz);
}

// C' = (R C) R^T for a 3x4 row-major [R|t] (PointCloud::Transform on covariances).
__device__ __forceinline__ Sym3 rotate_sym3(const double *U, const Sym3 &C)
{
    const double c[9] = {C.xx, C.xy, C.xz, C.xy, C.yy, C.yz, C.xz, C.yz, C.zz};
    double rc[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            rc[r * 3 + k] = U[r * 4] * c[k] + U[r * 4 + 1] * c[3 + k] + U[r * 4 + 2] * c[6 + k];
    Sym3 o;
    o.xx = rc[0] * U[0] + rc[1] * U[1] + rc[2] * U[2];
    o.xy = rc[0] * U[4] + rc[1] * U[5] + rc[2] * U[6];
    o.xz = rc[0] * U[8] + rc[1] * U[9] + rc[2] * U[10];
    o.yy = rc[3] * U[4] + rc[4] * U[5] + rc[5] * U[6];
    o.yz = rc[3] * U[8] + rc[4] * U[9] + rc[5] * U[10];
    o.zz = rc[6] * U[8] + rc[7] * U[9] + rc[8] * U[10];
    return o;
}

// p' = [R|t] (x, y, z, 1), 3x4 row-major.
__device__ __forceinline__ void transform_point(const double *U, double &x, double &y, double &z)
{
    const double nx = U[0] * x + U[1] * y + U[2] * z + U[3];
    const double ny = U[4] * x + U[5] * y + U[6] * z + U[7];
    const double nz = U[8] * x + U[9] * y + U[10] * z + U[11];
    x = nx; y = ny; z = nz;
}
