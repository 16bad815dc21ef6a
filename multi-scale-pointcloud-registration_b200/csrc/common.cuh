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

// ----------------------------------------------------------------------------
// Deterministic acos / cos (bit-identical twins of det_acos / det_cos in
// oracle/orpcd_oracle.c; Sun fdlibm's e_acos.c / k_cos.c / k_sin.c algorithms restated
// with + - * / sqrt and bit masks only, < 1 ulp).  Open3D calls libm here and feeds the
// results into discontinuous decisions (FastEigen3x3's branches, the source/target role
// swap of FPFH's pair features); the CUDA math library and glibc differ in the last ulp,
// and range-grid scans hold structural near-ties, so both sides of the parity tests use
// this arithmetic.  Requires -fmad=false (no FMA contraction).
// ----------------------------------------------------------------------------
__device__ __forceinline__ double det_acos(double x)
{
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17,
                 pi = 3.14159265358979311600e+00;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    const double ax = fabs(x);
    if (ax >= 1.0) {
        if (ax == 1.0) return x > 0 ? 0.0 : pi + 2.0 * pio2_lo;
        return (x - x) / (x - x);
    }
    if (ax < 0.5) {
        if (ax < 6.938893903907228e-18) return pio2_hi + pio2_lo;   // 2^-57
        const double z = x * x;
        const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const double r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    } else if (x < 0) {
        const double z = (1.0 + x) * 0.5;
        const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const double s = sqrt(z);
        const double r = p / q;
        const double w = r * s - pio2_lo;
        return pi - 2.0 * (s + w);
    } else {
        const double z = (1.0 - x) * 0.5;
        const double s = sqrt(z);
        const double df = __hiloint2double(__double2hiint(s), 0);
        const double c = (z - df * df) / (s + df);
        const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const double r = p / q;
        const double w = r * s + c;
        return 2.0 * (df + w);
    }
}

__device__ __forceinline__ double det_kcos(double x, double y)
{
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double ax = fabs(x);
    if (ax < 7.450580596923828e-09) return 1.0;   // 2^-27
    const double z = x * x;
    const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    if (ax < 0.3) return 1.0 - (0.5 * z - (z * r - x * y));
    const double qx = ax > 0.78125 ? 0.28125 : __hiloint2double(__double2hiint(ax) - 0x00200000, 0);
    const double hz = 0.5 * z - qx;
    const double a = 1.0 - qx;
    return a - (hz - (z * r - x * y));
}

__device__ __forceinline__ double det_ksin(double x, double y)
{
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = x * x;
    const double v = z * x;
    const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

// cos(x) for 0 <= x <= pi
__device__ __forceinline__ double det_cos(double x)
{
    const double pio4 = 7.85398163397448278999e-01, three_pio4 = 2.35619449019234483700e+00;
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
    const double pi_hi = 3.14159265358979311600e+00, pi_lo = 1.22464679914735317723e-16;
    if (x <= pio4) return det_kcos(x, 0.0);
    if (x < three_pio4) {
        const double t = x - pio2_hi;
        const double y = t - pio2_lo;
        const double yt = (t - y) - pio2_lo;
        return -det_ksin(y, yt);
    }
    const double t = x - pi_hi;
    const double y = t - pi_lo;
    const double yt = (t - y) - pi_lo;
    return -det_kcos(y, yt);
}

// atan / atan2: bit-identical twins of det_atan / det_atan2 in oracle/orpcd_oracle.c (fdlibm's s_atan.c /
// e_atan2.c restated).  Feeds the floor() that bins FPFH's f0 angle (Feature.cpp ComputePairFeatures).
__device__ __forceinline__ double det_atan(double x)
{
    const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01, aT2 = 1.42857142725034663711e-01,
                 aT3 = -1.11111104054623557880e-01, aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
                 aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02, aT8 = 4.97687799461593236017e-02,
                 aT9 = -3.65315727442169155270e-02, aT10 = 1.62858201153657823623e-02;
    const unsigned hx = (unsigned)__double2hiint(x), ix = hx & 0x7fffffffu;
    double hi = 0.0, lo = 0.0;
    bool reduced = true;
    if (ix >= 0x44100000u) {                        // |x| >= 2^66 (or NaN)
        if (x != x) return x + x;
        const double r = 1.57079632679489655800e+00 + 6.12323399573676603587e-17;
        return (hx >> 31) ? -1.57079632679489655800e+00 - 6.12323399573676603587e-17 : r;
    }
    if (ix < 0x3fdc0000u) {                         // |x| < 0.4375
        if (ix < 0x3e200000u) return x;             // |x| < 2^-29
        reduced = false;
    } else {
        x = fabs(x);
        if (ix < 0x3ff30000u) {
            if (ix < 0x3fe60000u) {
                hi = 4.63647609000806093515e-01; lo = 2.26987774529616870924e-17;
                x = (2.0 * x - 1.0) / (2.0 + x);
            } else {
                hi = 7.85398163397448278999e-01; lo = 3.06161699786838301793e-17;
                x = (x - 1.0) / (x + 1.0);
            }
        } else {
            if (ix < 0x40038000u) {
                hi = 9.82793723247329054082e-01; lo = 1.39033110312309984516e-17;
                x = (x - 1.5) / (1.0 + 1.5 * x);
            } else {
                hi = 1.57079632679489655800e+00; lo = 6.12323399573676603587e-17;
                x = -1.0 / x;
            }
        }
    }
    const double z = x * x;
    const double w = z * z;
    const double s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const double s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (!reduced) return x - x * (s1 + s2);
    const double r = hi - ((x * (s1 + s2) - lo) - x);
    return (hx >> 31) ? -r : r;
}

__device__ __forceinline__ double det_atan2(double y, double x)
{
    const double pi_o_2 = 1.5707963267948965580e+00, pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16;
    if (x != x || y != y) return x + y;
    if (x == 1.0) return det_atan(y);
    const unsigned hx = (unsigned)__double2hiint(x), hy = (unsigned)__double2hiint(y);
    const int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);
    if (y == 0.0) return m < 2 ? y : (m == 2 ? pi : -pi);
    if (x == 0.0) return (hy >> 31) ? -pi_o_2 : pi_o_2;
    const int ix = (int)(hx & 0x7fffffffu), iy = (int)(hy & 0x7fffffffu);
    if (ix >= 0x7ff00000 || iy >= 0x7ff00000) {
        if (iy >= 0x7ff00000 && ix < 0x7ff00000) return (hy >> 31) ? -pi_o_2 : pi_o_2;
        if (ix >= 0x7ff00000 && iy < 0x7ff00000) return m == 0 ? 0.0 : (m == 1 ? -0.0 : (m == 2 ? pi : -pi));
        return m == 0 ? 0.5 * pi_o_2 : (m == 1 ? -0.5 * pi_o_2 : (m == 2 ? 1.5 * pi_o_2 : -1.5 * pi_o_2));
    }
    const int k = (iy - ix) >> 20;
    double z;
    if (k > 60) z = pi_o_2 + 0.5 * pi_lo;
    else if ((hx >> 31) && k < -60) z = 0.0;
    else z = det_atan(fabs(y / x));
    if (m == 0) return z;
    if (m == 1) return -z;
    if (m == 2) return pi - (z - pi_lo);
    return (z - pi_lo) - pi;
}

// x = A.ldlt().solve(b) for a symmetric 6x6 system, Eigen's LDLT with diagonal pivoting restated, run by a whole warp:
// A (36 doubles, row-major, full), b, x (6 doubles each) and sperm (6 ints) live in shared memory.  The pivot search is
// a shuffle arg-max (lowest index on ties), the 15 trailing updates of a step run on 15 lanes, the substitutions on
// lane 0.  Every entry sees exactly the operations of the single-thread version (ldlt_solve6 in gicp.cu, used by the
// unfused reduce_kernel) in the same order: bit-identical results at a tenth of the latency (that one keeps the
// matrix in local memory: ~19 us per solve, every access a dependent L1 round trip).
__device__ __forceinline__ void ldlt_solve6_warp(double *A, const double *b, double *x, int *sperm, int lane)
{
    int perm_mine = lane;                                    // lanes 0..5: perm[lane]
    const int pi_ = 1 + lane / 5, pj_ = 1 + lane % 5;        // lanes 0..24: the pair (i, j) of the trailing update
    for (int k = 0; k < 6; ++k) {
        double v = (lane >= k && lane < 6) ? fabs(A[lane * 6 + lane]) : -1.0;
        int idx = lane;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {                    // arg-max over lanes 0..7, lowest index on ties
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        const int piv = __shfl_sync(0xffffffffu, idx, 0);
        if (piv != k) {
            if (lane < 6) { const double t = A[k * 6 + lane]; A[k * 6 + lane] = A[piv * 6 + lane]; A[piv * 6 + lane] = t; }
            __syncwarp();
            if (lane < 6) { const double t = A[lane * 6 + k]; A[lane * 6 + k] = A[lane * 6 + piv]; A[lane * 6 + piv] = t; }
            __syncwarp();
            const int pk = __shfl_sync(0xffffffffu, perm_mine, k), pp = __shfl_sync(0xffffffffu, perm_mine, piv);
            if (lane == k) perm_mine = pp;
            else if (lane == piv) perm_mine = pk;
        }
        const double d = A[k * 6 + k];
        if (d == 0.0) continue;                              // (uniform)
        const bool act = lane < 25 && pj_ > k && pj_ <= pi_;     // k < j <= i < 6
        const bool colw = lane > k && lane < 6;                  // lane i writes L[i][k]
        double nv = 0.0, li = 0.0;
        if (act) {
            const double l = A[pi_ * 6 + k] / d;
            nv = A[pi_ * 6 + pj_] - l * A[pj_ * 6 + k];
        }
        if (colw) li = A[lane * 6 + k] / d;
        __syncwarp();                                        // column k is read before it is overwritten
        if (act) { A[pi_ * 6 + pj_] = nv; A[pj_ * 6 + pi_] = nv; }
        if (colw) A[lane * 6 + k] = li;
        __syncwarp();
    }
    if (lane < 6) sperm[lane] = perm_mine;
    __syncwarp();
    if (lane == 0) {
        double y[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) y[i] = b[sperm[i]];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < i; ++j) y[i] -= A[i * 6 + j] * y[j];
#pragma unroll
        for (int i = 0; i < 6; ++i) y[i] = A[i * 6 + i] != 0.0 ? y[i] / A[i * 6 + i] : 0.0;
#pragma unroll
        for (int i = 5; i >= 0; --i)
#pragma unroll
            for (int j = i + 1; j < 6; ++j) y[i] -= A[j * 6 + i] * y[j];
#pragma unroll
        for (int i = 0; i < 6; ++i) x[sperm[i]] = y[i];
    }
    __syncwarp();
}


