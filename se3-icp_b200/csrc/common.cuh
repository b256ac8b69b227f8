// common.cuh — shared device helpers for the SE(3)-ICP kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#define SE3_WARP 32
#define SE3_FULL 0xffffffffu

namespace se3 {

// ---- error plumbing -------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define SE3_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t err__ = (call);                                                                \
        if (err__ != cudaSuccess) {                                                                \
            se3::set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
            return SE3ICP_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

#define SE3_TRY(call)                 \
    do {                              \
        int rc__ = (call);            \
        if (rc__ != 0) return rc__;   \
    } while (0)

// ---- exact (non-contracted) squared distance ----------------------------------------------------
// The oracle is built with -ffp-contract=off; keys that drive discrete decisions (kNN order,
// argmin) use explicit round-to-nearest mul/add so both sides see identical bits.
__device__ __forceinline__ double sqdist3(double ax, double ay, double az, double bx, double by, double bz) {
    double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
    double s = __dmul_rn(dx, dx);
    s = __dadd_rn(s, __dmul_rn(dy, dy));
    s = __dadd_rn(s, __dmul_rn(dz, dz));
    return s;
}

// ---- warp reductions --------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SE3_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(SE3_FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(SE3_FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_minf(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(SE3_FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_maxf(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(SE3_FULL, v, o));
    return v;
}

// lexicographic (d2, idx) argmin across the warp; every lane receives the winner
__device__ __forceinline__ void warp_argmin(double& d2, int& idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double od = __shfl_xor_sync(SE3_FULL, d2, o);
        int oi = __shfl_xor_sync(SE3_FULL, idx, o);
        if (od < d2 || (od == d2 && oi < idx)) {
            d2 = od;
            idx = oi;
        }
    }
}

#ifndef EIG_FAST_ROTATION
#define EIG_FAST_ROTATION 1
#endif
// ---- 3x3 symmetric eigen-solver (cyclic Jacobi, FP64) ----------------------------------------------
// a = {a00,a01,a02,a11,a12,a22}.  evals ascending; V columns are unit eigenvectors (V[r][c]).
__device__ inline void eig3_sym(const double a[6], double evals[3], double V[3][3]) {
    double A[3][3] = {{a[0], a[1], a[2]}, {a[1], a[3], a[4]}, {a[2], a[4], a[5]}};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 32; sweep++) {
        double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
        if (off == 0.0 || off <= 1e-32 * diag) break;
#pragma unroll
        for (int p = 0; p < 2; p++) {
#pragma unroll
            for (int q = p + 1; q < 3; q++) {
                double apq = A[p][q];
#if EIG_FAST_ROTATION
                // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)) with theta = d / apq, d = (aqq - app) / 2, written as
                // sgn(d) apq / (|d| + sqrt(d^2 + apq^2)): one square root, one division and one reciprocal square root on
                // the dependent chain instead of two divisions, two square roots and a reciprocal (the solve of a block is
                // a serial latency chain the other warps wait for)
                const double d = 0.5 * (A[q][q] - A[p][p]);
                const double h = sqrt(d * d + apq * apq);
                if (apq != 0.0 && h > 0.0) {
                    double t = apq / (fabs(d) + h);
                    if (d < 0.0) t = -t;
                    double c = rsqrt(t * t + 1.0), s = t * c;
#else
                if (apq != 0.0) {
                    double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                    double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#endif
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        double akp = A[k][p], akq = A[k][q];
                        A[k][p] = c * akp - s * akq;
                        A[k][q] = s * akp + c * akq;
                    }
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        double apk = A[p][k], aqk = A[q][k];
                        A[p][k] = c * apk - s * aqk;
                        A[q][k] = s * apk + c * aqk;
                    }
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        double vkp = V[k][p], vkq = V[k][q];
                        V[k][p] = c * vkp - s * vkq;
                        V[k][q] = s * vkp + c * vkq;
                    }
                }
            }
        }
    }
    // sort ascending (3 elements)
    double d0 = A[0][0], d1 = A[1][1], d2 = A[2][2];
    int i0 = 0, i1 = 1, i2 = 2;
    if (d1 < d0) { double t = d0; d0 = d1; d1 = t; int ti = i0; i0 = i1; i1 = ti; }
    if (d2 < d1) { double t = d1; d1 = d2; d2 = t; int ti = i1; i1 = i2; i2 = ti; }
    if (d1 < d0) { double t = d0; d0 = d1; d1 = t; int ti = i0; i0 = i1; i1 = ti; }
    double Vs[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        Vs[r][0] = V[r][i0];
        Vs[r][1] = V[r][i1];
        Vs[r][2] = V[r][i2];
    }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) V[r][c] = Vs[r][c];
    evals[0] = d0;
    evals[1] = d1;
    evals[2] = d2;
}

// inverse of a symmetric 3x3 {m00,m01,m02,m11,m12,m22} -> same packing
__device__ __forceinline__ void sym3_inverse(const double m[6], double inv[6]) {
    double c00 = m[3] * m[5] - m[4] * m[4];
    double c01 = m[2] * m[4] - m[1] * m[5];
    double c02 = m[1] * m[4] - m[2] * m[3];
    double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    double id = 1.0 / det;
    inv[0] = c00 * id;
    inv[1] = c01 * id;
    inv[2] = c02 * id;
    inv[3] = (m[0] * m[5] - m[2] * m[2]) * id;
    inv[4] = (m[1] * m[2] - m[0] * m[4]) * id;
    inv[5] = (m[0] * m[3] - m[1] * m[1]) * id;
}

// GICP covariance of a point from its normal: Rx diag(eps,1,1) Rx^T with Rx = GetRotationFromE1ToX(n)
// (reference .cpp:4-14 incl. the c < -0.99 -> Identity branch, .cpp:45-51); packed 00,01,02,11,12,22
__device__ inline void gicp_cov_from_normal(double nvx, double nvy, double nvz, double eps, double out6[6]) {
    double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double c = nvx;  // e1 . n
    if (!(c < -0.99)) {
        double v0 = 0.0, v1 = -nvz, v2 = nvy;  // e1 x n
        double S[3][3] = {{0, -v2, v1}, {v2, 0, -v0}, {-v1, v0, 0}};
        double f = 1.0 / (1.0 + c);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int cc = 0; cc < 3; cc++) {
                double s2 = S[r][0] * S[0][cc] + S[r][1] * S[1][cc] + S[r][2] * S[2][cc];
                R[r][cc] += S[r][cc] + s2 * f;
            }
    }
    double dg[3] = {eps, 1.0, 1.0};
    int e = 0;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int cc = r; cc < 3; cc++)
            out6[e++] = R[r][0] * dg[0] * R[cc][0] + R[r][1] * dg[1] * R[cc][1] + R[r][2] * dg[2] * R[cc][2];
}

// order-preserving map of a non-negative float to uint32 (plain bit pattern)
__device__ __forceinline__ uint32_t float_key(float f) { return __float_as_uint(f); }

}  // namespace se3
