// se3_key.cuh — 6-D Morton key of an SE(3) element: (unit quaternion xyz with w >= 0, position).
// The rotation leads by two bit levels because it carries most of the variance of the alpha-weighted
// 12-vector (9 alpha^2 vs. ~3 for the position after normalisation).
#pragma once

#include <stdint.h>

namespace se3 {

#ifndef SE3_KEY_LEAD
#define SE3_KEY_LEAD 2
#endif
constexpr int kSe3KeyLead = SE3_KEY_LEAD;

// R in 12-vector order: R[3*c + r] = entry (r, c)
__device__ __forceinline__ void quat_from_columns(const double* R, double q[4]) {
    double r00 = R[0], r10 = R[1], r20 = R[2], r01 = R[3], r11 = R[4], r21 = R[5], r02 = R[6], r12 = R[7], r22 = R[8];
    double tr = r00 + r11 + r22;
    double w, x, y, z;
    if (tr > 0.0) {
        double s = sqrt(tr + 1.0) * 2.0;
        w = 0.25 * s;
        x = (r21 - r12) / s;
        y = (r02 - r20) / s;
        z = (r10 - r01) / s;
    } else if (r00 > r11 && r00 > r22) {
        double s = sqrt(fmax(1.0 + r00 - r11 - r22, 1e-300)) * 2.0;
        w = (r21 - r12) / s;
        x = 0.25 * s;
        y = (r01 + r10) / s;
        z = (r02 + r20) / s;
    } else if (r11 > r22) {
        double s = sqrt(fmax(1.0 + r11 - r00 - r22, 1e-300)) * 2.0;
        w = (r02 - r20) / s;
        x = (r01 + r10) / s;
        y = 0.25 * s;
        z = (r12 + r21) / s;
    } else {
        double s = sqrt(fmax(1.0 + r22 - r00 - r11, 1e-300)) * 2.0;
        w = (r10 - r01) / s;
        x = (r02 + r20) / s;
        y = (r12 + r21) / s;
        z = 0.25 * s;
    }
    if (w < 0.0) {
        w = -w;
        x = -x;
        y = -y;
        z = -z;
    }
    q[0] = w, q[1] = x, q[2] = y, q[3] = z;
}

__device__ __forceinline__ uint32_t quant10(double v, double lo, double inv) {
    double f = (v - lo) * inv;
    f = fmin(fmax(f, 0.0), 1023.0);
    return isfinite(f) ? (uint32_t)f : 0u;
}

// bbox = lo[3], hi[3] of the cloud the key is quantised against (the target's working-frame box)
__device__ __forceinline__ uint64_t se3_key(const double* R, double px, double py, double pz, const double* __restrict__ bbox) {
    double q[4];
    quat_from_columns(R, q);
    uint32_t c[6];
    c[0] = quant10(q[1], -1.0, 511.5);
    c[1] = quant10(q[2], -1.0, 511.5);
    c[2] = quant10(q[3], -1.0, 511.5);
    double ext = fmax(fmax(bbox[3] - bbox[0], bbox[4] - bbox[1]), fmax(bbox[5] - bbox[2], 1e-300));
    double inv = 1023.0 / ext;
    c[3] = quant10(px, bbox[0], inv);
    c[4] = quant10(py, bbox[1], inv);
    c[5] = quant10(pz, bbox[2], inv);
    uint64_t key = 0;
    // kSe3KeyLead leading rotation-only levels, then rotation bit b interleaved with position bit b + kSe3KeyLead
    // (all 60 bits are used: the position bits left over at the end follow on their own)
#pragma unroll
    for (int b = 9; b > 9 - kSe3KeyLead; b--)
#pragma unroll
        for (int d = 0; d < 3; d++) key = (key << 1) | ((c[d] >> b) & 1u);
#pragma unroll
    for (int b = 9 - kSe3KeyLead; b >= 0; b--) {
#pragma unroll
        for (int d = 0; d < 3; d++) key = (key << 1) | ((c[d] >> b) & 1u);
#pragma unroll
        for (int d = 3; d < 6; d++) key = (key << 1) | ((c[d] >> (b + kSe3KeyLead)) & 1u);
    }
#pragma unroll
    for (int b = kSe3KeyLead - 1; b >= 0; b--)
#pragma unroll
        for (int d = 3; d < 6; d++) key = (key << 1) | ((c[d] >> b) & 1u);
    return key;
}

}  // namespace se3
