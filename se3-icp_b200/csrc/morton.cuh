// morton.cuh — 63-bit Morton code of a point inside the cloud's bounding cube.
#pragma once

#include <stdint.h>

namespace se3 {

__device__ __forceinline__ uint64_t spread21(uint64_t v) {  // 21 bits -> every third bit
    v &= 0x1fffffULL;
    v = (v | (v << 32)) & 0x1f00000000ffffULL;
    v = (v | (v << 16)) & 0x1f0000ff0000ffULL;
    v = (v | (v << 8)) & 0x100f00f00f00f00fULL;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ULL;
    v = (v | (v << 2)) & 0x1249249249249249ULL;
    return v;
}

__device__ __forceinline__ uint64_t morton63(double px, double py, double pz, const double* __restrict__ bbox) {
    double ext = fmax(fmax(bbox[3] - bbox[0], bbox[4] - bbox[1]), fmax(bbox[5] - bbox[2], 1e-300));
    double inv = 2097151.0 / ext;  // same cell size on all axes keeps the cells cubic
    double fx = fmin(fmax((px - bbox[0]) * inv, 0.0), 2097151.0);
    double fy = fmin(fmax((py - bbox[1]) * inv, 0.0), 2097151.0);
    double fz = fmin(fmax((pz - bbox[2]) * inv, 0.0), 2097151.0);
    return spread21((uint64_t)fx) | (spread21((uint64_t)fy) << 1) | (spread21((uint64_t)fz) << 2);
}

}  // namespace se3
