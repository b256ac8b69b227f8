// distance_store.cuh — the one place a correspondence distance is written (reference .cpp:411-413, 465-467): FP64 value,
// the float PCL stores, and — when the single-pass trimmed rejection is armed (CorrBuffers::thist) — one count in the
// 16-bit histogram of the float's sort key, so that the rejection needs no histogram passes of its own.
#pragma once

#include "internal.h"

namespace se3 {

// order-preserving key of a non-negative float distance; complemented when the LARGEST distances are kept, so that
// "keep the n smallest keys" serves both comparator directions
__device__ __forceinline__ unsigned int trim_key(float d, int keep_largest) {
    unsigned int b = __float_as_uint(d);
    return keep_largest ? ~b : b;
}

__device__ __forceinline__ void store_distance(const RunConfig& cfg, const CorrBuffers& cb, int i, double d) {
    const float f = (float)d;
    cb.dist[i] = d;
    cb.distf[i] = f;
    if (cb.thist) atomicAdd(&cb.thist[trim_key(f, cfg.keep_largest) >> 16], 1u);
}

// (key, source index) order of the trimmed rejection: correspondence i survives iff its pair is <= the selected one
__device__ __forceinline__ bool trim_keeps(unsigned int key, int i, unsigned int thr_bits, int tie_limit) {
    return key < thr_bits || (key == thr_bits && i <= tie_limit);
}

}  // namespace se3
