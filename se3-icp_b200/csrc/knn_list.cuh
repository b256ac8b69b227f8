// knn_list.cuh — exact (distance, original index) ordering of up to 128 candidates per warp, shared by the kNN / feature
// pass (knn_features.cu) and the SHOT frame's median vote (shot_lrf.cu).
#pragma once

#include "common.cuh"

namespace se3 {

// The k-nearest list lives in registers: 128 (key, id) slots striped over the warp, element e in lane
// e % 32, slot e / 32, ascending by (distance bits, original index).  Squared distances are >= 0, so
// their IEEE bit patterns order like the values and integer compares replace FP64 compares.
typedef unsigned long long key_t;
constexpr key_t kInfKey = 0x7ff0000000000000ULL;

// distances are non-negative and never NaN, so the FP64 compare (idle FP64 pipe) orders like the bit pattern
__device__ __forceinline__ bool key_less(key_t da, int ia, key_t db, int ib) {
    double xa = __longlong_as_double((long long)da), xb = __longlong_as_double((long long)db);
    return xa < xb || (xa == xb && ia < ib);
}

__device__ __forceinline__ void ce_lane(key_t& dl, int& il, key_t& dh, int& ih) {  // in-lane: low slot gets the min
    if (key_less(dh, ih, dl, il)) {
        key_t t = dl; dl = dh; dh = t;
        int u = il; il = ih; ih = u;
    }
}

__device__ __forceinline__ void ce_shfl(key_t& d, int& i, int j, bool keep_min) {  // with lane ^ j
    key_t od = __shfl_xor_sync(SE3_FULL, d, j);
    int oi = __shfl_xor_sync(SE3_FULL, i, j);
    bool other_less = key_less(od, oi, d, i);
    if (other_less == keep_min) {
        d = od;
        i = oi;
    }
}

// sorts one (key, id) per lane ascending by lane
__device__ __forceinline__ void warp_sort32(key_t& d, int& i, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            bool up = (lane & k) == 0;
            bool lower = (lane & j) == 0;
            ce_shfl(d, i, j, lower == up);
        }
    }
}

// merges 32 candidates (one per lane, any order; unused lanes hold kInfKey) into the sorted 128-list
__device__ __forceinline__ void merge32(key_t (&Ld)[4], int (&Li)[4], key_t cd, int ci, int lane) {
    warp_sort32(cd, ci, lane);
    // element-wise min of the list tail (ascending) with the reversed batch (descending) -> bitonic 128
    key_t bd = __shfl_sync(SE3_FULL, cd, 31 - lane);
    int bi = __shfl_sync(SE3_FULL, ci, 31 - lane);
    if (key_less(bd, bi, Ld[3], Li[3])) {
        Ld[3] = bd;
        Li[3] = bi;
    }
    ce_lane(Ld[0], Li[0], Ld[2], Li[2]);  // distance 64
    ce_lane(Ld[1], Li[1], Ld[3], Li[3]);
    ce_lane(Ld[0], Li[0], Ld[1], Li[1]);  // distance 32
    ce_lane(Ld[2], Li[2], Ld[3], Li[3]);
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        bool lower = (lane & j) == 0;
#pragma unroll
        for (int t = 0; t < 4; t++) ce_shfl(Ld[t], Li[t], j, lower);
    }
}

// element e of the striped list, broadcast to every lane
__device__ __forceinline__ void list_at(const key_t (&Ld)[4], const int (&Li)[4], int e, key_t& d, int& i) {
    int slot = e >> 5;
    key_t sd = slot == 0 ? Ld[0] : slot == 1 ? Ld[1] : slot == 2 ? Ld[2] : Ld[3];
    int si = slot == 0 ? Li[0] : slot == 1 ? Li[1] : slot == 2 ? Li[2] : Li[3];
    d = __shfl_sync(SE3_FULL, sd, e & 31);
    i = __shfl_sync(SE3_FULL, si, e & 31);
}

// Exact (distance, original index) order of the candidate pool by the 64+32-bit merge network: the first `K` entries
// are written back to the pool in ascending order and the K-th distance is returned.  Rare path on both of its uses —
// (1) shrink_pool(): the bisection on the distance VALUE cannot separate candidates that tie at the threshold
// (hundreds of coincident points, e.g. invalid-depth pixels mapped to one xyz), so the pool is cut to exactly the K
// smallest and stays bounded; (2) final ordering: two candidates share a quantised key (sort_pool_quantised).
// Not inlined: keeps the network out of the hot kernel's register budget.
static __device__ __noinline__ double knn_exact_trim(unsigned long long* pd, int* pi, int pool, int K, int lane) {
    key_t Ld[4] = {kInfKey, kInfKey, kInfKey, kInfKey};
    int Li[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int base = 0; base < pool; base += 32) {
        int t = base + lane;
        key_t cd = kInfKey;
        int ci = 0x7fffffff;
        if (t < pool) {
            cd = pd[t];
            ci = pi[t];
        }
        merge32(Ld, Li, cd, ci, lane);
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 4; t++) {
        int j = lane + 32 * t;
        if (j < K) {
            pd[j] = Ld[t];
            pi[j] = Li[t];
        }
    }
    key_t kd;
    int ki;
    list_at(Ld, Li, K - 1, kd, ki);
    __syncwarp();
    return __longlong_as_double((long long)kd);
}

}  // namespace se3
