// se3_index.cu — the 12-D search structure over the target's SE(3) rows.
//
// Replaces the reference's 12 x M matrix + nanoflann kd-tree (reference .cpp:610-626).  The rows
// live on a 6-dimensional manifold (3 rotation + 3 translation degrees of freedom), so the target is
// ordered by a 6-D Morton code of (unit-quaternion xyz, position); 32 consecutive rows form a leaf
// and every node of the implicit 32-wide hierarchy stores an outward-rounded 12-D box.  Measured on
// the KITTI-like workload this visits ~10 of 3 730 leaves per query (a 3-D order visits ~300).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "index_storage.h"
#include "internal.h"
#include "se3_key.cuh"

namespace se3 {

// all 60 key bits are sorted: dropping the low 12 (six radix passes instead of eight, ~10 us each at 120 k rows) made the
// SE(3)-phase search of a KITTI-size pair 2.5 % (50 us) slower
constexpr int kSe3SortLow = 0;

__global__ void __launch_bounds__(256) se3_key_kernel(const double* __restrict__ frame, const double* __restrict__ x,
                                                       const double* __restrict__ y, const double* __restrict__ z, int n,
                                                       const double* __restrict__ bbox, uint64_t* __restrict__ keys,
                                                       int* __restrict__ vals) {
    const size_t nn = (size_t)n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double R[9];
#pragma unroll
        for (int k = 0; k < 9; k++) R[k] = frame[k * nn + i];
        keys[i] = se3_key(R, x[i], y[i], z[i], bbox) & ~((1ULL << kSe3SortLow) - 1ULL);  // only the sorted bits are kept
        vals[i] = i;
    }
}

// rows in 6-D Morton order: [alpha R (column-major 9) | tscale * p]   (reference .cpp:597-625)
__global__ void __launch_bounds__(256) pack_se3_rows_kernel(const double* __restrict__ frame, const double* __restrict__ x,
                                                             const double* __restrict__ y, const double* __restrict__ z,
                                                             const int* __restrict__ perm12, int n, double alpha,
                                                             double tscale, float4* __restrict__ rows32,
                                                             double* __restrict__ rows64, int* __restrict__ inv12,
                                                             IterState* __restrict__ state) {
    const size_t nn = (size_t)n;
    double amax = 0.0;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        int o = perm12[s];
        inv12[o] = s;
        double v[12];
#pragma unroll
        for (int k = 0; k < 9; k++) v[k] = frame[k * nn + o] * alpha;
        v[9] = x[o] * tscale;
        v[10] = y[o] * tscale;
        v[11] = z[o] * tscale;
#pragma unroll
        for (int k = 0; k < 12; k++) {
            rows64[k * nn + s] = v[k];
            amax = fmax(amax, fabs(v[k]));
        }
        rows32[s] = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
        rows32[nn + s] = make_float4((float)v[4], (float)v[5], (float)v[6], (float)v[7]);
        rows32[2 * nn + s] = make_float4((float)v[8], (float)v[9], (float)v[10], (float)v[11]);
    }
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0 && amax > 0.0)
        atomicMax(reinterpret_cast<unsigned long long*>(&state->tgt_absmax), (unsigned long long)__double_as_longlong(amax));
}

__global__ void __launch_bounds__(256) box12_leaf_kernel(const double* __restrict__ rows64, int n, int n_leaves,
                                                          int total_nodes, float2* __restrict__ box12) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_leaves) return;
    int p = warp * 32 + lane;
    const size_t nn = (size_t)n, tn = (size_t)total_nodes;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        double lo = 1e300, hi = -1e300;
        if (p < n) lo = hi = rows64[k * nn + p];
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (lane == 0) box12[box12_slot(k, warp, tn)] = make_float2(__double2float_rd(lo), __double2float_ru(hi));
    }
}

__global__ void __launch_bounds__(256) box12_upper_kernel(int child_off, int child_cnt, int node_off, int node_cnt,
                                                           int total_nodes, float2* __restrict__ box12) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= node_cnt) return;
    int c = warp * 32 + lane;
    const size_t tn = (size_t)total_nodes;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        float lo = 3.0e38f, hi = -3.0e38f;
        if (c < child_cnt) {
            const float2 lh = box12[box12_slot(k, child_off + c, tn)];
            lo = lh.x;
            hi = lh.y;
        }
        lo = warp_minf(lo);
        hi = warp_maxf(hi);
        if (lane == 0) box12[box12_slot(k, node_off + warp, tn)] = make_float2(lo, hi);
    }
}

int Se3IndexStorage::reserve(int n, const CloudIndex& levels) {
    size_t nn = (size_t)(n > 0 ? n : 1);
    SE3_TRY(perm12.ensure(nn * sizeof(int)));
    SE3_TRY(keys12.ensure(nn * sizeof(uint64_t)));
    SE3_TRY(keys_tmp.ensure(nn * sizeof(uint64_t)));
    SE3_TRY(vals_tmp.ensure(nn * sizeof(int)));
    SE3_TRY(box12.ensure((size_t)(levels.total_nodes > 0 ? levels.total_nodes : 1) * 24 * sizeof(float)));
    SE3_TRY(rows32.ensure(3 * nn * sizeof(float4)));
    SE3_TRY(rows64.ensure(12 * nn * sizeof(double)));
    SE3_TRY(inv12.ensure(nn * sizeof(int)));
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, n, 0, 64);
    SE3_TRY(sort_tmp.ensure(tmp_bytes + 16));
    sort_tmp_bytes = tmp_bytes;
    return 0;
}

// frame: [9][n] rotation planes (original order); xyz from the cloud's 3-D index (original order)
int Se3IndexStorage::build(const CloudIndex& I, const double* frame, double alpha, double tscale, IterState* state,
                           cudaStream_t st, long long* launches) {
    int n = I.n;
    int g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    se3_key_kernel<<<g, 256, 0, st>>>(frame, I.x, I.y, I.z, n, I.bbox, keys_tmp.as<uint64_t>(), vals_tmp.as<int>());
    SE3_CUDA(cudaGetLastError());
    size_t tb = sort_tmp_bytes;
    SE3_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp.ptr, tb, keys_tmp.as<uint64_t>(), keys12.as<uint64_t>(),
                                             vals_tmp.as<int>(), perm12.as<int>(), n, kSe3SortLow, 60, st));
    pack_se3_rows_kernel<<<g, 256, 0, st>>>(frame, I.x, I.y, I.z, perm12.as<int>(), n, alpha, tscale, rows32.as<float4>(),
                                             rows64.as<double>(), inv12.as<int>(), state);
    int n_leaves = I.level_cnt[0];
    box12_leaf_kernel<<<(n_leaves * 32 + 255) / 256, 256, 0, st>>>(rows64.as<double>(), n, n_leaves, I.total_nodes,
                                                                   box12.as<float2>());
    for (int l = 1; l < I.n_levels; l++)
        box12_upper_kernel<<<(I.level_cnt[l] * 32 + 255) / 256, 256, 0, st>>>(
            I.level_off[l - 1], I.level_cnt[l - 1], I.level_off[l], I.level_cnt[l], I.total_nodes, box12.as<float2>());
    SE3_CUDA(cudaGetLastError());
    if (launches) *launches += 3 + I.n_levels + 3;
    return 0;
}

}  // namespace se3
