// spatial_index.cu — normalisation and the on-device spatial index (a3).
//
// Replaces reference .cpp:568-582 (GetCenter / largestDistanceFromGivenPoint / Translate / Scale)
// and the nanoflann kd-tree builds at .cpp:482,586-587 with a Morton-ordered, implicit 32-wide
// bounding-box hierarchy: leaves are 32 consecutive points of the Morton order, each upper level
// groups 32 nodes of the level below.  One warp tests 32 boxes per step during traversal.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "index_storage.h"
#include "internal.h"
#include "morton.cuh"

namespace se3 {

// ------------------------------------------------------------------------------------------------
// reductions for the normalisation.  Fixed grid of kReduceBlocks blocks; every consumer kernel
// re-derives the final value from the per-block partials in a fixed order (deterministic, no
// extra "finalise" launches).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sum_xyz_kernel(const double* __restrict__ aos, int n, double* __restrict__ partial) {
    double sx = 0, sy = 0, sz = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        sx += aos[3 * (size_t)i];
        sy += aos[3 * (size_t)i + 1];
        sz += aos[3 * (size_t)i + 2];
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    sz = warp_sum(sz);
    __shared__ double sm[3][8];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        sm[0][w] = sx;
        sm[1][w] = sy;
        sm[2][w] = sz;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0;
        for (int k = 0; k < 8; k++) s += sm[threadIdx.x][k];
        partial[blockIdx.x * 3 + threadIdx.x] = s;
    }
}

// Block-cooperative (256 threads), fixed-order reduction of the per-block partial sums -> centre.
// Must be called by all threads of the block; c lives in shared memory.
__device__ __forceinline__ void center_from_partials(const double* __restrict__ partial, int n, double* c, double (*scratch)[8]) {
    double s0 = 0, s1 = 0, s2 = 0;
    for (int b = threadIdx.x; b < kReduceBlocks; b += blockDim.x) {
        s0 += partial[3 * b];
        s1 += partial[3 * b + 1];
        s2 += partial[3 * b + 2];
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        scratch[0][w] = s0;
        scratch[1][w] = s1;
        scratch[2][w] = s2;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0;
        for (int k = 0; k < 8; k++) t += scratch[threadIdx.x][k];
        c[threadIdx.x] = t * (1.0 / (double)n);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) maxdist_kernel(const double* __restrict__ aos, int n, const double* __restrict__ psum,
                                                       double* __restrict__ pmax) {
    __shared__ double c[3];
    __shared__ double sm[8];
    __shared__ double scratch[3][8];
    center_from_partials(psum, n, c, scratch);
    double best = -1.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double dx = aos[3 * (size_t)i] - c[0], dy = aos[3 * (size_t)i + 1] - c[1], dz = aos[3 * (size_t)i + 2] - c[2];
        best = fmax(best, sqrt(dx * dx + dy * dy + dz * dz));
    }
    best = warp_max(best);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = sm[0];
        for (int k = 1; k < 8; k++) b = fmax(b, sm[k]);
        pmax[blockIdx.x] = b;
    }
}

// p' = (p - c) * s with s = scale_pre / max(r_src, r_tgt)   (reference .cpp:574-582)
__global__ void __launch_bounds__(256) normalise_kernel(const double* __restrict__ aos, int n, const double* __restrict__ psum_self,
                                                         const double* __restrict__ pmax_src, const double* __restrict__ pmax_tgt,
                                                         double scale_pre, int which, IterState* __restrict__ state,
                                                         double* __restrict__ x, double* __restrict__ y, double* __restrict__ z) {
    __shared__ double c[3];
    __shared__ double s_scale;
    __shared__ double scratch[3][8];
    center_from_partials(psum_self, n, c, scratch);
    double rmax = -1.0;
    for (int b = threadIdx.x; b < kReduceBlocks; b += blockDim.x) rmax = fmax(rmax, fmax(pmax_src[b], pmax_tgt[b]));
    rmax = warp_max(rmax);
    if ((threadIdx.x & 31) == 0) scratch[0][threadIdx.x >> 5] = rmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = scratch[0][0];
        for (int k = 1; k < 8; k++) r = fmax(r, scratch[0][k]);
        s_scale = scale_pre * (1.0 / r);
        if (blockIdx.x == 0) {
            double* dst = which == SE3ICP_SOURCE ? state->c_src : state->c_tgt;
            dst[0] = c[0];
            dst[1] = c[1];
            dst[2] = c[2];
            state->scale = s_scale;
        }
    }
    __syncthreads();
    double s = s_scale;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // Translate(-c) then Scale(s): (p + (-c)) * s
        x[i] = (aos[3 * (size_t)i] + (-c[0])) * s;
        y[i] = (aos[3 * (size_t)i + 1] + (-c[1])) * s;
        z[i] = (aos[3 * (size_t)i + 2] + (-c[2])) * s;
    }
}

__global__ void __launch_bounds__(256) aos_to_soa_kernel(const double* __restrict__ aos, int n, double* __restrict__ x,
                                                          double* __restrict__ y, double* __restrict__ z) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        x[i] = aos[3 * (size_t)i];
        y[i] = aos[3 * (size_t)i + 1];
        z[i] = aos[3 * (size_t)i + 2];
    }
}

// reference .cpp:16-30 lounge_point_confidence (depth only; p1*min_depth deliberately not squared)
__global__ void __launch_bounds__(256) confidence_kernel(const double* __restrict__ aos, int n, double* __restrict__ conf) {
    const double p1 = 0.002203, p2 = -0.001028, p3 = 0.0005351, min_depth = 0.4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double depth = aos[3 * (size_t)i + 2];
        double error = p1 * depth * depth + p2 * depth + p3;
        conf[i] = (p1 * min_depth + p2 * min_depth + p3) / error;
    }
}

static inline int grid_for(int n, int threads, int cap) {
    int g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    return g > cap ? cap : g;
}

int launch_sum_xyz(const double* aos, int n, double* partial, cudaStream_t st) {
    sum_xyz_kernel<<<kReduceBlocks, 256, 0, st>>>(aos, n, partial);
    SE3_CUDA(cudaGetLastError());
    return 0;
}
int launch_maxdist(const double* aos, int n, const double* psum, double* pmax, cudaStream_t st) {
    maxdist_kernel<<<kReduceBlocks, 256, 0, st>>>(aos, n, psum, pmax);
    SE3_CUDA(cudaGetLastError());
    return 0;
}
int launch_normalise(const double* aos, int n, const double* psum_self, const double* pmax_src, const double* pmax_tgt,
                     int, int, double scale_pre, int which, IterState* state, double* x, double* y, double* z,
                     cudaStream_t st) {
    normalise_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, st>>>(aos, n, psum_self, pmax_src, pmax_tgt, scale_pre, which,
                                                                 state, x, y, z);
    SE3_CUDA(cudaGetLastError());
    return 0;
}
int launch_aos_to_soa(const double* aos, int n, double* x, double* y, double* z, cudaStream_t st) {
    aos_to_soa_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, st>>>(aos, n, x, y, z);
    SE3_CUDA(cudaGetLastError());
    return 0;
}
int launch_confidence(const double* aos, int n, double* conf, cudaStream_t st) {
    confidence_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, st>>>(aos, n, conf);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Morton order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bbox_partial_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                            const double* __restrict__ z, int n, double* __restrict__ part) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double v[3] = {x[i], y[i], z[i]};
#pragma unroll
        for (int d = 0; d < 3; d++) {
            lo[d] = fmin(lo[d], v[d]);
            hi[d] = fmax(hi[d], v[d]);
        }
    }
    __shared__ double sm[6][8];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double a = warp_min(lo[d]), b = warp_max(hi[d]);
        if (l == 0) {
            sm[d][w] = a;
            sm[3 + d][w] = b;
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double r = sm[threadIdx.x][0];
        for (int k = 1; k < 8; k++) r = threadIdx.x < 3 ? fmin(r, sm[threadIdx.x][k]) : fmax(r, sm[threadIdx.x][k]);
        part[blockIdx.x * 6 + threadIdx.x] = r;
    }
}

// 6 warps, warp d reduces component d of the per-block boxes
__global__ void bbox_final_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ bbox) {
    int d = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (d >= 6) return;
    double r = d < 3 ? 1e300 : -1e300;
    for (int b = lane; b < nblocks; b += 32) r = d < 3 ? fmin(r, part[b * 6 + d]) : fmax(r, part[b * 6 + d]);
    r = d < 3 ? warp_min(r) : warp_max(r);
    if (lane == 0) bbox[d] = r;
}

__global__ void __launch_bounds__(256) morton_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                      const double* __restrict__ z, int n, const double* __restrict__ bbox,
                                                      uint64_t* __restrict__ keys, int* __restrict__ vals, uint64_t mask) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        keys[i] = morton63(x[i], y[i], z[i], bbox) & mask;  // only the sorted bits are kept: the key array is then fully ordered
        vals[i] = i;
    }
}

// Bits of the 63-bit code the radix sort looks at (one 8-bit pass each ~10 us at 120 k points, more than the rest of an
// index build): 48, i.e. cells of extent / 65536; points sharing a cell keep their input order.
static int morton_sort_bits(int n) { return 48; }

__global__ void __launch_bounds__(256) gather_sorted_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                             const double* __restrict__ z, const int* __restrict__ perm, int n,
                                                             double* __restrict__ sx, double* __restrict__ sy,
                                                             double* __restrict__ sz, int* __restrict__ inv) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        int o = perm[s];
        inv[o] = s;  // Morton position of an original index (the sort's value input is dead by now and holds it)
        sx[s] = x[o];
        sy[s] = y[o];
        sz[s] = z[o];
    }
}

// leaves: one warp per 32 consecutive Morton points; boxes rounded outward to float
__global__ void __launch_bounds__(256) leaf_box_kernel(const double* __restrict__ sx, const double* __restrict__ sy,
                                                        const double* __restrict__ sz, int n, int n_leaves, int total_nodes,
                                                        float2* __restrict__ box) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_leaves) return;
    int p = warp * 32 + lane;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    if (p < n) {
        lo[0] = hi[0] = sx[p];
        lo[1] = hi[1] = sy[p];
        lo[2] = hi[2] = sz[p];
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double a = warp_min(lo[d]), b = warp_max(hi[d]);
        if (lane == 0) box[(size_t)d * total_nodes + warp] = make_float2(__double2float_rd(a), __double2float_ru(b));
    }
}

// upper levels: one warp per node, union of up to 32 child boxes
__global__ void __launch_bounds__(256) upper_box_kernel(int child_off, int child_cnt, int node_off, int node_cnt,
                                                         int total_nodes, float2* __restrict__ box) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= node_cnt) return;
    int c = warp * 32 + lane;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    if (c < child_cnt) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const float2 lh = box[(size_t)d * total_nodes + child_off + c];
            lo[d] = lh.x;
            hi[d] = lh.y;
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        float a = warp_minf(lo[d]), b = warp_maxf(hi[d]);
        if (lane == 0) box[(size_t)d * total_nodes + node_off + warp] = make_float2(a, b);
    }
}

void IndexStorage::plan_levels(int n) {
    view.n = n;
    int cnt = (n + 31) / 32;
    int lvl = 0, off = 0;
    while (true) {
        view.level_off[lvl] = off;
        view.level_cnt[lvl] = cnt;
        off += cnt;
        lvl++;
        if (cnt <= 32 || lvl >= kMaxLevels) break;
        cnt = (cnt + 31) / 32;
    }
    view.n_levels = lvl;
    view.total_nodes = off;
    // Traversals start at the lowest level that is still small enough to be tested whole (a few independent rounds of
    // 32 boxes) instead of walking down from the root one dependent expansion at a time; the bound keeps the initial
    // pushes plus 31 per level below inside the traversal stack (kStackEntries = 192).
    view.start_level = lvl - 1;
    for (int l = 0; l < lvl; l++) {
        int cap = 191 - 31 * l;
        if (view.level_cnt[l] <= (cap < 128 ? cap : 128)) {
            view.start_level = l;
            break;
        }
    }
}

int IndexStorage::reserve(int n) {
    plan_levels(n);
    size_t nn = (size_t)(n > 0 ? n : 1);
    SE3_TRY(x.ensure(nn * sizeof(double)));
    SE3_TRY(y.ensure(nn * sizeof(double)));
    SE3_TRY(z.ensure(nn * sizeof(double)));
    SE3_TRY(sx.ensure(nn * sizeof(double)));
    SE3_TRY(sy.ensure(nn * sizeof(double)));
    SE3_TRY(sz.ensure(nn * sizeof(double)));
    SE3_TRY(perm.ensure(nn * sizeof(int)));
    SE3_TRY(keys.ensure(nn * sizeof(uint64_t)));
    SE3_TRY(keys_tmp.ensure(nn * sizeof(uint64_t)));
    SE3_TRY(vals_tmp.ensure(nn * sizeof(int)));
    SE3_TRY(box.ensure((size_t)(view.total_nodes > 0 ? view.total_nodes : 1) * 6 * sizeof(float)));
    SE3_TRY(bbox.ensure(6 * sizeof(double)));
    SE3_TRY(bbox_part.ensure((size_t)kReduceBlocks * 6 * sizeof(double)));
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, n, 0, 63);
    SE3_TRY(sort_tmp.ensure(tmp_bytes + 16));
    sort_tmp_bytes = tmp_bytes;
    view.inv = vals_tmp.as<int>();
    view.x = x.as<double>();
    view.y = y.as<double>();
    view.z = z.as<double>();
    view.sx = sx.as<double>();
    view.sy = sy.as<double>();
    view.sz = sz.as<double>();
    view.perm = perm.as<int>();
    view.keys = keys.as<uint64_t>();
    view.box = box.as<float2>();
    view.bbox = bbox.as<double>();
    return 0;
}

// x,y,z must already hold the working-frame coordinates.
int IndexStorage::build(cudaStream_t st, long long* launches) {
    int n = view.n;
    if (n <= 0) return SE3ICP_ERR_ARG;
    int g = grid_for(n, 256, 148 * 8);
    bbox_partial_kernel<<<kReduceBlocks, 256, 0, st>>>(x.as<double>(), y.as<double>(), z.as<double>(), n,
                                                        bbox_part.as<double>());
    bbox_final_kernel<<<1, 192, 0, st>>>(bbox_part.as<double>(), kReduceBlocks, bbox.as<double>());
    const int low = 63 - morton_sort_bits(n);
    morton_kernel<<<g, 256, 0, st>>>(x.as<double>(), y.as<double>(), z.as<double>(), n, bbox.as<double>(),
                                      keys_tmp.as<uint64_t>(), vals_tmp.as<int>(), ~((1ULL << low) - 1ULL));
    SE3_CUDA(cudaGetLastError());
    size_t tb = sort_tmp_bytes;
    SE3_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp.ptr, tb, keys_tmp.as<uint64_t>(), keys.as<uint64_t>(),
                                             vals_tmp.as<int>(), perm.as<int>(), n, low, 63, st));
    gather_sorted_kernel<<<g, 256, 0, st>>>(x.as<double>(), y.as<double>(), z.as<double>(), perm.as<int>(), n,
                                             sx.as<double>(), sy.as<double>(), sz.as<double>(), vals_tmp.as<int>());
    int n_leaves = view.level_cnt[0];
    leaf_box_kernel<<<(n_leaves * 32 + 255) / 256, 256, 0, st>>>(sx.as<double>(), sy.as<double>(), sz.as<double>(), n,
                                                                 n_leaves, view.total_nodes, box.as<float2>());
    for (int l = 1; l < view.n_levels; l++) {
        upper_box_kernel<<<(view.level_cnt[l] * 32 + 255) / 256, 256, 0, st>>>(
            view.level_off[l - 1], view.level_cnt[l - 1], view.level_off[l], view.level_cnt[l], view.total_nodes,
            box.as<float2>());
    }
    SE3_CUDA(cudaGetLastError());
    if (launches) *launches += 6 + (view.n_levels - 1) + 3;  // + CUB's internal passes (approx.)
    return 0;
}

}  // namespace se3
