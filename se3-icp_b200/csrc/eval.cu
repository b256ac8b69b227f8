// eval.cu — the callers either side of the registration path, on the device:
//  * se3icp_eval_error_filterreg   reference src/cc.cpp:4-20  (mean ||T_gt p - T_est p|| over the source)
//  * se3icp_eval_corrs_with_gt     reference src/cc.cpp:116-143 (nearest target point of every T_gt-moved source point)
//  * se3icp_eval_lrf_quality       reference src/cc.cpp:63-88 with angularErrorSO3_alt, cc.cpp:39-61
//  * se3icp_random_downsample      Open3D PointCloud::RandomDownSample as the drivers call it
//                                  (examples/benchmark_synthetic.cpp:100,150): floor(ratio * n) points, uniformly, without
//                                  replacement, in shuffled order.  Open3D shuffles with its own std::mt19937 stream; the
//                                  stream here is a counter-based hash of (seed, index), so the SUBSET differs from Open3D's
//                                  for the same seed while its distribution does not.
// Host buffers in, host buffers out, like every entry point of se3icp.h; the work runs on the context's stream.
#include <cub/device/device_radix_sort.cuh>

#include <vector>

#include "common.cuh"
#include "context.h"
#include "internal.h"

namespace se3 {

__global__ void __launch_bounds__(256) filterreg_kernel(const double* __restrict__ aos, int n, const double* __restrict__ T2,
                                                         double* __restrict__ partial) {
    // T2: T_gt (16 doubles) followed by T_est (16 doubles), row-major
    __shared__ double sm[8];
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double x = aos[3 * (size_t)i], y = aos[3 * (size_t)i + 1], z = aos[3 * (size_t)i + 2];
        double d2 = 0.0;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double a = T2[4 * r] * x + T2[4 * r + 1] * y + T2[4 * r + 2] * z + T2[4 * r + 3];
            const double b = T2[16 + 4 * r] * x + T2[16 + 4 * r + 1] * y + T2[16 + 4 * r + 2] * z + T2[16 + 4 * r + 3];
            d2 += (a - b) * (a - b);
        }
        acc += sqrt(d2);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < 8; k++) s += sm[k];
        partial[blockIdx.x] = s;  // fixed grid, summed in order on the host: deterministic
    }
}

__global__ void __launch_bounds__(256) transform_aos_kernel(const double* __restrict__ in, int n, const double* __restrict__ T,
                                                             double* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double x = in[3 * (size_t)i], y = in[3 * (size_t)i + 1], z = in[3 * (size_t)i + 2];
#pragma unroll
        for (int r = 0; r < 3; r++) out[3 * (size_t)i + r] = T[4 * r] * x + T[4 * r + 1] * y + T[4 * r + 2] * z + T[4 * r + 3];
    }
}

// frames: [count][16] row-major 4x4; pairs: [k][2]; err[k] in degrees (angularErrorSO3_alt incl. its clamped acos)
__global__ void __launch_bounds__(256) lrf_quality_kernel(const double* __restrict__ src_frames, const double* __restrict__ tgt_frames,
                                                           const double* __restrict__ Tgt, const int* __restrict__ pairs, int k,
                                                           double* __restrict__ err) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < k; p += gridDim.x * blockDim.x) {
        const double* A = src_frames + 16 * (size_t)pairs[2 * p];
        const double* B = tgt_frames + 16 * (size_t)pairs[2 * p + 1];
        // R1 = R_gt * R_src (rotation block of map_gt * source_SE3), trace(R1^T R2) = sum_ij R1_ij R2_ij
        double tr = 0.0;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double r1 = Tgt[4 * r] * A[c] + Tgt[4 * r + 1] * A[4 + c] + Tgt[4 * r + 2] * A[8 + c];
                tr += r1 * B[4 * r + c];
            }
        const double a = (tr - 1.0) / 2.0;
        const double ang = a <= -1.0 ? 3.14159265358979323846 : (a >= 1.0 ? 0.0 : acos(a));
        err[p] = fabs(ang) * (180.0 / 3.14159265358979323846);
    }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) shuffle_keys_kernel(int n, uint64_t seed, uint64_t* __restrict__ keys, int* __restrict__ vals) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        keys[i] = mix64(mix64(seed) + 0x9e3779b97f4a7c15ULL * (uint64_t)(i + 1));
        vals[i] = i;
    }
}

__global__ void __launch_bounds__(256) gather_aos_kernel(const double* __restrict__ in, const int* __restrict__ idx, int k,
                                                          double* __restrict__ out) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
        const size_t i = (size_t)idx[j];
        out[3 * (size_t)j] = in[3 * i], out[3 * (size_t)j + 1] = in[3 * i + 1], out[3 * (size_t)j + 2] = in[3 * i + 2];
    }
}

static int grid_of(size_t n) {
    size_t g = (n + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace se3

using namespace se3;

namespace {
int eval_ctx(se3icp_ctx* c) {
    if (!c) {
        set_last_error("null context");
        return SE3ICP_ERR_ARG;
    }
    SE3_CUDA(cudaSetDevice(c->device));
    if (c->run_pending) {
        set_last_error("evaluation call while a run is pending (call se3icp_run_finish)");
        return SE3ICP_ERR_STATE;
    }
    return 0;
}
}  // namespace

extern "C" {

int se3icp_eval_error_filterreg(se3icp_ctx* c, const double* src_xyz, size_t n, const double* T_gt, const double* T_est,
                                double* error_out) {
    SE3_TRY(eval_ctx(c));
    if (!src_xyz || !T_gt || !T_est || !error_out || n == 0 || n > 0x7fffffffULL / 4) return SE3ICP_ERR_ARG;
    cudaStream_t st = c->stream;
    const int blocks = kReduceBlocks;
    SE3_TRY(c->scratch.ensure(n * 3 * sizeof(double) + (32 + blocks) * sizeof(double)));
    double* d_xyz = c->scratch.as<double>();
    double* d_T = d_xyz + 3 * n;
    double* d_part = d_T + 32;
    double hT[32];
    memcpy(hT, T_gt, 16 * sizeof(double));
    memcpy(hT + 16, T_est, 16 * sizeof(double));
    SE3_CUDA(cudaMemcpyAsync(d_xyz, src_xyz, n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    SE3_CUDA(cudaMemcpyAsync(d_T, hT, sizeof(hT), cudaMemcpyHostToDevice, st));
    filterreg_kernel<<<blocks, 256, 0, st>>>(d_xyz, (int)n, d_T, d_part);
    SE3_CUDA(cudaGetLastError());
    std::vector<double> part(blocks);
    SE3_CUDA(cudaMemcpyAsync(part.data(), d_part, blocks * sizeof(double), cudaMemcpyDeviceToHost, st));
    SE3_CUDA(cudaStreamSynchronize(st));
    double s = 0.0;
    for (double v : part) s += v;
    *error_out = s / (double)n;
    return SE3ICP_OK;
}

int se3icp_eval_corrs_with_gt(se3icp_ctx* c, const double* src_xyz, size_t n, const double* tgt_xyz, size_t m, const double* T_gt,
                              int32_t* tgt_idx) {
    SE3_TRY(eval_ctx(c));
    if (!src_xyz || !tgt_xyz || !T_gt || !tgt_idx || n == 0 || m == 0 || n > 0x7fffffffULL / 4) return SE3ICP_ERR_ARG;
    // move the source on the device, then the exact 3-D nearest-neighbour pass of the registration path
    cudaStream_t st = c->stream;
    SE3_TRY(c->eval_buf.ensure(2 * n * 3 * sizeof(double) + 16 * sizeof(double)));
    double* d_in = c->eval_buf.as<double>();
    double* d_out = d_in + 3 * n;
    double* d_T = d_out + 3 * n;
    SE3_CUDA(cudaMemcpyAsync(d_in, src_xyz, n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    SE3_CUDA(cudaMemcpyAsync(d_T, T_gt, 16 * sizeof(double), cudaMemcpyHostToDevice, st));
    transform_aos_kernel<<<grid_of(n), 256, 0, st>>>(d_in, (int)n, d_T, d_out);
    SE3_CUDA(cudaGetLastError());
    std::vector<double> moved(3 * n);
    SE3_CUDA(cudaMemcpyAsync(moved.data(), d_out, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    SE3_CUDA(cudaStreamSynchronize(st));
    return se3icp_nn_xyz(c, moved.data(), n, tgt_xyz, m, tgt_idx, nullptr);
}

int se3icp_eval_lrf_quality(se3icp_ctx* c, const double* src_frames16, size_t n, const double* tgt_frames16, size_t m,
                            const double* T_gt, const int32_t* pairs, size_t n_pairs, double* mean_error_deg,
                            double* per_pair_error_deg) {
    SE3_TRY(eval_ctx(c));
    if (!src_frames16 || !tgt_frames16 || !T_gt || !pairs || !mean_error_deg || n == 0 || m == 0 || n_pairs == 0 ||
        n_pairs > 0x7fffffffULL)
        return SE3ICP_ERR_ARG;
    for (size_t p = 0; p < n_pairs; p++)
        if (pairs[2 * p] < 0 || (size_t)pairs[2 * p] >= n || pairs[2 * p + 1] < 0 || (size_t)pairs[2 * p + 1] >= m) {
            set_last_error("se3icp_eval_lrf_quality: pair %zu out of range", p);
            return SE3ICP_ERR_ARG;
        }
    cudaStream_t st = c->stream;
    const size_t bytes = (16 * (n + m) + 16 + n_pairs) * sizeof(double) + 2 * n_pairs * sizeof(int);
    SE3_TRY(c->eval_buf.ensure(bytes));
    double* d_src = c->eval_buf.as<double>();
    double* d_tgt = d_src + 16 * n;
    double* d_T = d_tgt + 16 * m;
    double* d_err = d_T + 16;
    int* d_pairs = reinterpret_cast<int*>(d_err + n_pairs);
    SE3_CUDA(cudaMemcpyAsync(d_src, src_frames16, 16 * n * sizeof(double), cudaMemcpyHostToDevice, st));
    SE3_CUDA(cudaMemcpyAsync(d_tgt, tgt_frames16, 16 * m * sizeof(double), cudaMemcpyHostToDevice, st));
    SE3_CUDA(cudaMemcpyAsync(d_T, T_gt, 16 * sizeof(double), cudaMemcpyHostToDevice, st));
    SE3_CUDA(cudaMemcpyAsync(d_pairs, pairs, 2 * n_pairs * sizeof(int), cudaMemcpyHostToDevice, st));
    lrf_quality_kernel<<<grid_of(n_pairs), 256, 0, st>>>(d_src, d_tgt, d_T, d_pairs, (int)n_pairs, d_err);
    SE3_CUDA(cudaGetLastError());
    std::vector<double> err(n_pairs);
    SE3_CUDA(cudaMemcpyAsync(err.data(), d_err, n_pairs * sizeof(double), cudaMemcpyDeviceToHost, st));
    SE3_CUDA(cudaStreamSynchronize(st));
    double s = 0.0;
    for (double v : err) s += v;  // pair order, as the reference accumulates
    *mean_error_deg = s / (double)n_pairs;
    if (per_pair_error_deg) memcpy(per_pair_error_deg, err.data(), n_pairs * sizeof(double));
    return SE3ICP_OK;
}

int se3icp_random_downsample(se3icp_ctx* c, const double* xyz, size_t n, double sampling_ratio, uint64_t seed, double* xyz_out,
                             int32_t* index_out, size_t* n_out) {
    SE3_TRY(eval_ctx(c));
    if (!xyz || !n_out || n == 0 || n > 0x7fffffffULL / 4 || !(sampling_ratio >= 0.0) || sampling_ratio > 1.0) return SE3ICP_ERR_ARG;
    size_t k = (size_t)((double)n * sampling_ratio);  // Open3D: (size_t)(n * ratio)
    if (k > n) k = n;
    *n_out = k;
    if (k == 0 || (!xyz_out && !index_out)) return SE3ICP_OK;
    cudaStream_t st = c->stream;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, (int)n, 0, 64);
    const size_t need = 2 * n * sizeof(uint64_t) + 2 * n * sizeof(int) + (n + k) * 3 * sizeof(double) + tmp_bytes + 64;
    SE3_TRY(c->eval_buf.ensure(need));
    uint64_t* d_keys = c->eval_buf.as<uint64_t>();
    uint64_t* d_keys2 = d_keys + n;
    double* d_xyz = reinterpret_cast<double*>(d_keys2 + n);
    double* d_sel = d_xyz + 3 * n;
    int* d_vals = reinterpret_cast<int*>(d_sel + 3 * k);
    int* d_vals2 = d_vals + n;
    void* d_tmp = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(d_vals2 + n) + 15) & ~(uintptr_t)15);
    shuffle_keys_kernel<<<grid_of(n), 256, 0, st>>>((int)n, seed, d_keys, d_vals);
    SE3_CUDA(cudaGetLastError());
    SE3_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 64, st));
    if (xyz_out) {
        SE3_CUDA(cudaMemcpyAsync(d_xyz, xyz, n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
        gather_aos_kernel<<<grid_of(k), 256, 0, st>>>(d_xyz, d_vals2, (int)k, d_sel);
        SE3_CUDA(cudaGetLastError());
        SE3_CUDA(cudaMemcpyAsync(xyz_out, d_sel, k * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (index_out) SE3_CUDA(cudaMemcpyAsync(index_out, d_vals2, k * sizeof(int), cudaMemcpyDeviceToHost, st));
    SE3_CUDA(cudaStreamSynchronize(st));
    return SE3ICP_OK;
}

}  // extern "C"
