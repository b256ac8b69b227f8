// nn_search.cu — correspondence search.
//
//  * nn_se3_tree        reference .cpp:444-470 (update_correspondences_raw_flann_SE3): warp-per-query pruned
//                       traversal of the 12-D box hierarchy built by se3_index.cu; FP32 directed-rounding
//                       box bounds, exact FP64 leaf distances, warm-started from the previous match.
//  * nn_se3_brute       reference .cpp:444-470 (update_correspondences_raw_flann_SE3): tiled FP32
//                       sweep with top-2 tracking and a rigorous certification test; the query is
//                       T_total * X0 formed on the fly (the reference's rewrite of source_se3_cloud_,
//                       .cpp:713-716, is fused away).
//  * nn_se3_repair      exact FP64 sweep for the queries the FP32 pass could not certify
//                       (also the whole search in SE3ICP_NN_EXACT_F64 mode).
//  * nn_xyz             reference .cpp:402-416 (update_correspondences_kd_tree_XYZ): warp-per-query
//                       pruned traversal, FP64, warm-started from the previous correspondence.
// Ties resolve to the smallest original index (the oracle's rule).
#include "common.cuh"
#include "distance_store.cuh"
#include "internal.h"
#include "morton.cuh"
#include "se3_key.cuh"
#include "traverse.cuh"

namespace se3 {

#if NN_COUNT_VISITS  // development build only (profiles/experiments): how many boxes / leaves the 12-D search opens
__device__ unsigned long long g_nn_visits[4];  // queries, node tests (32 boxes each), leaves evaluated, exact rows
#define NN_COUNT(slot, v) do { const unsigned long long v__ = (unsigned long long)(v); if (lane == 0) atomicAdd(&g_nn_visits[slot], v__); } while (0)
#else
#define NN_COUNT(slot, v) do { } while (0)
#endif

__device__ __forceinline__ bool se3_phase_active(const RunConfig& cfg, const IterState* st) {
    return cfg.has_se3 && (cfg.pure || !st->switch_icp);
}

// query i: 12-vector of T_total * [alpha R0 | beta p0]  (reference .cpp:450-453 after .cpp:713-716)
__device__ __forceinline__ void make_query(const SourceView& S, const RunConfig& cfg, const double* __restrict__ Tm, int i,
                                           double q[12]) {
    const size_t n = (size_t)S.n;
    double R0[9];
#pragma unroll
    for (int k = 0; k < 9; k++) R0[k] = S.frame[k * n + i] * cfg.alpha;
    double p[3] = {S.x[i] * cfg.beta, S.y[i] * cfg.beta, S.z[i] * cfg.beta};
#pragma unroll
    for (int c = 0; c < 3; c++) {  // column c of the rotation block
#pragma unroll
        for (int r = 0; r < 3; r++)
            q[3 * c + r] = Tm[4 * r] * R0[3 * c] + Tm[4 * r + 1] * R0[3 * c + 1] + Tm[4 * r + 2] * R0[3 * c + 2];
    }
#pragma unroll
    for (int r = 0; r < 3; r++) q[9 + r] = Tm[4 * r] * p[0] + Tm[4 * r + 1] * p[1] + Tm[4 * r + 2] * p[2] + Tm[4 * r + 3];
}

// |T X0 - T_ref X0|^2 = |(T - T_ref) X0|^2 for source element i: how far its query has moved since the iteration whose
// estimate was T_ref (coherence certificate).  Formed from the difference of the two estimates, so no second 12-vector is
// needed; it differs from the distance of the two rounded queries by ~1e-16 |q|, far inside the certificate's slack.
__device__ __forceinline__ double query_shift_sq(const SourceView& S, const RunConfig& cfg, const double* __restrict__ Tm,
                                                 const double* __restrict__ Tr, int i) {
    const size_t n = (size_t)S.n;
    double D[12];
#pragma unroll
    for (int k = 0; k < 12; k++) D[k] = Tm[k] - Tr[k];
    double dl = 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double a = S.frame[(3 * c) * n + i] * cfg.alpha, b = S.frame[(3 * c + 1) * n + i] * cfg.alpha,
                     d = S.frame[(3 * c + 2) * n + i] * cfg.alpha;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double v = D[4 * r] * a + D[4 * r + 1] * b + D[4 * r + 2] * d;
            dl += v * v;
        }
    }
    const double px = S.x[i] * cfg.beta, py = S.y[i] * cfg.beta, pz = S.z[i] * cfg.beta;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const double v = D[4 * r] * px + D[4 * r + 1] * py + D[4 * r + 2] * pz + D[4 * r + 3];
        dl += v * v;
    }
    return dl;
}

// exact FP64 squared distance between a query and Morton row j (sequential, non-contracted)
__device__ __forceinline__ double exact_d2_12(const double q[12], const double* __restrict__ rows64, size_t m, int j) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        double df = __dsub_rn(q[k], rows64[k * m + j]);
        s = __dadd_rn(s, __dmul_rn(df, df));
    }
    return s;
}

// same, query held in shared memory (frees 24 registers in the traversal kernel)
__device__ __forceinline__ double exact_d2_12_sm(const double* q, const double* __restrict__ rows64, size_t m, int j) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        double df = __dsub_rn(q[k], rows64[k * m + j]);
        s = __dadd_rn(s, __dmul_rn(df, df));
    }
    return s;
}

__device__ __forceinline__ void write_se3_match(const TargetView& T, const RunConfig& cfg, CorrBuffers& cb, int i,
                                                const double q[12], int j, double d2_12) {
    const size_t m = (size_t)T.n;
    // reference .cpp:465-467: stored distance is the 3-D distance of the translation columns
    // (target_se3_cloud_ column = beta * p even in the _with_cf variant, whose rows hold the unscaled p)
    double tx = T.rows64[9 * m + j] * T.dist_scale, ty = T.rows64[10 * m + j] * T.dist_scale,
           tz = T.rows64[11 * m + j] * T.dist_scale;
    double d3 = sqrt(sqdist3(q[9], q[10], q[11], tx, ty, tz));
    cb.idx[i] = T.perm12[j];
    store_distance(cfg, cb, i, d3);
    if (cb.d2_nd) cb.d2_nd[i] = d2_12;
}

// ------------------------------------------------------------------------------------------------
constexpr int kBfThreads = 128;
constexpr int kBfQ = 2;       // queries per thread
constexpr int kBfTile = 256;  // target rows per shared-memory tile

__global__ void __launch_bounds__(kBfThreads) nn_se3_brute_kernel(SourceView S, TargetView T, RunConfig cfg,
                                                                    IterState* __restrict__ state, CorrBuffers cb,
                                                                    int force_all) {
    if (state->done || !se3_phase_active(cfg, state)) return;
    __shared__ float4 tile[3][kBfTile];
    __shared__ double Tm[16];
    if (threadIdx.x < 16) Tm[threadIdx.x] = state->T_total[threadIdx.x];
    __syncthreads();

    const int M = T.n;
    float qf[kBfQ][12];
    float b1[kBfQ], b2[kBfQ];
    int i1[kBfQ];
    int qi[kBfQ];
#pragma unroll
    for (int u = 0; u < kBfQ; u++) {
        qi[u] = S.begin + blockIdx.x * (kBfThreads * kBfQ) + u * kBfThreads + threadIdx.x;
        b1[u] = b2[u] = 3.0e38f;
        i1[u] = 0;
        if (qi[u] < S.end) {
            double q[12];
            make_query(S, cfg, Tm, qi[u], q);
#pragma unroll
            for (int k = 0; k < 12; k++) qf[u][k] = (float)q[k];
        } else {
#pragma unroll
            for (int k = 0; k < 12; k++) qf[u][k] = 0.f;
        }
    }

    for (int base = 0; base < M; base += kBfTile) {
        for (int t = threadIdx.x; t < 3 * kBfTile; t += kBfThreads) {
            int plane = t / kBfTile, j = t - plane * kBfTile;
            int g = base + j;
            tile[plane][j] = g < M ? T.rows32[(size_t)plane * M + g] : make_float4(1e15f, 1e15f, 1e15f, 1e15f);
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < kBfTile; j++) {
            float4 a = tile[0][j], b = tile[1][j], c = tile[2][j];
#pragma unroll
            for (int u = 0; u < kBfQ; u++) {
                float e, d;
                e = qf[u][0] - a.x;  d = e * e;
                e = qf[u][1] - a.y;  d = fmaf(e, e, d);
                e = qf[u][2] - a.z;  d = fmaf(e, e, d);
                e = qf[u][3] - a.w;  d = fmaf(e, e, d);
                e = qf[u][4] - b.x;  d = fmaf(e, e, d);
                e = qf[u][5] - b.y;  d = fmaf(e, e, d);
                e = qf[u][6] - b.z;  d = fmaf(e, e, d);
                e = qf[u][7] - b.w;  d = fmaf(e, e, d);
                e = qf[u][8] - c.x;  d = fmaf(e, e, d);
                e = qf[u][9] - c.y;  d = fmaf(e, e, d);
                e = qf[u][10] - c.z; d = fmaf(e, e, d);
                e = qf[u][11] - c.w; d = fmaf(e, e, d);
                b2[u] = fminf(b2[u], fmaxf(d, b1[u]));
                i1[u] = d < b1[u] ? base + j : i1[u];
                b1[u] = fminf(b1[u], d);
            }
        }
        __syncthreads();
    }

    const double tgt_absmax = state->tgt_absmax;
#pragma unroll
    for (int u = 0; u < kBfQ; u++) {
        int i = qi[u];
        if (i >= S.end) continue;
        double q[12];
        make_query(S, cfg, Tm, i, q);
        double amax = 0.0;
#pragma unroll
        for (int k = 0; k < 12; k++) amax = fmax(amax, fabs(q[k]));
        // |s - d2| <= eps(s): FP32 rounding of the 24 inputs (delta per coordinate difference) plus
        // the rounding of the 24 FP32 operations; generous constants, see DESIGN.md "certification".
        double delta = (amax + tgt_absmax) * 5.9604644775390625e-08;  // 2^-24
        double s1 = b1[u], s2 = b2[u];
        double e1 = s1 * 1.9073486328125e-06 + 8.0 * delta * sqrt(s1) + 16.0 * delta * delta;
        double e2 = s2 * 1.9073486328125e-06 + 8.0 * delta * sqrt(s2) + 16.0 * delta * delta;
        bool certified = !force_all && (s2 - e2 > s1 + e1);
        if (certified) {
            write_se3_match(T, cfg, cb, i, q, i1[u], cb.d2_nd ? exact_d2_12(q, T.rows64, (size_t)M, i1[u]) : 0.0);
        } else {
            int pos = atomicAdd(&state->repair_count, 1);
            cb.repair[pos] = i;
        }
    }
}

int launch_nn_se3_brute(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state, CorrBuffers cb,
                        int force_all_repair, cudaStream_t st) {
    int per_block = kBfThreads * kBfQ;
    int g = (S.end - S.begin + per_block - 1) / per_block;
    if (g < 1) g = 1;
    nn_se3_brute_kernel<<<g, kBfThreads, 0, st>>>(S, T, cfg, state, cb, force_all_repair);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
constexpr int kRepairThreads = 256;

__global__ void __launch_bounds__(kRepairThreads) nn_se3_repair_kernel(SourceView S, TargetView T, RunConfig cfg,
                                                                        IterState* __restrict__ state, CorrBuffers cb) {
    if (state->done || !se3_phase_active(cfg, state)) return;
    __shared__ double Tm[16];
    __shared__ double sd[kRepairThreads / 32];
    __shared__ int sid[kRepairThreads / 32];
    __shared__ int sj[kRepairThreads / 32];
    if (threadIdx.x < 16) Tm[threadIdx.x] = state->T_total[threadIdx.x];
    __syncthreads();
    const int count = state->repair_count;
    const int M = T.n;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int r = blockIdx.x; r < count; r += gridDim.x) {
        const int i = cb.repair[r];
        double q[12];
        make_query(S, cfg, Tm, i, q);
        double best = inf;
        int best_id = 0x7fffffff, best_j = 0;
        for (int j = threadIdx.x; j < M; j += kRepairThreads) {
            double d2 = exact_d2_12(q, T.rows64, (size_t)M, j);
            if (d2 < best) {
                best = d2;
                best_j = j;
                best_id = T.perm12[j];
            } else if (d2 == best) {
                int id = T.perm12[j];
                if (id < best_id) {
                    best_id = id;
                    best_j = j;
                }
            }
        }
        // block argmin on (d2, original id)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double od = __shfl_xor_sync(SE3_FULL, best, o);
            int oi = __shfl_xor_sync(SE3_FULL, best_id, o);
            int oj = __shfl_xor_sync(SE3_FULL, best_j, o);
            if (od < best || (od == best && oi < best_id)) {
                best = od;
                best_id = oi;
                best_j = oj;
            }
        }
        int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) {
            sd[w] = best;
            sid[w] = best_id;
            sj[w] = best_j;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < kRepairThreads / 32; k++) {
                if (sd[k] < best || (sd[k] == best && sid[k] < best_id)) {
                    best = sd[k];
                    best_id = sid[k];
                    best_j = sj[k];
                }
            }
            write_se3_match(T, cfg, cb, i, q, best_j, best);
        }
        __syncthreads();
    }
}

int launch_nn_se3_repair(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state, CorrBuffers cb,
                         cudaStream_t st) {
    int cnt = S.end - S.begin;
    int g = cnt < 148 * 4 ? (cnt > 0 ? cnt : 1) : 148 * 4;
    nn_se3_repair_kernel<<<g, kRepairThreads, 0, st>>>(S, T, cfg, state, cb);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// row position of the target SE(3) row whose 6-D Morton key is closest to the query's (binary search over the sorted keys)
__device__ __forceinline__ int seed_position_se3(const TargetView& T, const RunConfig& cfg, const double q[12]) {
    double Ru[9];
    const double inv_a = cfg.alpha != 0.0 ? 1.0 / cfg.alpha : 0.0;
#pragma unroll
    for (int k = 0; k < 9; k++) Ru[k] = q[k] * inv_a;
    const double inv_t = T.tscale != 0.0 ? 1.0 / T.tscale : 0.0;
    const uint64_t key = se3_key(Ru, q[9] * inv_t, q[10] * inv_t, q[11] * inv_t, T.idx.bbox);
    int lo = 0, hi = T.n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (T.keys12[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo >= T.n ? T.n - 1 : lo;
}

__device__ __forceinline__ int seed_position_xyz(const CloudIndex& I, double qx, double qy, double qz) {
    const uint64_t key = morton63(qx, qy, qz, I.bbox);
    int lo = 0, hi = I.n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (I.keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo >= I.n ? I.n - 1 : lo;
}

// Per-query preparation of a correspondence pass, one THREAD per query (coalesced plane loads), run before both searches:
//  * coherence filter — settles every query whose remembered match is provably still its unique nearest neighbour and
//    appends the others to the work list the search kernels consume;
//  * seeding — an unsettled query gets a starting point for its search from the Morton order of the search structure
//    (key construction + a 17-step binary search, once per thread here instead of 32-fold redundantly in every lane of
//    the warp-per-query search kernel, which is left with a single, warm-started code path): always when it has no
//    remembered match (first pass of a run) and on the first ICP-phase pass (the remembered match is the 12-D one,
//    possibly far in position); and, while the estimate still moves a lot (||T_prev - T||_F > cfg.reseed_thr), whenever
//    the seed is closer than the remembered match — measured: with the remembered match alone the second pass of a
//    KITTI-size pair cost more than the cold first one.  The search is exact from any starting point.
#ifndef NN_FILTER_THREADS
#define NN_FILTER_THREADS 256
#endif
#ifndef NN_FILTER_BLOCKS
#define NN_FILTER_BLOCKS 3
#endif
__global__ void __launch_bounds__(NN_FILTER_THREADS, NN_FILTER_BLOCKS) nn_filter_kernel(SourceView S, TargetView T, RunConfig cfg,
                                                         IterState* __restrict__ state, CorrBuffers cb) {
    if (state->done) return;
    const bool se3 = se3_phase_active(cfg, state);
    const bool enabled = (se3 ? cfg.coherence : cfg.coherence_xyz) && state->T_change < cfg.coherence_thr;
    // first kernel of every iteration: file the estimate this iteration's queries are formed with (certificates recorded
    // now are checked against it in later iterations)
    if (blockIdx.x == 0 && threadIdx.x < 16 && cb.t_table && state->iter < cb.t_table_cap)
        cb.t_table[16 * (size_t)state->iter + threadIdx.x] = state->T_total[threadIdx.x];
    const int t = S.begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S.end) return;
    // spatially sorted processing order: the work list then hands neighbouring queries to neighbouring warps
    const int i = S.order ? S.order[t] : t;
    const double* Tm = state->T_total;
    const size_t m = (size_t)T.n;
    int prev = cb.idx[i];
    const bool have_prev = prev >= 0 && prev < T.n;
    const bool first_icp_pass = !se3 && cfg.has_se3 && state->iter == state->switch_iter;
    const bool want_seed = !have_prev || first_icp_pass || state->T_change > cfg.reseed_thr;
    // the first ICP-phase iteration still sees the 12-D references of the SE(3) phase: ignore them once
    const bool try_settle = enabled && have_prev && !first_icp_pass && cb.ref_d2nd[i] >= 0.0;  // (NaN / -1: no certificate)
    if (!try_settle && !want_seed) {
        if (enabled) cb.work[atomicAdd(&state->work_count, 1)] = i;
        return;
    }
    bool settled = false;
    if (se3) {
        double q[12];
        make_query(S, cfg, Tm, i, q);
        double d_prev = -1.0;
        if (try_settle) {
            // distance to the query the certificate was recorded for: same source element, that iteration's estimate
            const double dl = query_shift_sq(S, cfg, Tm, cb.t_table + 16 * (size_t)cb.ref_iter[i], i);
            const int j = T.inv12[prev];
            d_prev = exact_d2_12(q, T.rows64, m, j);
            if ((sqrt(d_prev) + sqrt(dl)) * (1.0 + 1e-12) + 1e-300 < cb.ref_d2nd[i]) {
                write_se3_match(T, cfg, cb, i, q, j, d_prev);
                settled = true;
            }
        }
        if (!settled && want_seed) {
            const int js = seed_position_se3(T, cfg, q);
            if (have_prev) {
                if (d_prev < 0.0) d_prev = exact_d2_12(q, T.rows64, m, T.inv12[prev]);
                if (exact_d2_12(q, T.rows64, m, js) < d_prev) cb.idx[i] = T.perm12[js];
            } else {
                cb.idx[i] = T.perm12[js];
            }
        }
    } else {
        const double px = S.x[i], py = S.y[i], pz = S.z[i];
        const double qx = Tm[0] * px + Tm[1] * py + Tm[2] * pz + Tm[3];
        const double qy = Tm[4] * px + Tm[5] * py + Tm[6] * pz + Tm[7];
        const double qz = Tm[8] * px + Tm[9] * py + Tm[10] * pz + Tm[11];
        double d_prev = -1.0;
        if (have_prev) d_prev = sqdist3(qx, qy, qz, T.idx.x[prev], T.idx.y[prev], T.idx.z[prev]);
        if (try_settle) {
            const double* Tr = cb.t_table + 16 * (size_t)cb.ref_iter[i];
            const double ex = (Tm[0] - Tr[0]) * px + (Tm[1] - Tr[1]) * py + (Tm[2] - Tr[2]) * pz + (Tm[3] - Tr[3]),
                         ey = (Tm[4] - Tr[4]) * px + (Tm[5] - Tr[5]) * py + (Tm[6] - Tr[6]) * pz + (Tm[7] - Tr[7]),
                         ez = (Tm[8] - Tr[8]) * px + (Tm[9] - Tr[9]) * py + (Tm[10] - Tr[10]) * pz + (Tm[11] - Tr[11]);
            const double d1 = sqrt(d_prev);
            if ((d1 + sqrt(ex * ex + ey * ey + ez * ez)) * (1.0 + 1e-12) + 1e-300 < cb.ref_d2nd[i]) {
                store_distance(cfg, cb, i, d1);
                if (cb.d2_nd) cb.d2_nd[i] = d_prev;
                settled = true;
            }
        }
        if (!settled && want_seed) {
            const int seed = T.idx.perm[seed_position_xyz(T.idx, qx, qy, qz)];
            if (!have_prev || sqdist3(qx, qy, qz, T.idx.x[seed], T.idx.y[seed], T.idx.z[seed]) < d_prev) cb.idx[i] = seed;
        }
    }
    if (enabled && !settled) cb.work[atomicAdd(&state->work_count, 1)] = i;
}

int launch_nn_filter(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state, CorrBuffers cb,
                     cudaStream_t st) {
    int g = (S.end - S.begin + NN_FILTER_THREADS - 1) / NN_FILTER_THREADS;
    if (g < 1) g = 1;
    nn_filter_kernel<<<g, NN_FILTER_THREADS, 0, st>>>(S, T, cfg, state, cb);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
constexpr int kTreeWarps = 4;  // round 2, SE(3)-phase search of a KITTI-size pair: 4 warps x 10 blocks per SM (48 registers) 2.06 ms,
                               // 4 x 8 (64 registers) 2.11, 2 x 16 2.08, 8 x 4 2.22, 16 x 2 2.40: queries differ in length, so small blocks
                               // (a block holds its slot until its slowest warp is done) and more resident warps win here

__device__ __forceinline__ void se3_tree_body(const SourceView& S, const TargetView& T, const RunConfig& cfg,
                                              IterState* __restrict__ state, CorrBuffers& cb,
                                              int2 (*stacks)[kStackEntries], const double* Tm, double (*qs)[12]) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // coherent iterations only search the queries the filter kernel could not settle (cb.work list)
    const bool coherent = cfg.coherence && state->T_change < cfg.coherence_thr;
    const int w = blockIdx.x * kTreeWarps + wib;
    int i;
    if (coherent) {
        if (w >= state->work_count) return;
        i = cb.work[w];
    } else {
        i = S.begin + w;
        if (i >= S.end) return;
    }
    const int M = T.n;
    const size_t m = (size_t)M, tn = (size_t)T.idx.total_nodes;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);

    NN_COUNT(0, 1);
    double* q = qs[wib];  // the FP64 query lives in shared memory; only its FP32 image stays in registers
    {
        double qr[12];
        make_query(S, cfg, Tm, i, qr);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 12; k++) q[k] = qr[k];
        }
        __syncwarp();
    }
    // FP32 query with one scalar margin covering its rounding: |q_k - qf_k| <= 2^-24 |q_k| <= qeps
    float qf[12];
    float qamax = 0.f;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        qf[k] = (float)q[k];
        qamax = fmaxf(qamax, fabsf(qf[k]));
    }
    const float qeps = qamax * 1.1920929e-07f;  // 2^-23 |q|max: twice the worst rounding of any coordinate
    // Rigorous FP32 lower bound of the squared distance from the query to a node's box.  Per dimension the gap
    // d_k = max(0, lo_k - qf_k, qf_k - hi_k) is rounded down; the true gap (for the FP64 query) is at least
    // (d_k - qeps)+, and sum (d_k - qeps)+^2 >= S - 2 qeps sum d_k >= S - 2 sqrt(12) qeps sqrt(S) with S = sum d_k^2
    // (Cauchy-Schwarz) >= k1 S - k2 (sqrt(S) <= (S + 1) / 2: no square root, traverse.cuh rounding_correction), so the
    // rounding of the query costs two operations per node instead of two per dimension.  S is accumulated with round-down
    // FMAs; k1 S - k2 grows with S, so the smaller computed S keeps the bound valid.  Two dimensions per 16-byte load.
    float qk1, qk2;
    rounding_correction(__fmul_ru(7.f, qeps), qk1, qk2);  // 7 qeps > 2 sqrt(12) qeps
    auto lb_fn = [&](int node) -> double {
        NN_COUNT(1, 1);
        const float4* b = reinterpret_cast<const float4*>(T.box12) + node;
        float acc = 0.f;
#pragma unroll
        for (int kk = 0; kk < 6; kk++) {
            const float4 v = b[(size_t)kk * tn];
            float d0 = fmaxf(fmaxf(__fsub_rd(v.x, qf[2 * kk]), __fsub_rd(qf[2 * kk], v.y)), 0.f);
            float d1 = fmaxf(fmaxf(__fsub_rd(v.z, qf[2 * kk + 1]), __fsub_rd(qf[2 * kk + 1], v.w)), 0.f);
            acc = __fmaf_rd(d0, d0, acc);  // exact d*d + acc rounded down: still a lower bound, one instruction
            acc = __fmaf_rd(d1, d1, acc);
        }
        return (double)fmaxf(0.f, __fsub_rd(__fmul_rd(acc, qk1), qk2));
    };

    // ---- temporal coherence (DESIGN.md "coherence filter") -----------------------------------------------
    // Once the estimate moves little (T_change below cfg.coherence_thr) the search also tracks the exact
    // SECOND-nearest distance d2nd of each query and remembers the query q_ref it was found for.  In later
    // iterations the triangle inequality gives |q - r| >= d2nd - |q - q_ref| for every row r other than the
    // remembered match, so if the remembered match is strictly closer than that bound it is still the unique
    // nearest neighbour and the traversal is skipped.  Exact: a failed test just falls through to the search.
    double tau = inf;  // best squared distance
    double b2 = inf;   // second-best squared distance (coherent mode)
    int best_id = 0x7fffffff, best_j = 0;
    int prev = cb.idx[i];
    const bool have_prev = prev >= 0 && prev < M;

    // FP32 pre-filter of the leaf rows: |d32 - d2| <= eps(d32) (rounding of the 24 inputs, delta per coordinate
    // difference, plus the 24 FP32 operations; same bound as the certified sweep, slightly inflated).  Only rows
    // that could still beat the current bound pay for the exact FP64 evaluation, which decides.
    const float fdelta = (qamax + (float)state->tgt_absmax) * 5.9604645e-08f * 1.0001f;
    int skip_leaf = -1;
    auto leaf_fn = [&](int leaf) {
        if (leaf == skip_leaf) return;
        NN_COUNT(2, 1);
        int p = leaf * 32 + lane;
        bool cand = false;
        if (p < M) {
            const float4 a = T.rows32[p], b = T.rows32[m + p], c = T.rows32[2 * m + p];
            float e, d32;
            e = qf[0] - a.x;  d32 = e * e;
            e = qf[1] - a.y;  d32 = fmaf(e, e, d32);
            e = qf[2] - a.z;  d32 = fmaf(e, e, d32);
            e = qf[3] - a.w;  d32 = fmaf(e, e, d32);
            e = qf[4] - b.x;  d32 = fmaf(e, e, d32);
            e = qf[5] - b.y;  d32 = fmaf(e, e, d32);
            e = qf[6] - b.z;  d32 = fmaf(e, e, d32);
            e = qf[7] - b.w;  d32 = fmaf(e, e, d32);
            e = qf[8] - c.x;  d32 = fmaf(e, e, d32);
            e = qf[9] - c.y;  d32 = fmaf(e, e, d32);
            e = qf[10] - c.z; d32 = fmaf(e, e, d32);
            e = qf[11] - c.w; d32 = fmaf(e, e, d32);
            float eps = d32 * 3.8146973e-06f + 9.f * fdelta * sqrtf(d32) + 20.f * fdelta * fdelta;
            cand = (double)(d32 - eps) <= (coherent ? b2 : tau);
        }
        if (__ballot_sync(SE3_FULL, cand) == 0u) return;
        NN_COUNT(3, __popc(__ballot_sync(SE3_FULL, cand)));
        double d2 = inf;
        int id = 0x7fffffff;
        if (cand) {
            d2 = exact_d2_12_sm(q, T.rows64, m, p);
            id = T.perm12[p];
        }
        if (!coherent) {
            // warm-started searches rarely improve: only reduce across the warp when some lane beats tau
            if (__ballot_sync(SE3_FULL, d2 < tau || (d2 == tau && id < best_id)) == 0u) return;
            double wd = d2;
            int wid = id;
            warp_argmin(wd, wid);
            unsigned who = __ballot_sync(SE3_FULL, id == wid && d2 == wd);
            tau = wd;
            best_id = wid;
            best_j = leaf * 32 + (__ffs(who) - 1);
            return;
        }
        // two nearest: nothing changes unless a lane beats the second-best bound (or ties the best)
        if (__ballot_sync(SE3_FULL, d2 < b2 || (d2 == tau && id < best_id)) == 0u) return;
        double wd = d2;
        int wid = id;
        warp_argmin(wd, wid);
        unsigned who = __ballot_sync(SE3_FULL, id == wid && d2 == wd);
        int win = __ffs(who) - 1;
        double m2 = warp_min(lane == win ? inf : d2);  // second smallest of this leaf
        if (wid == best_id) {
            b2 = fmin(b2, m2);  // the leaf holds the current best itself
        } else if (wd < tau || (wd == tau && wid < best_id)) {
            b2 = fmin(b2, fmin(tau, m2));  // old best and the leaf's runner-up compete for second place
            tau = wd;
            best_id = wid;
            best_j = leaf * 32 + win;
        } else {
            b2 = fmin(b2, wd);
        }
    };

    // warm start: every query has a remembered match, or the stand-in nn_filter_kernel seeded from the Morton order
    if (have_prev && !coherent) {
        best_j = T.inv12[prev];
        best_id = prev;
        tau = exact_d2_12_sm(q, T.rows64, m, best_j);
    } else if (have_prev) {
        const int first = T.inv12[prev] >> 5;  // the remembered match's own leaf: best and a first runner-up
        leaf_fn(first);
        skip_leaf = first;
    }
    // prune against the second-best bound when it is being tracked
    traverse_nodes<true, false>(T.idx, lb_fn, coherent ? b2 : tau, stacks[wib], lane, leaf_fn);
    if (lane == 0) {
        write_se3_match(T, cfg, cb, i, q, best_j, tau);
        if (cfg.coherence) {  // certificate for later iterations: second-nearest distance + the iteration it belongs to
            const int it = state->iter;
            const bool keep = coherent && it < cb.t_table_cap;
            cb.ref_d2nd[i] = keep ? sqrt(b2) : -1.0;
            if (keep) cb.ref_iter[i] = it;
        }
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int kXyzWarps = kTreeWarps;

__device__ __forceinline__ void xyz_body(const SourceView& S, const TargetView& T, const RunConfig& cfg,
                                         IterState* __restrict__ state, CorrBuffers& cb, int2 (*stacks)[kStackEntries],
                                         const double* Tm) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const bool coherent = cfg.coherence_xyz && state->T_change < cfg.coherence_thr;
    const int w = blockIdx.x * kXyzWarps + wib;
    int i;
    if (coherent) {
        if (w >= state->work_count) return;
        i = cb.work[w];
    } else {
        i = S.begin + w;
        if (i >= S.end) return;
    }
    const CloudIndex& I = T.idx;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);

    // source_moving_ point: T_total * p0 (the reference re-transforms the cloud every iteration, .cpp:541,706)
    const double px = S.x[i], py = S.y[i], pz = S.z[i];
    const double qx = Tm[0] * px + Tm[1] * py + Tm[2] * pz + Tm[3];
    const double qy = Tm[4] * px + Tm[5] * py + Tm[6] * pz + Tm[7];
    const double qz = Tm[8] * px + Tm[9] * py + Tm[10] * pz + Tm[11];

    double tau = inf, b2 = inf;
    int best = 0x7fffffff;
    int prev = cb.idx[i];
    const bool have_prev = prev >= 0 && prev < I.n;
    int skip_leaf = -1;
    auto leaf_fn = [&](int leaf) {
        if (leaf == skip_leaf) return;
        int p = leaf * 32 + lane;
        double d2 = inf;
        int id = 0x7fffffff;
        if (p < I.n) {
            d2 = sqdist3(qx, qy, qz, I.sx[p], I.sy[p], I.sz[p]);
            id = I.perm[p];
        }
        if (!coherent) {
            if (__ballot_sync(SE3_FULL, d2 < tau || (d2 == tau && id < best)) == 0u) return;
            warp_argmin(d2, id);
            tau = d2;
            best = id;
            return;
        }
        if (__ballot_sync(SE3_FULL, d2 < b2 || (d2 == tau && id < best)) == 0u) return;
        double wd = d2;
        int wid = id;
        warp_argmin(wd, wid);
        unsigned who = __ballot_sync(SE3_FULL, id == wid && d2 == wd);
        int win = __ffs(who) - 1;
        double m2 = warp_min(lane == win ? inf : d2);
        if (wid == best) {
            b2 = fmin(b2, m2);
        } else if (wd < tau || (wd == tau && wid < best)) {
            b2 = fmin(b2, fmin(tau, m2));
            tau = wd;
            best = wid;
        } else {
            b2 = fmin(b2, wd);
        }
    };

    // warm start: the previous match (a single candidate) ...
    if (have_prev) {
        tau = sqdist3(qx, qy, qz, I.x[prev], I.y[prev], I.z[prev]);
        best = prev;
    }
    // ... plus its whole leaf (one load finds it) whenever a runner-up is needed.  Every query has a previous match:
    // nn_filter_kernel seeds the first pass of a run from the Morton order, and the first ICP-phase pass starts from
    // the SE(3) phase's match, which is close in position as well.
    if (coherent && have_prev) {
        const int first = I.inv[prev] >> 5;
        leaf_fn(first);
        skip_leaf = first;
    }
    traverse_boxes<true>(I, qx, qy, qz, coherent ? b2 : tau, stacks[wib], lane, leaf_fn);

    if (lane == 0) {
        double d = sqrt(tau);  // .cpp:411
        cb.idx[i] = best;
        store_distance(cfg, cb, i, d);  // .cpp:413
        if (cb.d2_nd) cb.d2_nd[i] = tau;
        if (cfg.coherence_xyz) {
            const int it = state->iter;
            const bool keep = coherent && it < cb.t_table_cap;
            cb.ref_d2nd[i] = keep ? sqrt(b2) : -1.0;
            if (keep) cb.ref_iter[i] = it;
        }
    }
}

// One kernel for the correspondence stage of an iteration: the phase flag lives on the device, so the kernel
// picks the 12-D or the 3-D search itself (which = 0), instead of launching both and letting one early-out.
// which = 1 / 2 restricts it to the SE(3) / XYZ search (stage-level entry points, timing hook).
__global__ void __launch_bounds__(kTreeWarps * 32, 10) nn_search_kernel(SourceView S, TargetView T, RunConfig cfg,
                                                                        IterState* __restrict__ state, CorrBuffers cb,
                                                                        int which) {
    if (state->done) return;
    const bool se3 = se3_phase_active(cfg, state);
    if ((which == 1 && !se3) || (which == 2 && se3)) return;
    __shared__ int2 stacks[kTreeWarps][kStackEntries];
    __shared__ double Tm[16];
    __shared__ double qs[kTreeWarps][12];
    if (threadIdx.x < 16) Tm[threadIdx.x] = state->T_total[threadIdx.x];
    __syncthreads();
    if (se3)
        se3_tree_body(S, T, cfg, state, cb, stacks, Tm, qs);
    else
        xyz_body(S, T, cfg, state, cb, stacks, Tm);
}

static int launch_nn_search_which(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state,
                                  CorrBuffers cb, int which, cudaStream_t st) {
    if (T.idx.n_levels > 6) {
        set_last_error("cloud too large for the traversal stack");
        return SE3ICP_ERR_UNSUPPORTED;
    }
    int g = (S.end - S.begin + kTreeWarps - 1) / kTreeWarps;
    if (g < 1) g = 1;
    nn_search_kernel<<<g, kTreeWarps * 32, 0, st>>>(S, T, cfg, state, cb, which);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

#if NN_COUNT_VISITS
extern "C" int se3icp_debug_nn_visits(unsigned long long out[4], int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_nn_visits, sizeof(g_nn_visits));
    if (reset) {
        unsigned long long z[4] = {0, 0, 0, 0};
        cudaMemcpyToSymbol(g_nn_visits, z, sizeof(z));
    }
    return 0;
}
#endif

int launch_nn_search(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state, CorrBuffers cb,
                     cudaStream_t st) {
    return launch_nn_search_which(S, T, cfg, state, cb, 0, st);
}
int launch_nn_se3_tree(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state, CorrBuffers cb,
                       cudaStream_t st) {
    return launch_nn_search_which(S, T, cfg, state, cb, 1, st);
}
int launch_nn_xyz(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state, CorrBuffers cb,
                  cudaStream_t st) {
    return launch_nn_search_which(S, T, cfg, state, cb, 2, st);
}

}  // namespace se3
