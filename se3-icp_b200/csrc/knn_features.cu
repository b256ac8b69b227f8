// knn_features.cu — neighbourhood estimation.
//
// One warp per query point.  The exact k nearest neighbours (ascending by (d2, original index),
// FP64, same non-contracted arithmetic as the oracle) are collected in a per-warp shared-memory
// list, then the same warp derives, from prefixes of that one list,
//   - the TOLDI local reference frame        (reference .cpp:241-316, k = number_of_nn_for_LRF_)
//   - the Open3D-style unoriented normal     (reference .cpp:43,494,643, k = 20 / 30, incl. self)
//   - the GICP covariance Rx diag(eps,1,1) Rx^T (reference .cpp:4-14,45-51)
// replacing 2 x (kd-tree kNN + per-point OpenMP loop) of the reference with one pass per cloud.
#include "common.cuh"
#include "internal.h"
#include "knn_list.cuh"
#include "traverse.cuh"

namespace se3 {

#ifndef KNN_WARPS
#define KNN_WARPS 16
#endif
#ifndef KNN_BLOCKS
#define KNN_BLOCKS 2
#endif
// 16 warps x 2 blocks per SM at 64 registers: neighbouring queries share L1 lines, and 32 warps per SM hide the latency of
// the serial selection loops better than 24 do, spills included.  Per 119 k-point cloud: 0.826 ms against 12 x 2 (80
// registers) 0.876, 8 x 4 0.880, 10 x 3 0.882, 20 x 1 0.956 (re-swept at the end of round 2; same ranking as before the
// kernel changes of that round).  The macros exist for such sweeps (profiles/experiments/build_variant.sh).
constexpr int kKnnWarps = KNN_WARPS;
constexpr int kPool = 256;  // unsorted candidate pool per warp

#ifndef KNN_QPW
#define KNN_QPW 1
#endif
constexpr int kQpw = KNN_QPW;  // queries per warp between two block barriers (2 / 3: 1.01 / 1.15 ms against 0.885 ms for 1)
#ifndef KNN_SLACK
#define KNN_SLACK 12
#endif
// a pool shrink accepts any radius that keeps K .. K + kSlack candidates: the interpolated count search hits a narrow
// window in as few rounds as a wide one, and fewer survivors mean a tighter radius (per 119 k-point cloud, K = 90:
// kSlack 4 / 8 / 12 / 16 / 24 / 32 / 38 -> 0.829 / 0.826 / 0.828 / 0.830 / 0.836 / 0.845 / 0.852 ms)
constexpr int kSlack = KNN_SLACK;

struct KnnScratch {
    unsigned long long d[kPool];
    int id[kPool];
    int2 stack[kStackEntries];
};

// Final ordering of the <= 128 pool entries, fast path.  Each squared distance is mapped to a 24-bit integer,
// q = floor(d2 * (2^24 - 1) / max d2), which is monotone: q_a < q_b implies d2_a < d2_b.  The words (q << 7 | pool slot)
// go through a 32-bit bitonic network (one SHFL and one min/max per compare-exchange instead of three SHFLs and a
// 96-bit comparison), element e of the sorted sequence ending in lane e % 32, register e / 32.  Two equal q among the
// entries leave their order undecided (equal distances, or distances closer than 2^-24 of the radius): the function
// then returns false and the caller orders the pool exactly.  ~0.4 k instead of ~2.1 k instructions per query.
__device__ __forceinline__ bool sort_pool_quantised(const unsigned long long* pd, int pool, int lane, unsigned int (&w)[4]) {
    double e[4], dmax = 0.0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const int j = lane + 32 * t;
        e[t] = j < pool ? __longlong_as_double((long long)pd[j]) : 0.0;
        dmax = fmax(dmax, e[t]);
    }
    dmax = warp_max(dmax);
    if (!(dmax > 0.0) || !(dmax < 1e300)) return false;  // all distances zero (or not finite): nothing to quantise
    const double scale = 16777215.0 / dmax;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const int j = lane + 32 * t;
        unsigned int q = (unsigned int)fmin(e[t] * scale, 16777215.0);
        w[t] = j < pool ? ((q << 7) | (unsigned int)j) : 0xffffffffu;
    }
#pragma unroll
    for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {  // partner in the same lane: registers t and t ^ (j / 32)
                const int dt = j >> 5;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    if ((t & dt) == 0) {
                        const bool up = ((32 * t) & k) == 0;
                        unsigned int lo = min(w[t], w[t + dt]), hi = max(w[t], w[t + dt]);
                        w[t] = up ? lo : hi;
                        w[t + dt] = up ? hi : lo;
                    }
                }
            } else {
                const bool lower = (lane & j) == 0;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const bool up = ((lane + 32 * t) & k) == 0;
                    unsigned int o = __shfl_xor_sync(SE3_FULL, w[t], j);
                    w[t] = (lower == up) ? min(w[t], o) : max(w[t], o);
                }
            }
        }
    }
    // undecided order: neighbours of the sorted sequence with the same quantised key
    bool tie = false;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        unsigned int nxt = __shfl_down_sync(SE3_FULL, w[t], 1);
        unsigned int wrap = __shfl_sync(SE3_FULL, t < 3 ? w[t < 3 ? t + 1 : 3] : 0xffffffffu, 0);
        if (lane == 31) nxt = wrap;
        tie |= w[t] != 0xffffffffu && nxt != 0xffffffffu && (w[t] >> 7) == (nxt >> 7);
    }
    return __ballot_sync(SE3_FULL, tie) == 0u;
}

__global__ void __launch_bounds__(kKnnWarps * 32, KNN_BLOCKS) knn_features_kernel(CloudIndex I, FeatureArgs fa) {
    extern __shared__ __align__(16) unsigned char knn_smem[];  // dynamic: more than 48 KB from 12 warps per block on
    KnnScratch* scratch = reinterpret_cast<KnnScratch*>(knn_smem);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // query = Morton position s: warp w takes positions kQpw w .. kQpw w + kQpw - 1, or those entries of the compacted list
    // of a partial range.  Warps past the end redo the last query without writing, so the whole block reaches the
    // __syncthreads() of the batched eigen-solves below.
    const int n_queries = fa.active_count ? *fa.active_count : I.n;
    if (blockIdx.x * (kKnnWarps * kQpw) >= n_queries) return;  // block-uniform
    KnnScratch& W = scratch[wib];
    const int K = fa.K < I.n ? fa.K : I.n;
    const int n_leaves = I.level_cnt[0];
    const double inf = __longlong_as_double((long long)kInfKey);
    const size_t n = (size_t)I.n;
    const int cnt = K;  // a finished list holds the min(K, n) nearest, ascending, then padding
    const int cl = fa.k_lrf > 0 ? (fa.k_lrf < cnt ? fa.k_lrf : cnt) : 0;
    const int rz = cl / 3;
    const int cn = fa.k_nrm > 0 ? (fa.k_nrm < cnt ? fa.k_nrm : cnt) : 0;

    // per-block hand-over to the batched 3x3 eigen-solves (one thread per system instead of one redundant warp-wide
    // instruction stream per query); with kQpw > 1 the sorted neighbour positions and the radius wait in shared memory
    __shared__ double eig_in[kKnnWarps][kQpw][2][6];
    __shared__ double eig_out[kKnnWarps][kQpw][2][3];
    __shared__ int eig_valid[kKnnWarps][kQpw];
    __shared__ int nb_pos[kQpw > 1 ? kKnnWarps : 1][kQpw > 1 ? kQpw : 1][kQpw > 1 ? 128 : 1];
    __shared__ double nb_radius2[kKnnWarps][kQpw];

    // state of the query in flight
    int s = 0, self = 0;
    bool active = false;
    double qx = 0, qy = 0, qz = 0;
    double nx[4], ny[4], nz[4];  // neighbour coordinates, list position j = lane + 32 t

    auto select_query = [&](int qi) {
        const int w_raw = (blockIdx.x * kKnnWarps + wib) * kQpw + qi;
        const bool in_range = w_raw < n_queries;
        s = !in_range ? I.n - 1 : (fa.active_list ? fa.active_list[w_raw] : w_raw);
        qx = I.sx[s], qy = I.sy[s], qz = I.sz[s];
        self = I.perm[s];
        // sharded source: only the rank's own range of original indices is processed; block-uniform exit
        // is not possible (neighbouring Morton positions belong to different ranks), so foreign queries
        // fall through as inactive warps
        active = in_range && self >= fa.q_begin && self < fa.q_end;
    };
    auto gather_neighbours = [&](const int (&Li)[4]) {
#pragma unroll
        for (int t = 0; t < 4; t++) {
            int j = lane + 32 * t;
            nx[t] = ny[t] = nz[t] = 0.0;
            if (j < cnt && active) {
                int id = Li[t];  // Morton position: the gathers hit the sorted planes the search has just read
                nx[t] = I.sx[id];
                ny[t] = I.sy[id];
                nz[t] = I.sz[id];
            }
        }
    };

    // ---- search + second moments of one query ------------------------------------------------------------------------
    auto search_query = [&](int qi) {
        key_t Ld[4] = {kInfKey, kInfKey, kInfKey, kInfKey};
        int Li[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
        // Search phase: candidates within the current radius go to an UNSORTED pool in shared memory.  When the
        // pool fills up, a search on the distance value (compares + one integer warp reduction per step, no data
        // movement) finds a radius that keeps between K and K + kSlack of them and the pool is compacted.  Sorting happens
        // once, at the end, on the ~K survivors.
        int pool = 0;
        double tau = inf;
        // exact (distance, ORIGINAL index) cut of the pool (rare: ties); the pool itself carries Morton positions
        auto exact_trim = [&](int count, int keep) -> double {
            for (int t = lane; t < count; t += 32) W.id[t] = I.perm[W.id[t]];
            __syncwarp();
            const double r = knn_exact_trim(W.d, W.id, count, keep, lane);
            for (int t = lane; t < keep; t += 32) W.id[t] = I.inv[W.id[t]];
            __syncwarp();
            return r;
        };
        auto shrink_pool = [&]() {
            if (pool <= K + kSlack) return;
            double e[8];
            int eid[8];
            // Bounds of the search need not be tight: 0 below, the current radius above (every pool entry passed
            // d2 <= tau); on the first shrink (no radius yet) the largest high word + 1, one integer warp reduction.
            unsigned hmax = 0u;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                int t = lane + 32 * k;
                e[k] = inf;
                eid[k] = 0;
                if (t < pool) {
                    e[k] = __longlong_as_double((long long)W.d[t]);
                    eid[k] = W.id[t];
                    hmax = max(hmax, (unsigned)__double2hiint(e[k]));
                }
            }
            double lo = 0.0, hi = tau;
            if (!(tau < inf)) hi = __hiloint2double((int)(__reduce_max_sync(SE3_FULL, hmax) + 1u), 0);
            // Squared distances of points on a surface are spread almost uniformly, so interpolating the count (regula
            // falsi on the empirical distribution, aimed at the middle of the accepted window K .. K + kSlack) needs 2-3
            // rounds where halving the interval needs 6-8; every other round from the fourth on halves, which bounds
            // the worst case.  count(d <= hi) >= K holds throughout.
            int c_lo = 0, c_hi = pool;
            for (int it = 0; it < 24; it++) {
                double mid;
                if (it < 3 || (it & 1)) {
                    const float f = __fdividef((float)(K + kSlack / 2 - c_lo), (float)(c_hi - c_lo));
                    mid = fma(hi - lo, (double)f, lo);
                } else {
                    mid = 0.5 * (lo + hi);
                }
                if (!(mid > lo && mid < hi)) break;
                int c = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) c += e[k] <= mid ? 1 : 0;
                c = __reduce_add_sync(SE3_FULL, c);
                if (c < K) {
                    lo = mid;
                    c_lo = c;
                } else {
                    hi = mid;
                    c_hi = c;
                    if (c <= K + kSlack) break;
                }
            }
            __syncwarp();
            int out = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                bool keep = e[k] <= hi;  // padding is +inf and hi is finite
                unsigned m = __ballot_sync(SE3_FULL, keep);
                if (keep) {
                    int pos = out + __popc(m & ((1u << lane) - 1u));
                    W.d[pos] = (key_t)__double_as_longlong(e[k]);
                    W.id[pos] = eid[k];
                }
                out += __popc(m);
            }
            pool = out;
            tau = hi;
            __syncwarp();
            if (pool > K + kSlack) {  // ties at the threshold (pool >= K holds, so the K-th entry exists)
                tau = exact_trim(pool, K);
                pool = K;
            }
        };
        auto eval_leaf = [&](int leaf) {
            int p = leaf * 32 + lane;
            bool pass = false;
            double d2 = inf;
            if (p < I.n) {
                d2 = sqdist3(qx, qy, qz, I.sx[p], I.sy[p], I.sz[p]);
                pass = d2 <= tau;
            }
            unsigned m = __ballot_sync(SE3_FULL, pass);
            if (m == 0u) return;
            if (pass) {
                int pos = pool + __popc(m & ((1u << lane) - 1u));
                W.d[pos] = (key_t)__double_as_longlong(d2);
                W.id[pos] = p;  // Morton position: no index load here
            }
            pool += __popc(m);
            __syncwarp();
            if (pool > kPool - 32) shrink_pool();
        };

        // seed: the leaves around the query in Morton order give a near-final search radius
        const int L = s >> 5;
        // Window of 2 ceil(K / 32) + 1 leaves (7 for K = 90: 224 points, all the pool takes without shrinking in
        // between): its K-th distance is the first radius, and a tight one saves more traversal and pool work than the
        // extra coalesced leaves cost.  Measured per 119 k-point cloud, K = 90: 3 / 4 / 5 / 7 / 9 / 11 / 15 leaves
        // 0.936 / 0.924 / 0.875 / 0.835 / 0.843 / 0.850 / 0.876 ms.
#ifdef KNN_SEED_LEAVES
        const int nw = KNN_SEED_LEAVES;
#else
        const int nw = min(2 * ((K + 31) / 32) + 1, (kPool - 32) / 32);
#endif
        // (the window keeps its width at the ends of the Morton order)
        int w0 = L - nw / 2;
        if (w0 > n_leaves - nw) w0 = n_leaves - nw;
        if (w0 < 0) w0 = 0;
        const int w1 = w0 + nw - 1 < n_leaves - 1 ? w0 + nw - 1 : n_leaves - 1;
        if (active) {
            for (int leaf = w0; leaf <= w1; leaf++) eval_leaf(leaf);
            shrink_pool();
            if (pool >= K && !(tau < inf)) {  // K .. K + kSlack seeds: nothing to shrink, but they do bound the radius
                unsigned hmax = 0u;
                for (int t = lane; t < pool; t += 32) hmax = max(hmax, (unsigned)(W.d[t] >> 32));
                tau = __hiloint2double((int)(__reduce_max_sync(SE3_FULL, hmax) + 1u), 0);
            }
            traverse_boxes<false>(I, qx, qy, qz, tau, W.stack, lane, [&](int leaf) {
                if (leaf >= w0 && leaf <= w1) return;
                eval_leaf(leaf);
            });
            shrink_pool();
            // exact order of the survivors (K .. K + kSlack of them): the list is striped over the warp, element e in lane
            // e % 32, register e / 32
            if (pool > 128) {  // K + kSlack > 128: cut to exactly K first
                tau = exact_trim(pool, K);
                pool = K;
            }
            unsigned int w[4];
            if (sort_pool_quantised(W.d, pool, lane, w)) {
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    if (w[t] != 0xffffffffu) {
                        const int slot = (int)(w[t] & 127u);
                        Ld[t] = W.d[slot];
                        Li[t] = W.id[slot];
                    }
                }
            } else {  // equal quantised keys somewhere: the 96-bit network decides, sorted entries come back in the pool
                exact_trim(pool, pool);
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int j = lane + 32 * t;
                    if (j < pool) {
                        Ld[t] = W.d[j];
                        Li[t] = W.id[j];
                    }
                }
            }
            __syncwarp();  // the pool is reused by the warp's next query
        }

        if (fa.knn_idx && active) {
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int j = lane + 32 * t;
                if (j < fa.K) {
                    fa.knn_idx[(size_t)self * fa.K + j] = j < cnt ? I.perm[Li[t]] : -1;
                    if (fa.knn_d2) fa.knn_d2[(size_t)self * fa.K + j] = j < cnt ? __longlong_as_double((long long)Ld[t]) : -1.0;
                }
            }
        }
        if (fa.k_lrf <= 0 && fa.k_nrm <= 0) return;

        gather_neighbours(Li);
        if (kQpw > 1) {
#pragma unroll
            for (int t = 0; t < 4; t++) nb_pos[wib][qi][lane + 32 * t] = Li[t];
        }
        if (lane == 0) eig_valid[wib][qi] = active ? 1 : 0;
        if (fa.k_lrf > 0 && active) {
            key_t rad_bits;
            int rad_id;
            list_at(Ld, Li, cl - 1, rad_bits, rad_id);
            if (lane == 0) nb_radius2[wib][qi] = __longlong_as_double((long long)rad_bits);
            // .cpp:259-265 centroid of neighbours 1..rz-1 divided by rz
            double cx = 0, cy = 0, cz = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int j = lane + 32 * t;
                if (j >= 1 && j < rz) {
                    cx += nx[t];
                    cy += ny[t];
                    cz += nz[t];
                }
            }
            double inv_rz = 1.0 / (double)rz;
            cx = warp_sum(cx) * inv_rz;
            cy = warp_sum(cy) * inv_rz;
            cz = warp_sum(cz) * inv_rz;
            // .cpp:268-272 scatter of neighbours 1..rz
            double c6[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int j = lane + 32 * t;
                if (j >= 1 && j <= rz && j < cnt) {
                    double dx = nx[t] - cx, dy = ny[t] - cy, dz = nz[t] - cz;
                    c6[0] += dx * dx;
                    c6[1] += dx * dy;
                    c6[2] += dx * dz;
                    c6[3] += dy * dy;
                    c6[4] += dy * dz;
                    c6[5] += dz * dz;
                }
            }
#pragma unroll
            for (int e = 0; e < 6; e++) {
                double v = warp_sum(c6[e]);
                if (lane == 0) eig_in[wib][qi][0][e] = v;
            }
        }
        if (fa.k_nrm > 0 && cn >= 3 && active) {
            // Open3D ComputeCovariance: cumulants over the neighbourhood including the point itself
            double cu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int j = lane + 32 * t;
                if (j < cn) {
                    cu[0] += nx[t];
                    cu[1] += ny[t];
                    cu[2] += nz[t];
                    cu[3] += nx[t] * nx[t];
                    cu[4] += nx[t] * ny[t];
                    cu[5] += nx[t] * nz[t];
                    cu[6] += ny[t] * ny[t];
                    cu[7] += ny[t] * nz[t];
                    cu[8] += nz[t] * nz[t];
                }
            }
            double invc = 1.0 / (double)cn;
#pragma unroll
            for (int e = 0; e < 9; e++) cu[e] = warp_sum(cu[e]) * invc;
            if (lane == 0) {
                eig_in[wib][qi][1][0] = cu[3] - cu[0] * cu[0];
                eig_in[wib][qi][1][1] = cu[4] - cu[0] * cu[1];
                eig_in[wib][qi][1][2] = cu[5] - cu[0] * cu[2];
                eig_in[wib][qi][1][3] = cu[6] - cu[1] * cu[1];
                eig_in[wib][qi][1][4] = cu[7] - cu[1] * cu[2];
                eig_in[wib][qi][1][5] = cu[8] - cu[2] * cu[2];
            }
        }
    };

    // ---- frame / normal / covariance of one query from the solved eigenvectors --------------------------------------
    auto finish_query = [&](int qi) {
        if (!active) return;
        if (fa.k_lrf > 0) {
            const double radius = sqrt(nb_radius2[wib][qi]);  // .cpp:256
            double zx = eig_out[wib][qi][0][0], zy = eig_out[wib][qi][0][1], zz = eig_out[wib][qi][0][2];
            // .cpp:286-297
            double ax = 0, ay = 0, az = 0, wx = 0, wy = 0, wz = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int j = lane + 32 * t;
                if (j >= 1 && j < cl) {
                    double vx = nx[t] - qx, vy = ny[t] - qy, vz = nz[t] - qz;
                    ax += vx;
                    ay += vy;
                    az += vz;
                    double nd = zx * vx + zy * vy + zz * vz;
                    double an = sqrt(vx * vx + vy * vy + vz * vz);
                    double w = (radius - an) * (radius - an) * (nd * nd);
                    wx += w * vx;
                    wy += w * vy;
                    wz += w * vz;
                }
            }
            ax = warp_sum(ax);
            ay = warp_sum(ay);
            az = warp_sum(az);
            wx = warp_sum(wx);
            wy = warp_sum(wy);
            wz = warp_sum(wz);
            if (zx * ax + zy * ay + zz * az < 0.0) {  // .cpp:298
                zx = -zx;
                zy = -zy;
                zz = -zz;
            }
            double pd = wx * zx + wy * zy + wz * zz;  // .cpp:302-303
            double xx = wx - pd * zx, xy = wy - pd * zy, xz = wz - pd * zz;
            double inv = 1.0 / sqrt(xx * xx + xy * xy + xz * xz);
            xx *= inv;
            xy *= inv;
            xz *= inv;
            double yx = zy * xz - zz * xy, yy = zz * xx - zx * xz, yz = zx * xy - zy * xx;  // y = z x x (.cpp:306)
            if (lane == 0) {
                double* f = fa.frame + self;
                f[0] = xx, f[n] = xy, f[2 * n] = xz;
                f[3 * n] = yx, f[4 * n] = yy, f[5 * n] = yz;
                f[6 * n] = zx, f[7 * n] = zy, f[8 * n] = zz;
            }
        }

        if (fa.k_nrm > 0 && lane == 0) {
            double nvx = 0.0, nvy = 0.0, nvz = 1.0;  // fewer than 3 neighbours: Open3D's identity covariance -> (0,0,1)
            if (cn >= 3) {
                nvx = eig_out[wib][qi][1][0], nvy = eig_out[wib][qi][1][1], nvz = eig_out[wib][qi][1][2];
                if (nvx * nvx + nvy * nvy + nvz * nvz == 0.0) {
                    nvx = 0.0, nvy = 0.0, nvz = 1.0;
                }
            }
            if (fa.nrm) {
                fa.nrm[self] = nvx;
                fa.nrm[n + self] = nvy;
                fa.nrm[2 * n + self] = nvz;
            }
            if (fa.want_cov && fa.cov) {
                double C6[6];
                gicp_cov_from_normal(nvx, nvy, nvz, fa.gicp_eps, C6);
                double* o = fa.cov + self;
                for (int e = 0; e < 6; e++) o[e * n] = C6[e];
            }
        }
    };

#pragma unroll 1
    for (int qi = 0; qi < kQpw; qi++) {
        select_query(qi);
        search_query(qi);
    }
    if (fa.k_lrf <= 0 && fa.k_nrm <= 0) return;
    __syncthreads();
    if (threadIdx.x < 2 * kKnnWarps * kQpw) {
        const int which = threadIdx.x & 1, qi = (threadIdx.x >> 1) % kQpw, w = threadIdx.x / (2 * kQpw);
        if (eig_valid[w][qi] && ((which == 0 && fa.k_lrf > 0) || (which == 1 && fa.k_nrm > 0))) {
            double a6[6], ev[3], V[3][3];
#pragma unroll
            for (int e = 0; e < 6; e++) a6[e] = eig_in[w][qi][which][e];
            eig3_sym(a6, ev, V);  // .cpp:275-281 / Open3D ComputeNormal: eigenvector of the smallest eigenvalue
            eig_out[w][qi][which][0] = V[0][0];
            eig_out[w][qi][which][1] = V[1][0];
            eig_out[w][qi][which][2] = V[2][0];
        }
    }
    __syncthreads();
    if (kQpw == 1) {
        finish_query(0);  // the neighbour coordinates are still in registers
    } else {
#pragma unroll 1
        for (int qi = 0; qi < kQpw; qi++) {
            select_query(qi);
            int Li[4];
#pragma unroll
            for (int t = 0; t < 4; t++) Li[t] = nb_pos[wib][qi][lane + 32 * t];
            gather_neighbours(Li);
            finish_query(qi);
        }
    }
}

// Morton positions whose original index lies in [q_begin, q_end), compacted (order kept inside a warp, roughly kept
// across warps: neighbouring queries stay neighbours, which is all the locality the search wants)
__global__ void __launch_bounds__(256) knn_active_kernel(const int* __restrict__ perm, int n, int q_begin, int q_end,
                                                          int* __restrict__ list, int* __restrict__ count) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    bool on = false;
    if (s < n) {
        const int o = perm[s];
        on = o >= q_begin && o < q_end;
    }
    const unsigned m = __ballot_sync(SE3_FULL, on);
    if (m == 0u) return;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(SE3_FULL, base, __ffs(m) - 1);
    if (on) list[base + __popc(m & ((1u << lane) - 1u))] = s;
}

__global__ void __launch_bounds__(256) cov_from_normals_kernel(const double* __restrict__ nrm, int n, double eps,
                                                                double* __restrict__ cov) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double C6[6];
        gicp_cov_from_normal(nrm[i], nrm[(size_t)n + i], nrm[2 * (size_t)n + i], eps, C6);
        for (int e = 0; e < 6; e++) cov[(size_t)e * n + i] = C6[e];
    }
}

int launch_cov_from_normals(const double* nrm, int n, double eps, double* cov, cudaStream_t st) {
    int g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    cov_from_normals_kernel<<<g, 256, 0, st>>>(nrm, n, eps, cov);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

int launch_knn_features(const CloudIndex& I, const FeatureArgs& fa, cudaStream_t st) {
    if (fa.K > SE3ICP_MAX_KNN || fa.K <= 0) {
        set_last_error("kNN list length %d outside 1..%d", fa.K, SE3ICP_MAX_KNN);
        return SE3ICP_ERR_UNSUPPORTED;
    }
    if (I.n_levels > 6) {
        set_last_error("cloud too large for the traversal stack");
        return SE3ICP_ERR_UNSUPPORTED;
    }
    int blocks = (I.n + kKnnWarps * kQpw - 1) / (kKnnWarps * kQpw);
    const size_t smem = sizeof(KnnScratch) * kKnnWarps;
    static bool configured[64] = {false};  // function attributes are per device
    int dev = 0;
    SE3_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        SE3_CUDA(cudaFuncSetAttribute(knn_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = true;
    }
    FeatureArgs args = fa;
    const bool partial = fa.q_begin > 0 || fa.q_end < I.n;
    if (partial && fa.active_list && fa.active_count) {
        SE3_CUDA(cudaMemsetAsync(fa.active_count, 0, sizeof(int), st));
        knn_active_kernel<<<(I.n + 255) / 256, 256, 0, st>>>(I.perm, I.n, fa.q_begin, fa.q_end, fa.active_list, fa.active_count);
    } else {
        args.active_list = nullptr;
        args.active_count = nullptr;
    }
    knn_features_kernel<<<blocks, kKnnWarps * 32, smem, st>>>(I, args);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace se3
