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
#include "traverse.cuh"

namespace se3 {

constexpr int kCap = 256;  // candidate buffer entries per warp
constexpr int kKnnWarps = 8;

struct KnnScratch {
    double d[kCap];
    int id[kCap];
    int2 stack[kStackEntries];
};

// ascending bitonic sort of (d, id) pairs, P a power of two <= kCap, executed by one warp
__device__ __forceinline__ void warp_sort_pairs(double* d, int* id, int P, int lane) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (P >> 1); t += 32) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int p = i | j;
                bool up = (i & k) == 0;
                double di = d[i], dp = d[p];
                int ii = id[i], ip = id[p];
                bool gt = di > dp || (di == dp && ii > ip);
                if (gt == up) {
                    d[i] = dp;
                    d[p] = di;
                    id[i] = ip;
                    id[p] = ii;
                }
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(kKnnWarps * 32) knn_features_kernel(CloudIndex I, FeatureArgs fa) {
    __shared__ KnnScratch scratch[kKnnWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int s = blockIdx.x * kKnnWarps + wib;  // query = Morton position s
    if (s >= I.n) return;
    KnnScratch& W = scratch[wib];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);

    const double qx = I.sx[s], qy = I.sy[s], qz = I.sz[s];
    const int self = I.perm[s];
    const int K = fa.K < I.n ? fa.K : I.n;
    int cnt = 0;
    double tau = inf;
    int tau_id = 0x7fffffff;
    const int n_leaves = I.level_cnt[0];

    auto eval_leaf = [&](int leaf, bool filter) {
        int p = leaf * 32 + lane;
        bool pass = false;
        double d2 = 0.0;
        int id = 0;
        if (p < I.n) {
            d2 = sqdist3(qx, qy, qz, I.sx[p], I.sy[p], I.sz[p]);
            id = I.perm[p];
            pass = !filter || d2 < tau || (d2 == tau && id < tau_id);
        }
        unsigned m = __ballot_sync(SE3_FULL, pass);
        if (pass) {
            int pos = cnt + __popc(m & ((1u << lane) - 1u));
            W.d[pos] = d2;
            W.id[pos] = id;
        }
        cnt += __popc(m);
        __syncwarp();
    };
    auto compact = [&]() {
        int P = 32;
        while (P < cnt) P <<= 1;
        for (int t = cnt + lane; t < P; t += 32) {
            W.d[t] = inf;
            W.id[t] = 0x7fffffff;
        }
        __syncwarp();
        warp_sort_pairs(W.d, W.id, P, lane);
        if (cnt > K) cnt = K;
        if (cnt == K) {
            tau = W.d[K - 1];
            tau_id = W.id[K - 1];
        }
        __syncwarp();
    };

    // seed: the leaves around the query in Morton order give a near-final search radius
    const int L = s >> 5;
    const int half = (K + 63) / 64;
    const int w0 = L - half > 0 ? L - half : 0;
    const int w1 = L + half < n_leaves - 1 ? L + half : n_leaves - 1;
    for (int leaf = w0; leaf <= w1; leaf++) eval_leaf(leaf, false);
    compact();

    traverse_boxes(I, qx, qy, qz, tau, W.stack, lane, [&](int leaf) {
        if (leaf >= w0 && leaf <= w1) return;
        eval_leaf(leaf, true);
        if (cnt > kCap - 32) compact();
    });
    compact();  // final: ascending, cnt = min(K, n)

    if (fa.knn_idx) {
        for (int j = lane; j < fa.K; j += 32) {
            fa.knn_idx[(size_t)self * fa.K + j] = j < cnt ? W.id[j] : -1;
            if (fa.knn_d2) fa.knn_d2[(size_t)self * fa.K + j] = j < cnt ? W.d[j] : -1.0;
        }
    }
    if (fa.k_lrf <= 0 && fa.k_nrm <= 0) return;

    // neighbour coordinates, list position j = lane + 32 t
    double nx[4], ny[4], nz[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
        int j = lane + 32 * t;
        nx[t] = ny[t] = nz[t] = 0.0;
        if (j < cnt) {
            int id = W.id[j];
            nx[t] = I.x[id];
            ny[t] = I.y[id];
            nz[t] = I.z[id];
        }
    }
    const size_t n = (size_t)I.n;

    if (fa.k_lrf > 0) {
        const int cl = fa.k_lrf < cnt ? fa.k_lrf : cnt;
        const int rz = cl / 3;
        const double radius = sqrt(W.d[cl - 1]);  // .cpp:256
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
        for (int e = 0; e < 6; e++) c6[e] = warp_sum(c6[e]);
        double ev[3], V[3][3];
        eig3_sym(c6, ev, V);  // .cpp:275-281
        double zx = V[0][0], zy = V[1][0], zz = V[2][0];
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

    if (fa.k_nrm > 0) {
        const int cn = fa.k_nrm < cnt ? fa.k_nrm : cnt;
        double nvx = 0.0, nvy = 0.0, nvz = 1.0;
        if (cn >= 3) {
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
            double c6[6];
            c6[0] = cu[3] - cu[0] * cu[0];
            c6[1] = cu[4] - cu[0] * cu[1];
            c6[2] = cu[5] - cu[0] * cu[2];
            c6[3] = cu[6] - cu[1] * cu[1];
            c6[4] = cu[7] - cu[1] * cu[2];
            c6[5] = cu[8] - cu[2] * cu[2];
            double ev[3], V[3][3];
            eig3_sym(c6, ev, V);
            nvx = V[0][0], nvy = V[1][0], nvz = V[2][0];
            if (nvx * nvx + nvy * nvy + nvz * nvz == 0.0) {
                nvx = 0.0, nvy = 0.0, nvz = 1.0;
            }
        }
        if (lane == 0) {
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
    }
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
    int blocks = (I.n + kKnnWarps - 1) / kKnnWarps;
    knn_features_kernel<<<blocks, kKnnWarps * 32, 0, st>>>(I, fa);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace se3
