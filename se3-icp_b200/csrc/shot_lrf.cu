// shot_lrf.cu — SHOT local reference frame with radius support.
//
// Replaces reference .cpp:121-224 (computeSingleSHOTSE3Frame) and its OpenMP loop .cpp:226-239: the alternative to the
// TOLDI frame that the reference keeps next to it (its calls are commented out at .cpp:593-594,812-813; `lrf_radius_`
// .cpp:340 is its only parameter).  One warp per point, three traversals of the cloud's box hierarchy with the fixed
// squared radius instead of a sorted radius search:
//   1. weighted scatter  M = sum (r - d_i) a_i a_i^T / sum (r - d_i)  over the support d_i^2 < r^2, a_i = p_i - p,
//      without the point itself (the reference skips entry 0 of the sorted result, .cpp:151);
//   2. x+ / z+ = eigenvectors of the largest / smallest eigenvalue; votes  #{a_i . v >= 0}  for both;
//   3. only on an exact tie of a vote (.cpp:189,203): the five support points around the median DISTANCE vote.  Their
//      ranks are found by counting traversals (how many support points lie within t?) that narrow a distance window to
//      fewer than 256 points, which are then collected and ordered exactly by (distance, index).
// Fewer than 5 support points: undefined in the reference (a warning, then 0 / 0 or an out-of-range read); identity here
// and in the oracle.
#include "common.cuh"
#include "internal.h"
#include "knn_list.cuh"
#include "traverse.cuh"

namespace se3 {

constexpr int kShotWarps = 8;
constexpr int kShotPool = 256;
constexpr int kShotWindow = 100;  // slack of the rank window on either side: 2 * 100 + 5 candidates fit the pool

struct ShotScratch {
    unsigned long long d[kShotPool];
    int id[kShotPool];
    int2 stack[kStackEntries];
};

__global__ void __launch_bounds__(kShotWarps * 32) shot_lrf_kernel(CloudIndex I, double radius, double* __restrict__ frame,
                                                                     int* __restrict__ unresolved) {
    __shared__ ShotScratch scratch[kShotWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int s = blockIdx.x * kShotWarps + wib;  // Morton position: neighbouring warps read the same leaves
    if (s >= I.n) return;
    ShotScratch& W = scratch[wib];
    const double qx = I.sx[s], qy = I.sy[s], qz = I.sz[s];
    const int self = I.perm[s];
    const size_t n = (size_t)I.n;
    const double r2 = __dmul_rn(radius, radius);
    const double tau = r2;  // the traversal prunes on lower bound <= tau; membership (strict) is decided per point

    // every support point: fn(position, squared distance), called by the lanes that hold one
    auto for_support = [&](auto&& fn) {
        traverse_boxes<false>(I, qx, qy, qz, tau, W.stack, lane, [&](int leaf) {
            const int p = leaf * 32 + lane;
            if (p < I.n) {
                const double d2 = sqdist3(qx, qy, qz, I.sx[p], I.sy[p], I.sz[p]);
                if (d2 < r2) fn(p, d2);
            }
        });
        __syncwarp();
    };

    // ---- 1: weighted scatter (.cpp:151-157); the point itself is the skipped entry 0 (coincident duplicates contribute
    //         a = 0 with weight r whichever of them the reference skips)
    double c6[6] = {0, 0, 0, 0, 0, 0}, total = 0.0;
    int cnt = 0;
    for_support([&](int p, double d2) {
        if (p == s) return;
        const double w = radius - sqrt(d2);
        const double ax = I.sx[p] - qx, ay = I.sy[p] - qy, az = I.sz[p] - qz;
        c6[0] += w * ax * ax;
        c6[1] += w * ax * ay;
        c6[2] += w * ax * az;
        c6[3] += w * ay * ay;
        c6[4] += w * ay * az;
        c6[5] += w * az * az;
        total += w;
        cnt++;
    });
    const int n_considered = __reduce_add_sync(SE3_FULL, cnt);
    double xx = 1, xy = 0, xz = 0, zx = 0, zy = 0, zz = 1;
    if (n_considered >= 5) {
        total = warp_sum(total);
        double a6[6], ev[3], V[3][3];
#pragma unroll
        for (int e = 0; e < 6; e++) a6[e] = warp_sum(c6[e]) / total;
        eig3_sym(a6, ev, V);
        xx = V[0][2], xy = V[1][2], xz = V[2][2];  // largest eigenvalue  (.cpp:169)
        zx = V[0][0], zy = V[1][0], zz = V[2][0];  // smallest eigenvalue (.cpp:170)

        // ---- 2: votes (.cpp:172-179)
        int px = 0, pz = 0;
        for_support([&](int p, double) {
            if (p == s) return;
            const double ax = I.sx[p] - qx, ay = I.sy[p] - qy, az = I.sz[p] - qz;
            px += (ax * xx + ay * xy + az * xz >= 0.0) ? 1 : 0;
            pz += (ax * zx + ay * zy + az * zz >= 0.0) ? 1 : 0;
        });
        int sx_vote = 2 * __reduce_add_sync(SE3_FULL, px) - n_considered;
        int sz_vote = 2 * __reduce_add_sync(SE3_FULL, pz) - n_considered;

        // ---- 3: exact tie -> the neighbours diff[median - 2 .. median + 2], median = n_considered / 2, of the list sorted
        //         by distance vote (.cpp:189-197,203-211).  diff[j] is entry j + 1 of the sorted support (entry 0 = the
        //         point itself), so the wanted entries are F[R0 .. R0 + 4] with R0 = n_considered / 2 - 1.
        if (sx_vote == 0 || sz_vote == 0) {
            const int total_cnt = n_considered + 1;  // the support including the point itself
            const int R0 = n_considered / 2 - 1;
            // number of support points with squared distance <= t
            auto count_le = [&](double t) {
                int c = 0;
                for_support([&](int, double d2) { c += d2 <= t ? 1 : 0; });
                return __reduce_add_sync(SE3_FULL, c);
            };
            // a threshold t in (a, b) with want_lo <= count_le(t) <= want_hi; false when the counts jump over the window
            // (more than ~100 support points at exactly one distance)
            auto find_threshold = [&](double a, int ca, double b, int cb, int want_lo, int want_hi, double& t, int& ct) {
                for (int it = 0; it < 80; it++) {
                    double mid;
                    if (it & 1) {
                        mid = 0.5 * (a + b);
                    } else {
                        const float f = __fdividef((float)((want_lo + want_hi) / 2 - ca), (float)(cb - ca));
                        mid = fma(b - a, (double)fminf(fmaxf(f, 0.02f), 0.98f), a);
                    }
                    if (!(mid > a && mid < b)) return false;
                    const int c = count_le(mid);
                    if (c < want_lo) {
                        a = mid, ca = c;
                    } else if (c > want_hi) {
                        b = mid, cb = c;
                    } else {
                        t = mid, ct = c;
                        return true;
                    }
                }
                return false;
            };
            bool ok = true;
            double lo = -1.0, hi = r2;  // window (lo, hi]: nothing is <= -1, everything in the support is < r2
            int c_lo = 0;
            if (R0 > kShotWindow) ok = find_threshold(0.0, 0, r2, total_cnt, R0 - kShotWindow, R0, lo, c_lo);
            if (ok && total_cnt - c_lo > 2 * kShotWindow + 5) {
                int c_hi;
                ok = find_threshold(lo > 0.0 ? lo : 0.0, c_lo, r2, total_cnt, R0 + 5, R0 + 5 + kShotWindow, hi, c_hi);
            }
            int pool = 0;
            if (ok) {
                // collect the window; the traversal calls the leaf functor convergently, so the ballot is safe here
                traverse_boxes<false>(I, qx, qy, qz, tau, W.stack, lane, [&](int leaf) {
                    const int p = leaf * 32 + lane;
                    bool in = false;
                    double d2 = 0.0;
                    if (p < I.n) {
                        d2 = sqdist3(qx, qy, qz, I.sx[p], I.sy[p], I.sz[p]);
                        in = d2 < r2 && d2 > lo && d2 <= hi;
                    }
                    const unsigned m = __ballot_sync(SE3_FULL, in);
                    if (in) {
                        const int pos = pool + __popc(m & ((1u << lane) - 1u));
                        if (pos < kShotPool) {
                            W.d[pos] = (unsigned long long)__double_as_longlong(d2);
                            W.id[pos] = I.perm[p];  // ORIGINAL index: ties in distance resolve to the smaller one
                        }
                    }
                    pool += __popc(m);
                });
                __syncwarp();
                ok = pool <= kShotPool && R0 - c_lo >= 0 && R0 - c_lo + 5 <= pool && R0 - c_lo + 5 <= 128;
            }
            if (ok) {
                const int keep = pool < 128 ? pool : 128;
                knn_exact_trim(W.d, W.id, pool, keep, lane);  // pool[0 .. keep) ascending by (distance, index)
                int vx = 0, vz = 0;
                if (lane < 5) {
                    const int id = W.id[R0 - c_lo + lane];
                    const double ax = I.x[id] - qx, ay = I.y[id] - qy, az = I.z[id] - qz;
                    vx = (ax * xx + ay * xy + az * xz >= 0.0) ? 1 : 0;
                    vz = (ax * zx + ay * zy + az * zz >= 0.0) ? 1 : 0;
                }
                vx = __reduce_add_sync(SE3_FULL, vx);
                vz = __reduce_add_sync(SE3_FULL, vz);
                if (sx_vote == 0) sx_vote = vx < 3 ? -1 : 1;  // .cpp:194-196
                if (sz_vote == 0) sz_vote = vz < 3 ? -1 : 1;  // .cpp:208-210
            } else if (lane == 0 && unresolved) {
                atomicAdd(unresolved, 1);  // the tie stays unresolved (axis kept as the solver returned it)
            }
        }
        if (sx_vote < 0) xx = -xx, xy = -xy, xz = -xz;
        if (sz_vote < 0) zx = -zx, zy = -zy, zz = -zz;
    }
    if (lane == 0) {
        const double yx = zy * xz - zz * xy, yy = zz * xx - zx * xz, yz = zx * xy - zy * xx;  // y = z x x (.cpp:216)
        double* f = frame + self;
        if (n_considered < 5) {
            f[0] = 1, f[n] = 0, f[2 * n] = 0, f[3 * n] = 0, f[4 * n] = 1, f[5 * n] = 0, f[6 * n] = 0, f[7 * n] = 0, f[8 * n] = 1;
        } else {
            f[0] = xx, f[n] = xy, f[2 * n] = xz;
            f[3 * n] = yx, f[4 * n] = yy, f[5 * n] = yz;
            f[6 * n] = zx, f[7 * n] = zy, f[8 * n] = zz;
        }
    }
}

int launch_shot_lrf(const CloudIndex& I, double radius, double* frame, int* unresolved, cudaStream_t st) {
    if (!(radius > 0.0)) {
        set_last_error("SHOT frame: radius must be positive");
        return SE3ICP_ERR_ARG;
    }
    if (I.n_levels > 6) {
        set_last_error("cloud too large for the traversal stack");
        return SE3ICP_ERR_UNSUPPORTED;
    }
    shot_lrf_kernel<<<(I.n + kShotWarps - 1) / kShotWarps, kShotWarps * 32, 0, st>>>(I, radius, frame, unresolved);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace se3
