// traverse.cuh — warp-cooperative traversal of the implicit 32-wide box hierarchy.
// One warp owns one query; each step a lane tests one child box, survivors are compacted onto a
// per-warp stack in shared memory with their lower bound so they can be re-tested against the
// shrinking search radius when popped.
#pragma once

#include "common.cuh"
#include "internal.h"

namespace se3 {

constexpr int kStackEntries = 192;  // >= 31 * levels + 1 for up to 6 levels (n <= 32^6)

// FP32 image of a 3-D query for the box tests: the coordinates rounded to float and a correction term covering that
// rounding (see box_lower_bound)
struct BoxQuery {
    float x, y, z, k1, k2;
};

// c > 2 sqrt(dims) eps bounds the effect of rounding the query to FP32 on sum d_k (see box_lower_bound); with
// sqrt(S) <= (S + 1) / 2 the correction c sqrt(S) needs no square root: S - c sqrt(S) >= S (1 - c/2) - c/2 = k1 S - k2.
__device__ __forceinline__ void rounding_correction(float c, float& k1, float& k2) {
    k2 = __fmul_ru(0.5f, c);
    k1 = __fsub_rd(1.0f, k2);
}

__device__ __forceinline__ BoxQuery make_box_query(double qx, double qy, double qz) {
    BoxQuery q;
    q.x = (float)qx, q.y = (float)qy, q.z = (float)qz;
    // |q_k - float(q_k)| <= 2^-24 |q_k| <= eps := 2^-23 max|q|; c > 2 sqrt(3) eps
    rounding_correction(__fmul_ru(3.5f * 1.1920929e-07f, fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fabsf(q.z))), q.k1, q.k2);
    return q;
}

// Rigorous FP32 lower bound of the squared distance from the (FP64) query to a node's box, the boxes being rounded
// outwards: per axis the gap d_k = max(0, lo_k - qf_k, qf_k - hi_k) is rounded down, the true gap is at least
// (d_k - eps)+, and sum (d_k - eps)+^2 >= S - 2 eps sum d_k >= S - 2 sqrt(3) eps sqrt(S) >= k1 S - k2 with S = sum d_k^2
// (rounding_correction).  k1 S - k2 grows with S, so accumulating S with round-down FMAs keeps the bound valid.
__device__ __forceinline__ double box_lower_bound(const CloudIndex& I, int node, const BoxQuery& q) {
    const float2* b = I.box + node;
    const size_t tn = (size_t)I.total_nodes;
    const float2 bx = b[0], by = b[tn], bz = b[2 * tn];
    const float dx = fmaxf(fmaxf(__fsub_rd(bx.x, q.x), __fsub_rd(q.x, bx.y)), 0.f);
    const float dy = fmaxf(fmaxf(__fsub_rd(by.x, q.y), __fsub_rd(q.y, by.y)), 0.f);
    const float dz = fmaxf(fmaxf(__fsub_rd(bz.x, q.z), __fsub_rd(q.z, bz.y)), 0.f);
    const float s = __fmaf_rd(dz, dz, __fmaf_rd(dy, dy, __fmul_rd(dx, dx)));
    return (double)fmaxf(0.f, __fsub_rd(__fmul_rd(s, q.k1), q.k2));
}

// `tau` is read on every test, so the leaf functor may shrink it while the traversal runs.
// lb_fn(node) returns a lower bound of the squared distance from the query to anything below the
// node; leaf_fn(leaf_id) is called by the whole warp (convergent).
// kWideStart: begin at CloudIndex::start_level (all of its nodes, a few independent rounds) instead of the root level.
// Pays when the first radius is loose and the upper boxes rarely prune (1-NN searches: -3 %); the kNN search, whose
// seeded radius prunes whole top-level subtrees, keeps the root start (+2 % otherwise).
// kLeafMask: children that are leaves are not pushed; the ballot mask of the passing ones is kept and they are visited
// straight from it, smallest bound first, each re-tested against the radius as it stands by then (one integer warp
// reduction + one SHFL per leaf instead of a stack store, a stack load and the decoding of an entry).  Measured: the kNN
// search gains 5 % (0.99 -> 0.97 ms per 119 k-point cloud with lowest lane first; highest lane first 1.02 ms; smallest
// bound first another 1 %); the 1-NN searches lose 4 % either way (more spills at their 48 registers), so they keep the
// stack for their leaves — ordering THEIR stack pushes by bound (most promising child on top) opens 3 % fewer leaves and
// changes a KITTI-size pair by -1 %, a 150-iteration bunny run by +2 %: not kept.
template <bool kWideStart, bool kLeafMask, class LbFn, class LeafFn>
__device__ __forceinline__ void traverse_nodes(const CloudIndex& I, LbFn&& lb_fn, const double& tau, int2* stack, int lane,
                                               LeafFn&& leaf_fn) {
    const double kSlack = 1.0 - 1e-12;  // never prune on a rounding-level difference
    int sp = 0;
    {
        // every node of the start level, 32 at a time: the rounds are independent, so their loads overlap
        const int top = kWideStart ? I.start_level : I.n_levels - 1;
        const int cnt = I.level_cnt[top];
        for (int base = 0; base < cnt; base += 32) {
            int c = base + lane;
            double lb = 0.0;
            bool ok = false;
            if (c < cnt) {
                lb = lb_fn(I.level_off[top] + c);
                ok = lb * kSlack <= tau;
            }
            unsigned m = __ballot_sync(SE3_FULL, ok);
            if (ok) stack[sp + __popc(m & ((1u << lane) - 1u))] = make_int2((top << 27) | c, __float_as_int(__double2float_rd(lb)));
            sp += __popc(m);
        }
        __syncwarp();
    }
    // leaf_fn keeps a single call site in both variants
    unsigned pend = 0u;
    int pend_base = 0;
    float pend_lb = 0.f;
    for (;;) {
        int leaf;
        if (kLeafMask && pend) {
            // nearest pending leaf first: what it contributes tightens the radius before the farther ones are re-tested
            // (bounds are non-negative floats, so their bit patterns order like the values; 0.885 -> 0.875 ms per cloud
            // against lowest lane first)
            const unsigned int key = ((pend >> lane) & 1u) ? ((__float_as_uint(pend_lb) & 0xffffffe0u) | (unsigned int)lane) : 0xffffffffu;
            const int src = (int)(__reduce_min_sync(SE3_FULL, key) & 31u);
            pend &= ~(1u << src);
            const float l = __shfl_sync(SE3_FULL, pend_lb, src);
            if ((double)l * kSlack > tau) continue;
            leaf = pend_base + src;
        } else {
            if (sp == 0) break;
            int2 e = stack[sp - 1];
            sp--;
            __syncwarp();
            if ((double)__int_as_float(e.y) * kSlack > tau) continue;
            int lvl = e.x >> 27, node = e.x & ((1 << 27) - 1);
            if (lvl == 0) {
                leaf = node;
            } else {
                int cl = lvl - 1;
                int c = node * 32 + lane;
                double lb = 0.0;
                bool ok = false;
                if (c < I.level_cnt[cl]) {
                    lb = lb_fn(I.level_off[cl] + c);
                    ok = lb * kSlack <= tau;
                }
                unsigned m = __ballot_sync(SE3_FULL, ok);
                if (kLeafMask && cl == 0) {
                    pend = m;
                    pend_base = node * 32;
                    pend_lb = __double2float_rd(lb);
                } else {
                    if (ok) stack[sp + __popc(m & ((1u << lane) - 1u))] = make_int2((cl << 27) | c, __float_as_int(__double2float_rd(lb)));
                    sp += __popc(m);
                    __syncwarp();
                }
                continue;
            }
        }
        leaf_fn(leaf);
    }
}

template <bool kWideStart, class LeafFn>
__device__ __forceinline__ void traverse_boxes(const CloudIndex& I, double qx, double qy, double qz, const double& tau,
                                               int2* stack, int lane, LeafFn&& leaf_fn) {
    const BoxQuery bq = make_box_query(qx, qy, qz);
    traverse_nodes<kWideStart, !kWideStart>(I, [&](int node) { return box_lower_bound(I, node, bq); }, tau, stack, lane, leaf_fn);
}

}  // namespace se3
