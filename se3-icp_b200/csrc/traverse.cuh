// traverse.cuh — warp-cooperative traversal of the implicit 32-wide box hierarchy.
// One warp owns one query; each step a lane tests one child box, survivors are compacted onto a
// per-warp stack in shared memory with their lower bound so they can be re-tested against the
// shrinking search radius when popped.
#pragma once

#include "common.cuh"
#include "internal.h"

namespace se3 {

constexpr int kStackEntries = 192;  // >= 31 * levels + 1 for up to 6 levels (n <= 32^6)

__device__ __forceinline__ double box_lower_bound(const CloudIndex& I, int node, double qx, double qy, double qz) {
    const float2* b = I.box + node;
    size_t tn = (size_t)I.total_nodes;
    const float2 bx = b[0], by = b[tn], bz = b[2 * tn];
    double lox = bx.x, loy = by.x, loz = bz.x;
    double hix = bx.y, hiy = by.y, hiz = bz.y;
    double dx = fmax(0.0, fmax(lox - qx, qx - hix));
    double dy = fmax(0.0, fmax(loy - qy, qy - hiy));
    double dz = fmax(0.0, fmax(loz - qz, qz - hiz));
    return dx * dx + dy * dy + dz * dz;
}

// `tau` is read on every test, so the leaf functor may shrink it while the traversal runs.
// lb_fn(node) returns a lower bound of the squared distance from the query to anything below the
// node; leaf_fn(leaf_id) is called by the whole warp (convergent).
// kWideStart: begin at CloudIndex::start_level (all of its nodes, a few independent rounds) instead of the root level.
// Pays when the first radius is loose and the upper boxes rarely prune (1-NN searches: -3 %); the kNN search, whose
// seeded radius prunes whole top-level subtrees, keeps the root start (+2 % otherwise).
template <bool kWideStart, class LbFn, class LeafFn>
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
    while (sp > 0) {
        int2 e = stack[sp - 1];
        sp--;
        __syncwarp();
        if ((double)__int_as_float(e.y) * kSlack > tau) continue;
        int lvl = e.x >> 27, node = e.x & ((1 << 27) - 1);
        if (lvl == 0) {
            leaf_fn(node);
            continue;
        }
        int cl = lvl - 1;
        int c = node * 32 + lane;
        double lb = 0.0;
        bool ok = false;
        if (c < I.level_cnt[cl]) {
            lb = lb_fn(I.level_off[cl] + c);
            ok = lb * kSlack <= tau;
        }
        unsigned m = __ballot_sync(SE3_FULL, ok);
        if (ok) stack[sp + __popc(m & ((1u << lane) - 1u))] = make_int2((cl << 27) | c, __float_as_int(__double2float_rd(lb)));
        sp += __popc(m);
        __syncwarp();
    }
}

template <bool kWideStart, class LeafFn>
__device__ __forceinline__ void traverse_boxes(const CloudIndex& I, double qx, double qy, double qz, const double& tau,
                                               int2* stack, int lane, LeafFn&& leaf_fn) {
    traverse_nodes<kWideStart>(I, [&](int node) { return box_lower_bound(I, node, qx, qy, qz); }, tau, stack, lane, leaf_fn);
}

}  // namespace se3
