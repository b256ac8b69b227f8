"""Boxes / leaves the 12-D search opens per query (needs a library built with -DNN_COUNT_VISITS=1):
    SE3ICP_LIB=se3-icp_b200/variants/libse3icp_count.so [SE3ICP_SE3_ORDER=morton] python profiles/experiments/visits.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
import workloads as W
capi = graft.load_package().capi
lib = capi.lib()
out = (C.c_ulonglong * 4)()
HAVE = hasattr(lib, "se3icp_debug_nn_visits")
def report(tag):
    if not HAVE:
        print(tag, "(library built without NN_COUNT_VISITS)")
        return
    lib.se3icp_debug_nn_visits(out, 1)
    q = max(out[0], 1)
    print("%s: %d searches, %.2f node tests, %.2f leaves, %.2f exact rows per search" % (tag, out[0], out[1] / q, out[2] / q, out[3] / q))
src, tgt, _ = W.lidar_pair(seed=0)
ctx = capi.Context(0)
ctx.set_cloud(capi.SOURCE, src); ctx.set_cloud(capi.TARGET, tgt)
for it in (1, 2, 3, 10):
    p = capi.default_params(variant="gicp", entry=capi.RUN_SE3_PURE, reuse_features=1, use_graph=0, **dict(W.KITTI_PARAMS, max_num_se3_iterations=it, max_num_iterations=it))
    if HAVE:
        lib.se3icp_debug_nn_visits(out, 1)
    T, st = ctx.run(p)
    report("lidar pair, SE(3) iterations 1..%d (%d run)" % (it, st.num_iterations))
