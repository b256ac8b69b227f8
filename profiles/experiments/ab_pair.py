"""A/B probe with statistics: KITTI-like pair (N runs: median / min of the device time, set-up, SE(3)-phase search) and bunny
difficult pt2pt (median), plus a checksum of both transforms so that variants can be checked for bit-identical results.
    SE3ICP_LIB=se3-icp_b200/variants/libse3icp_X.so python profiles/experiments/ab_pair.py [runs]"""
import hashlib
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
import workloads as W

capi = graft.load_package().capi
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 25
src, tgt, _ = W.lidar_pair(seed=0)
ctx = capi.Context(0)
ctx.set_cloud(capi.SOURCE, src)
ctx.set_cloud(capi.TARGET, tgt)
p = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, reuse_features=0, **W.KITTI_PARAMS)
tot, setup, se3 = [], [], []
for r in range(runs + 3):
    T, st = ctx.run(p)
    if r >= 3:
        tot.append(st.time_total_ms)
        setup.append(st.time_setup_ms)
        se3.append(st.time_se3_phase_search_ms)
h1 = hashlib.md5(T.tobytes()).hexdigest()[:8]
it1 = st.num_iterations
src, tgt, _ = W.bunny_problem("difficult", seed=2)
ctx.set_cloud(capi.SOURCE, src)
ctx.set_cloud(capi.TARGET, tgt)
p = capi.default_params(variant="pt2pt", entry=capi.RUN_SE3_ICP, reuse_features=0, estimated_overlap=1.0,
                        max_num_se3_iterations=10, mse=1e-5, mse_switch_error=5e-5, number_of_nn_for_LRF=90)
bun = []
for r in range(9):
    T, st = ctx.run(p)
    if r >= 2:
        bun.append(st.time_total_ms)
h2 = hashlib.md5(T.tobytes()).hexdigest()[:8]
print("pair med %.3f min %.3f ms (setup %.3f, se3 search %.3f, %d it, T %s); bunny med %.2f min %.2f ms (%d it, T %s)" %
      (statistics.median(tot), min(tot), statistics.median(setup), statistics.median(se3), it1, h1,
       statistics.median(bun), min(bun), st.num_iterations, h2))
