"""A/B probe: KITTI-like pair (5 runs, last) and bunny difficult pt2pt (3 runs, last) on one context."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
import workloads as W
capi = graft.load_package().capi
src, tgt, T_gt = W.lidar_pair(seed=0)
ctx = capi.Context(0)
ctx.set_cloud(capi.SOURCE, src); ctx.set_cloud(capi.TARGET, tgt)
p = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, reuse_features=0, **W.KITTI_PARAMS)
for _ in range(5):
    T, st = ctx.run(p)
print("pair %.2f ms (setup %.2f, se3 search %.3f, corr %.3f, %d it)" % (st.time_total_ms, st.time_setup_ms, st.time_se3_phase_search_ms, st.time_se3_correspondence_search_ms, st.num_iterations), end="; ")
src, tgt, T_gt = W.bunny_problem("difficult", seed=2)
ctx.set_cloud(capi.SOURCE, src); ctx.set_cloud(capi.TARGET, tgt)
p = capi.default_params(variant="pt2pt", entry=capi.RUN_SE3_ICP, reuse_features=0, estimated_overlap=1.0, max_num_se3_iterations=10, mse=1e-5, mse_switch_error=5e-5, number_of_nn_for_LRF=90)
for _ in range(3):
    T, st = ctx.run(p)
print("bunny %.2f ms (corr %.2f, %d it)" % (st.time_total_ms, st.time_se3_correspondence_search_ms, st.num_iterations))
