"""Bunny difficult pt2pt and the KITTI-like pair with the coherence filter on / off (params.nn_coherence)."""
import os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
import workloads as W
capi = graft.load_package().capi
ctx = capi.Context(0)
src, tgt, _ = W.bunny_problem("difficult", seed=2)
ctx.set_cloud(capi.SOURCE, src); ctx.set_cloud(capi.TARGET, tgt)
for coh in (1, 0, 1, 0):
    p = capi.default_params(variant="pt2pt", entry=capi.RUN_SE3_ICP, reuse_features=0, estimated_overlap=1.0, max_num_se3_iterations=10,
                            mse=1e-5, mse_switch_error=5e-5, number_of_nn_for_LRF=90, nn_coherence=coh)
    t = []
    for r in range(6):
        T, st = ctx.run(p)
        if r: t.append(st.time_total_ms)
    print("bunny coherence=%d: med %.2f ms, corr %.2f ms, %d it, searched %d" % (coh, statistics.median(t), st.time_se3_correspondence_search_ms, st.num_iterations, st.queries_searched))
src, tgt, _ = W.lidar_pair(seed=0)
ctx.set_cloud(capi.SOURCE, src); ctx.set_cloud(capi.TARGET, tgt)
for coh in (1, 0):
    p = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, reuse_features=0, nn_coherence=coh, **W.KITTI_PARAMS)
    t = []
    for r in range(8):
        T, st = ctx.run(p)
        if r > 1: t.append(st.time_total_ms)
    print("pair coherence=%d: med %.3f ms, %d it, searched %d" % (coh, statistics.median(t), st.num_iterations, st.queries_searched))
