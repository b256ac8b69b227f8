"""kNN+features kernel timing on the KITTI-like target cloud: python profiles/knn_time.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
import workloads as W
capi = graft.load_package().capi
src, tgt, T_gt = W.lidar_pair(seed=0)
ctx = capi.Context(0)
ctx.set_cloud(capi.SOURCE, src); ctx.set_cloud(capi.TARGET, tgt)
p = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, reuse_features=0, **W.KITTI_PARAMS)
T, st = ctx.run(p)
print("knn_features(target) %.3f ms; pair total %.2f ms" % (ctx.time_stage(capi.STAGE_KNN_TARGET, 10), st.time_total_ms))
