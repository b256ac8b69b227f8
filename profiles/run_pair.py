"""One KITTI-like se3_gicp registration on cuda:0 (the unit of bench.py's workload), for ncu captures:
    python profiles/run_pair.py [n_runs]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import workloads as W  # noqa: E402

capi = graft.load_package().capi
src, tgt, T_gt = W.lidar_pair(seed=0)
ctx = capi.Context(0)
ctx.set_cloud(capi.SOURCE, src)
ctx.set_cloud(capi.TARGET, tgt)
p = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, reuse_features=0, **W.KITTI_PARAMS)  # every run is a full run
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    T, st = ctx.run(p)
print("pair: %d/%d points, %d iterations (%d SE3), %.2f ms total, %.2f ms setup, %d launches, rot err %.2e rad" %
      (len(src), len(tgt), st.num_iterations, st.num_pure_se3_iterations, st.time_total_ms, st.time_setup_ms,
       st.kernel_launches, W.rotation_error(T, T_gt)))
print("SE(3)-phase search %.3f ms, correspondence stage total %.3f ms" % (st.time_se3_phase_search_ms, st.time_se3_correspondence_search_ms))
print("queries searched: %d of %d (%.1f %%)" % (st.queries_searched, st.num_iterations * len(src), 100.0 * st.queries_searched / (st.num_iterations * len(src))))
