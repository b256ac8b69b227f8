"""Per-kernel CUDA-event timings on one KITTI-like pair (live, not under a profiler):
    python profiles/stage_times.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import workloads as W  # noqa: E402

capi = graft.load_package().capi
src, tgt, T_gt = W.lidar_pair(seed=0)
ctx = capi.Context(0)
ctx.set_cloud(capi.SOURCE, src)
ctx.set_cloud(capi.TARGET, tgt)
p = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, **W.KITTI_PARAMS)
for _ in range(3):
    T, st = ctx.run(p)
print("pair %d/%d: %d it (%d SE3) total %.2f ms setup %.2f ms launches %d" %
      (len(src), len(tgt), st.num_iterations, st.num_pure_se3_iterations, st.time_total_ms, st.time_setup_ms, st.kernel_launches))
for coh in (0, 1):
    pc = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, nn_coherence=coh, **W.KITTI_PARAMS)
    for _ in range(3):
        Tc, sc = ctx.run(pc)
    print("  nn_coherence=%d: total %.2f ms, identical: %s" % (coh, sc.time_total_ms, bool((Tc == T).all())))
pg = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, use_graph=1, **W.KITTI_PARAMS)
for _ in range(3):
    Tg, sg = ctx.run(pg)
print("  with use_graph=1: total %.2f ms, identical result: %s, launches %d" % (sg.time_total_ms, bool((Tg == T).all()), sg.kernel_launches))
for name, sid, rep in (("nn_se3", capi.STAGE_NN_SE3, 20), ("nn_xyz", capi.STAGE_NN_XYZ, 20), ("reduce", capi.STAGE_REDUCE, 20),
                       ("knn_features(target)", capi.STAGE_KNN_TARGET, 5)):
    print("  %-22s %.3f ms" % (name, ctx.time_stage(sid, rep)))
