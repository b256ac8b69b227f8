"""One GPU, every BASELINE.json configuration that fits it: device time per registration (CUDA events inside the library,
warm, median of 5) next to the CPU oracle on the host cores, with the parity of the two answers.

    python profiles/configs_table.py
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as graft
import workloads as W

capi = graft.load_package().capi
orc = graft.load_oracle()
try:
    orc.set_num_threads(len(os.sched_getaffinity(0)))
except AttributeError:
    pass
ctx = capi.Context(0)
RRM = dict(estimated_overlap=1.0, max_num_se3_iterations=10, mse=1e-5, mse_switch_error=5e-5, number_of_nn_for_LRF=90)
ENT = {"RUN_SE3_ICP": (capi.RUN_SE3_ICP, orc.RUN_SE3_ICP), "RUN_SE3_ICP_CF": (capi.RUN_SE3_ICP_CF, orc.RUN_SE3_ICP_CF)}


def row(name, src, tgt, entry, variant, params, cpu=True):
    pg = capi.default_params(variant=variant, entry=ENT[entry][0], reuse_features=0, **params)
    ctx.set_cloud(capi.SOURCE, src)
    ctx.set_cloud(capi.TARGET, tgt)
    times = []
    for _ in range(6):
        T, st = ctx.run(pg)
        times.append(st.time_total_ms)
    ms = float(np.median(times[1:]))
    cpu_txt, par = "-", "-"
    if cpu:
        t0 = time.perf_counter()
        To, so, _ = orc.run(src, tgt, orc.default_params(variant=variant, entry=ENT[entry][1], **params))
        cpu_ms = (time.perf_counter() - t0) * 1e3
        cpu_txt = "%.0f" % cpu_ms
        par = "%.1e rad, it %d/%d" % (W.rotation_error(T, To), st.num_iterations, so.num_iterations)
    print("| %s | %d / %d | %d (%d) | %.2f | %.2f | %s | %s |" % (name, len(src), len(tgt), st.num_iterations,
          st.num_pure_se3_iterations, ms, st.time_setup_ms, cpu_txt, par), flush=True)


print("| configuration | points | iterations (SE(3)) | GPU ms | of which set-up | CPU oracle ms (%d threads) | GPU vs oracle |" % orc.num_threads())
print("|---|---|---:|---:|---:|---:|---|")
s, t, _ = W.load_c1()
row("[0] fixture se3_pt2pl", s, t, "RUN_SE3_ICP", "pt2pl", RRM)
for level, seed in (("easy", 1), ("moderate", 2), ("difficult", 2)):
    s, t, _ = W.bunny_problem(level, seed=seed)
    for v in ("pt2pt", "pt2pl", "gicp"):
        row("[1] bunny %s se3_%s" % (level, v), s, t, "RUN_SE3_ICP", v, RRM)
s, t, _ = W.lidar_pair(seed=0)
row("[2] KITTI-like se3_gicp", s, t, "RUN_SE3_ICP", "gicp", W.KITTI_PARAMS)
s, t, _ = W.rgbd_pair(seed=0)
row("[3] lounge-like se3_gicp_with_cf", s, t, "RUN_SE3_ICP_CF", "gicp", W.LOUNGE_PARAMS)
