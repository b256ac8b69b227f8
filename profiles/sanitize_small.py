"""Small end-to-end exercise of every entry point for compute-sanitizer:
    compute-sanitizer --tool memcheck python profiles/sanitize_small.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import workloads as W  # noqa: E402

capi = graft.load_package().capi
src, tgt, T_gt = W.load_c1()
src, tgt = src[:1500], tgt[:1300]
ctx = capi.Context(0)
kw = dict(max_num_se3_iterations=4, max_num_iterations=8, mse=1e-5, mse_switch_error=5e-5, number_of_nn_for_LRF=90)
n = 0
for entry in (capi.RUN_ICP, capi.RUN_SE3_ICP, capi.RUN_SE3_ICP_CF, capi.RUN_SE3_PURE):
    for variant in ("pt2pt", "pt2pl", "gicp"):
        for overlap in (1.0, 0.7):
            for graph in (0, 1):
                ctx.set_cloud(capi.SOURCE, src)
                ctx.set_cloud(capi.TARGET, tgt)
                T, st = ctx.run(capi.default_params(variant=variant, entry=entry, estimated_overlap=overlap, use_graph=graph, **kw))
                n += 1
for mode in (capi.NN_BRUTE_F32, capi.NN_EXACT_F64):
    ctx.run(capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, nn_mode=mode, **kw))
ctx.knn(src, 90)
ctx.lrf(src, 90)
ctx.normals(src, 30)
rng = np.random.default_rng(0)
rows_s, rows_t = rng.normal(size=(300, 12)), rng.normal(size=(257, 12))
for mode in (capi.NN_TREE, capi.NN_BRUTE_F32, capi.NN_EXACT_F64):
    ctx.nn_se3(rows_s, rows_t, mode)
ctx.nn_xyz(src, tgt)
ctx.trim(rng.random(1000).astype(np.float32), 0.7)
ctx.comm_init(1, 0, capi.comm_unique_id())
ctx.set_cloud(capi.SOURCE, src)
ctx.set_cloud(capi.TARGET, tgt)
ctx.run_sharded(capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, estimated_overlap=0.7, **kw), 0, len(src))
ctx.close()
print("sanitize_small: %d registrations + stage calls done" % n)
