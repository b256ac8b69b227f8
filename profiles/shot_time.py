"""SHOT frames (se3icp_shot_lrf, reference .cpp:121-239) of the full bunny (34 834 points, normalised to radius 3) at the
class default radius 0.8 and at 0.3: GPU wall time through the C ABI (host buffers in and out) against the CPU oracle.
    python profiles/shot_time.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
import workloads as W
capi = graft.load_package().capi
orc = graft.load_oracle()
pts = W.load_bunny().astype(np.float64)
pts = (pts - pts.mean(0)) * (3.0 / np.linalg.norm(pts - pts.mean(0), axis=1).max())
ctx = capi.Context(0)
for r in (0.3, 0.8):
    ctx.shot_lrf(pts, r)
    t0 = time.perf_counter(); g, unresolved = ctx.shot_lrf(pts, r, return_unresolved=True); tg = time.perf_counter() - t0
    t0 = time.perf_counter(); o = orc.shot(pts, r); tc = time.perf_counter() - t0
    d2 = ((pts[:, None, :] - pts[None, ::97, :]) ** 2).sum(2)
    print("radius %.1f: support %d points on average; GPU %.1f ms (incl. upload / index / download), oracle %.0f ms on %d threads; "
          "max |difference| %.1e, unresolved ties %d" % (r, (d2 < r * r).sum(0).mean(), 1e3 * tg, 1e3 * tc, orc.num_threads(), np.abs(g - o).max(), unresolved))
