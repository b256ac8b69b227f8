"""Duration of the solve / update tail in isolation (stage-level calls; run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as graft
import workloads as W
capi = graft.load_package().capi
src, tgt, T = W.load_c1()
ctx = capi.Context(0)
rng = np.random.default_rng(0)
A = rng.normal(size=(6, 6)); A = A @ A.T + np.eye(6)
in27 = np.concatenate([A[np.triu_indices(6)], rng.normal(size=6)])
for _ in range(3):
    ctx.solve(in27)                                     # pt2pl / gicp tail: LDLT + Euler update
    ctx.reduce_pt2pt(src, W.apply_T(T, src), np.arange(len(src), dtype=np.int32))  # pt2pt tail: Kabsch
print("done")
