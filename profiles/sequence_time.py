"""Sequence driver vs independent pairs on KITTI-size scans (one context): python profiles/sequence_time.py [n_scans]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as graft
import workloads as W
capi = graft.load_package().capi
n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 6
scans, steps = W.lidar_sequence(seed=0, n_scans=n_scans)
dev = [torch.from_numpy(np.ascontiguousarray(s)).cuda() for s in scans]
ptrs = [(d.data_ptr(), d.shape[0]) for d in dev]
ctx = capi.Context(0)
for reuse in (1, 0, 1, 0):
    p = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, reuse_features=reuse, **W.KITTI_PARAMS)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    T, st = ctx.run_sequence(ptrs, p, device_inputs=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    err = max(W.rotation_error(T[i], steps[i]) for i in range(n_scans - 1))
    print("reuse=%d: %d registrations in %.1f ms (%.2f ms each, set-up %.2f ms avg, reuses %s), max rot err vs GT %.2e rad"
          % (reuse, n_scans - 1, dt * 1e3, dt * 1e3 / (n_scans - 1), np.mean([s.time_setup_ms for s in st]),
             [s.feature_reuses for s in st], err))
