"""Does the batch bench slow down at 8 GPUs because the ranks get different scenes?  Times, on ONE GPU, the 32-pair step
of bench.py for the scene set each rank of an 8-GPU run would generate (seeds rank * 8 + i).

    python profiles/rank_workloads.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import workloads as W  # noqa: E402

capi = graft.load_package().capi
params = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, **W.KITTI_PARAMS)
ctxs = [capi.Context(0) for _ in range(16)]
for rank in range(8):
    pairs = [W.lidar_pair(seed=rank * 8 + i) for i in range(8)]
    dev = [(torch.from_numpy(np.ascontiguousarray(s)).cuda(), torch.from_numpy(np.ascontiguousarray(t)).cuda()) for s, t, _ in pairs]
    lst = [(dev[i % 8][0].data_ptr(), dev[i % 8][0].shape[0], dev[i % 8][1].data_ptr(), dev[i % 8][1].shape[0]) for i in range(32)]
    for _ in range(2):
        capi.run_batch(ctxs, lst, params, device_inputs=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        T, st = capi.run_batch(ctxs, lst, params, device_inputs=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print("scene set of rank %d: %.1f ms per 32-pair step (%.1f registrations/s), iterations per pair %s" %
          (rank, 1e3 * dt, 32 / dt, [s.num_iterations for s in st[:8]]), flush=True)
