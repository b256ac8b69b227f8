"""Batch throughput vs number of concurrent contexts (streams) on one GPU: python profiles/ctx_scaling.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as graft
import workloads as W
capi = graft.load_package().capi
pairs = [W.lidar_pair(seed=i) for i in range(4)]
dev = [(torch.from_numpy(s).cuda(), torch.from_numpy(t).cuda()) for s, t, _ in pairs]
P = 32
lst = [(dev[i % 4][0].data_ptr(), dev[i % 4][0].shape[0], dev[i % 4][1].data_ptr(), dev[i % 4][1].shape[0]) for i in range(P)]
for graph in (1,):
    params = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, use_graph=graph, **W.KITTI_PARAMS)
    for nctx in (1, 2, 4, 8, 12, 16, 24, 32):
        ctxs = [capi.Context(0) for _ in range(nctx)]
        capi.run_batch(ctxs, lst, params, device_inputs=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            capi.run_batch(ctxs, lst, params, device_inputs=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print("use_graph=%d contexts=%d: %.2f ms/pair, %.1f reg/s" % (graph, nctx, 1e3 * dt / P, P / dt))
        for c in ctxs:
            c.close()
