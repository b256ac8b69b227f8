"""Quick A/B figure: one pair alone (3 runs, last reported) and batch throughput at 8 contexts (32 pairs x 3, 4 scenes).
    python profiles/experiments/quick_tput.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as graft
import workloads as W
capi = graft.load_package().capi
pairs = [W.lidar_pair(seed=i) for i in range(4)]
params = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, reuse_features=0, **W.KITTI_PARAMS)
ctx = capi.Context(0)
ctx.set_cloud(capi.SOURCE, pairs[0][0]); ctx.set_cloud(capi.TARGET, pairs[0][1])
for _ in range(4):
    T, st = ctx.run(params)
print("pair %.2f ms (setup %.2f, %d it)" % (st.time_total_ms, st.time_setup_ms, st.num_iterations), end="; ")
ctx.close()
dev = [(torch.from_numpy(s).cuda(), torch.from_numpy(t).cuda()) for s, t, _ in pairs]
P = 32
lst = [(dev[i % 4][0].data_ptr(), dev[i % 4][0].shape[0], dev[i % 4][1].data_ptr(), dev[i % 4][1].shape[0]) for i in range(P)]
ctxs = [capi.Context(0) for _ in range(8)]
capi.run_batch(ctxs, lst, params, device_inputs=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    capi.run_batch(ctxs, lst, params, device_inputs=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 4
print("batch %.1f reg/s" % (P / dt))
