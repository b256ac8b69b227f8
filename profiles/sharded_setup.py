"""Set-up stage times of the sharded 10 M-point pair (SE3ICP_SETUP_TIMING=1), one GPU or torchrun.
    SE3ICP_SETUP_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 profiles/sharded_setup.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as graft
import workloads as W
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
pkg = graft.load_package(); capi, sh = pkg.capi, pkg.sharding
import torch.distributed as dist
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
src, tgt, _ = W.rgbd_pair_device(seed=0, device=dev)
if os.environ.get("DEALT", "1") == "1":
    src = src[torch.from_numpy(sh.dealt_order(src.shape[0], world)).to(dev)].contiguous()
p = capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, reuse_features=0, estimated_overlap=1.0, max_num_se3_iterations=10,
                        mse=1e-5, mse_switch_error=5e-5, number_of_nn_for_LRF=90)
ctx = capi.Context(local)
ctx.set_cloud_device(capi.SOURCE, src.data_ptr(), src.shape[0])
ctx.set_cloud_device(capi.TARGET, tgt.data_ptr(), tgt.shape[0])
if world > 1:
    sh.init_sharded_comm(ctx, capi, dist, dev)
    b, e = sh.shard_range(src.shape[0], world, rank)
    for k in range(3):
        dist.barrier(); torch.cuda.synchronize()
        if rank == 0: print("--- sharded run %d" % k, file=sys.stderr, flush=True)
        T, s = ctx.run_sharded(p, b, e)
    if rank == 0: print("sharded: total %.1f setup %.1f search %.1f" % (s.time_total_ms, s.time_setup_ms, s.time_se3_correspondence_search_ms))
else:
    for k in range(2):
        print("--- run %d" % k, file=sys.stderr, flush=True)
        T, s = ctx.run(p)
    print("one GPU: total %.1f setup %.1f search %.1f" % (s.time_total_ms, s.time_setup_ms, s.time_se3_correspondence_search_ms))
ctx.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
