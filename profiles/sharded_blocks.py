"""Sharded 10 M-point pair: effect of the block size with which the source is dealt to the ranks (balance vs locality).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 profiles/sharded_blocks.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.distributed as dist
import __graft_entry__ as graft
import workloads as W
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
pkg = graft.load_package(); capi, sh = pkg.capi, pkg.sharding
dist.init_process_group("nccl", device_id=dev)
src0, tgt, _ = W.rgbd_pair_device(seed=0, device=dev)
p = capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, reuse_features=0, estimated_overlap=1.0, max_num_se3_iterations=10,
                        mse=1e-5, mse_switch_error=5e-5, number_of_nn_for_LRF=90)
ctx = capi.Context(local)
sh.init_sharded_comm(ctx, capi, dist, dev)
ctx.set_cloud_device(capi.TARGET, tgt.data_ptr(), tgt.shape[0])
b, e = sh.shard_range(src0.shape[0], world, rank)
for block in (0, 4096, 32768, 262144):
    src = src0 if block == 0 else src0[torch.from_numpy(sh.dealt_order(src0.shape[0], world, block)).to(dev)].contiguous()
    ctx.set_cloud_device(capi.SOURCE, src.data_ptr(), src.shape[0])
    best = None
    for k in range(3):
        dist.barrier(); torch.cuda.synchronize()
        T, s = ctx.run_sharded(p, b, e)
        if k and (best is None or s.time_total_ms < best.time_total_ms):
            best = s
    v = torch.tensor([best.time_total_ms, best.time_setup_ms, best.time_se3_correspondence_search_ms], dtype=torch.float64, device=dev)
    lo, hi = v.clone(), v.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("block %7d: total %.1f ms (max over ranks), set-up %.1f-%.1f, search %.1f-%.1f, iterations %d" %
              (block, hi[0], lo[1], hi[1], lo[2], hi[2], best.num_iterations), flush=True)
    del src
ctx.close()
dist.barrier(); dist.destroy_process_group()
