"""BASELINE.json configs[4] at its named size: ONE pair of ~10 M points, se3_pt2pl, target replicated, source queries
sharded over the ranks, one 29-double NCCL all-reduce per iteration.

    python profiles/big_pair.py [scale]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        profiles/big_pair.py [scale]                                      # sharded; also runs the whole pair on rank 0's GPU

scale 6 -> 3840 x 2880 depth image (~10 M valid points per frame)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as graft
import workloads as W

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 6
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
pkg = graft.load_package()
capi, sh = pkg.capi, pkg.sharding
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.time()
src, tgt, T_gt = W.rgbd_pair(seed=0, width=640 * scale, height=480 * scale, f=525.0 * scale, stride=1)
gen_s = time.time() - t0
ctx = capi.Context(local)
p = capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, estimated_overlap=1.0, max_num_se3_iterations=10, mse=1e-5,
                        mse_switch_error=5e-5, number_of_nn_for_LRF=90, reuse_features=0)  # every run is a full run
ctx.set_cloud(capi.SOURCE, src)
ctx.set_cloud(capi.TARGET, tgt)
lines = []
if world > 1:
    sh.init_sharded_comm(ctx, capi, dist, torch.device("cuda", local))
    b, e = sh.shard_range(len(src), world, rank)
    for _ in range(2):
        Ts, ss = ctx.run_sharded(p, b, e)
    lines.append("sharded over %d GPUs: %.1f ms (set-up %.1f ms), %d iterations (%d SE(3))" %
                 (world, ss.time_total_ms, ss.time_setup_ms, ss.num_iterations, ss.num_pure_se3_iterations))
    dist.barrier()
if rank == 0:
    for _ in range(2):
        T1, s1 = ctx.run(p)
    lines.append("one GPU: %.1f ms (set-up %.1f ms), %d iterations (%d SE(3)); vs GT rot %.2e rad transl %.4f m" %
                 (s1.time_total_ms, s1.time_setup_ms, s1.num_iterations, s1.num_pure_se3_iterations,
                  W.rotation_error(T1, T_gt), float(np.linalg.norm(T1[:3, 3] - T_gt[:3, 3]))))
    if world > 1:
        lines.append("sharded vs one GPU: rot %.1e rad, transl %.1e, iterations %d/%d" %
                     (W.rotation_error(Ts, T1), float(np.linalg.norm(Ts[:3, 3] - T1[:3, 3])), ss.num_iterations, s1.num_iterations))
    print("BIG PAIR %d / %d points (generated in %.0f s), %.1f GB allocated on the GPU\n  " %
          (len(src), len(tgt), gen_s, torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9) + "\n  ".join(lines))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
