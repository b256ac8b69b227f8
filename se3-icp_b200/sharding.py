"""Host-side sharding logic for the two multi-GPU modes.

  * batch of independent pairs: pair p belongs to rank p % world (no data-path collective)
  * one very large pair: contiguous source query ranges per rank, target replicated, one all-reduce of the
    normal-equation record per iteration inside libse3icp_cuda.so (NCCL)

Only bookkeeping lives here; it works with any torch.distributed backend (gloo in the CPU tests).
"""
import numpy as np


def shard_range(n, world, rank):
    """contiguous [begin, end) of `n` items for `rank`; sizes differ by at most one"""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def dealt_order(n, world, block=32768):
    """Permutation that deals `n` points to `world` ranks in blocks of `block` consecutive points and lays every rank's
    share out contiguously: after `cloud[dealt_order(n, world)]` the contiguous range shard_range(n, world, r) of rank r
    holds the blocks r, r + world, r + 2 world, ... of the original order.  A scan is ordered by image row or by laser
    ring, so plain contiguous ranges give the ranks different parts of the scene and different search costs; dealing
    the blocks out balances them while se3icp_run_sharded still gets one contiguous range per rank.  (A point cloud is
    an unordered set: the registration result does not depend on the order beyond summation order.)
    Measured on the 10.1 M-point pair over 8 GPUs (profiles/sharded_blocks.py): contiguous ranges 257 ms (search time
    per rank 152-185 ms), blocks of 4096 / 32768 / 262144 points 239 / 236 / 241 ms (larger blocks keep more of the
    scan order's locality, smaller ones balance better)."""
    n, world, block = int(n), int(world), int(block)
    if world <= 1:
        return np.arange(n, dtype=np.int64)
    blocks = np.arange((n + block - 1) // block, dtype=np.int64)
    dealt = np.concatenate([blocks[r::world] for r in range(world)])
    idx = (dealt[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).reshape(-1)
    return idx[idx < n]


def pairs_of_rank(n_pairs, world, rank):
    """indices of the independent pairs owned by `rank` (round robin, as SURVEY §8e)"""
    return list(range(rank, n_pairs, world))


def max_over_ranks(value, dist=None, device=None):
    """multi-GPU timings are the max over ranks"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_bytes(payload, n_bytes, dist, src=0, device=None):
    """broadcast a small byte string (the NCCL unique id) from `src` over torch.distributed"""
    import torch
    if dist.get_rank() == src:
        t = torch.tensor(list(payload), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(n_bytes, dtype=torch.uint8, device=device)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tolist())


def init_sharded_comm(ctx, capi, dist, device=None):
    """creates the library-owned NCCL communicator of `ctx` across all torch.distributed ranks"""
    rank, world = dist.get_rank(), dist.get_world_size()
    cid = capi.comm_unique_id() if rank == 0 else b""
    cid = broadcast_bytes(cid, capi.COMM_ID_BYTES, dist, 0, device)
    ctx.comm_init(world, rank, cid)
    return rank, world


def gather_results(T_local, owned, n_pairs, dist=None, device=None):
    """assemble per-pair transforms from all ranks (sum of disjoint contributions)"""
    out = np.zeros((n_pairs, 4, 4))
    out[owned] = T_local
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return out
    import torch
    t = torch.from_numpy(out).to(device) if device is not None else torch.from_numpy(out)
    dist.all_reduce(t)
    return t.cpu().numpy()
