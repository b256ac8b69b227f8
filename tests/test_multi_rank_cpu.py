"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: sharding bookkeeping, max-over-ranks
timing, result gathering, and the algebra of the sharded pair (per-rank partial normal equations,
all-reduce, identical solve on every rank) emulated with the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as graft
    import workloads as W
    pkg = graft.load_package()
    orc = graft.load_oracle()
    sh = pkg.sharding
    res = {}
    # 1. batch mode bookkeeping
    owned = sh.pairs_of_rank(7, world, rank)
    T_local = np.stack([np.eye(4) * (p + 1) for p in owned])
    res["gathered"] = sh.gather_results(T_local, owned, 7, dist)
    res["tmax"] = sh.max_over_ranks(10.0 + rank, dist)
    cid = sh.broadcast_bytes(bytes(range(128)) if rank == 0 else b"", 128, dist, 0)
    res["cid_ok"] = cid == bytes(range(128))
    # 2. sharded pair: partial point-to-plane systems over the rank's source range, all-reduced
    src, tgt, T_gt = W.load_c1()
    rng = np.random.default_rng(0)
    moved = W.apply_T(T_gt, src) + rng.normal(0, 0.01, src.shape)
    corr, _ = orc.nn(moved, tgt)
    nrm = orc.normals(tgt, 30)
    b, e = sh.shard_range(len(src), world, rank)
    part = orc.reduce_pt2pl(moved, tgt, nrm, np.arange(b, e, dtype=np.int32), corr[b:e])
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)
    res["sum27"] = t.numpy()
    res["T"] = orc.solve6(t.numpy())
    res["whole"] = orc.reduce_pt2pl(moved, tgt, nrm, np.arange(len(src), dtype=np.int32), corr)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), res, allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions(pkg):
    sh = pkg.sharding
    for n in (0, 1, 7, 10_000_000):
        for world in (1, 2, 3, 8):
            r = [sh.shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    assert sorted(sum((sh.pairs_of_rank(256, 8, k) for k in range(8)), [])) == list(range(256))


def test_dealt_order_balances_contiguous_ranges(pkg):
    """dealt_order is a permutation, and after it every rank's contiguous range is a regular sample of the whole cloud"""
    sh = pkg.sharding
    for n, world, block in ((0, 2, 4), (10, 4, 4), (20_000, 3, 4096), (1_000_003, 8, 4096)):
        p = sh.dealt_order(n, world, block)
        assert p.shape == (n,) and np.array_equal(np.sort(p), np.arange(n))
    np.testing.assert_array_equal(sh.dealt_order(37, 1), np.arange(37))
    n, world, block = 1_000_003, 8, 4096
    p = sh.dealt_order(n, world, block)
    for r in range(world):
        b, e = sh.shard_range(n, world, r)
        share = p[b:e]
        # the original positions of a rank's share spread over the whole cloud: every eighth of the cloud holds ~1/8 of them
        hist = np.histogram(share, bins=8, range=(0, n))[0]
        assert hist.min() > 0.7 * len(share) / 8 and hist.max() < 1.3 * len(share) / 8  # one 4096-point block is a quarter of a bin
        # and they come in runs of consecutive points (memory order inside a block is kept)
        assert np.mean(np.diff(share) == 1) > 0.99
    # default block size: still a permutation with balanced shares
    p = sh.dealt_order(10_000_000, 8)
    assert np.array_equal(np.sort(p), np.arange(10_000_000))


def test_world2_gloo(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / ("r%d.npy" % k), allow_pickle=True).item() for k in range(world)]
    for k in range(world):
        assert r[k]["cid_ok"] and r[k]["tmax"] == 11.0
        for p in range(7):
            np.testing.assert_array_equal(r[k]["gathered"][p], np.eye(4) * (p + 1))
        np.testing.assert_allclose(r[k]["sum27"], r[k]["whole"], rtol=1e-12, atol=1e-12)
    np.testing.assert_array_equal(r[0]["sum27"], r[1]["sum27"])  # all-reduce: identical bits on every rank
    np.testing.assert_array_equal(r[0]["T"], r[1]["T"])          # hence the identical solve, no broadcast needed


def test_reference_arm_under_torchrun_world2():
    """bench.py --impl reference launched like the driver launches it for N > 1: rank 0 alone times the CPU path and prints
    ONE JSON line (impl, cpu_baseline, e2e with zero copy bytes); the other rank exits 0 without output."""
    import json
    import subprocess

    port = 31500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0"]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "registrations/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert abs(d["e2e"]["value"] - d["value"]) < 1e-9 and "BASELINE.json configs[2]" in d["config"]["workload"]
