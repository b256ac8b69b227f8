"""GPU parity tests (run with -m gpu on a B200): every stage of the CUDA path through the C ABI
against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): correspondence / neighbour indices bit-exact except equal-distance
ties within 1e-6 relative; LRFs and normals within 1e-4 (normals up to sign); final transforms within
1e-5 rad and 1e-5 x cloud extent."""
import os

import numpy as np
import pytest

import workloads as W
from conftest import rot_err

pytestmark = pytest.mark.gpu

RRM = dict(estimated_overlap=1.0, max_num_se3_iterations=10, mse=1e-5, mse_switch_error=5e-5,
           number_of_nn_for_LRF=90)  # examples/run_registration_method.cpp:38-42


def assert_indices_match(idx_gpu, d_gpu, idx_ref, d_ref, what):
    """bit-exact indices, except where the two picks are at equal distance within 1e-6 relative"""
    diff = idx_gpu != idx_ref
    tol = 1e-6 * np.maximum(np.abs(d_ref), 1e-300)
    assert np.all(np.abs(d_gpu - d_ref)[diff] <= tol[diff]), "%s: %d index mismatches beyond ties" % (what, diff.sum())
    return int(diff.sum())


def normalised(src, tgt):
    cs, ct = src.mean(0), tgt.mean(0)
    r = max(np.linalg.norm(src - cs, axis=1).max(), np.linalg.norm(tgt - ct, axis=1).max())
    s = 3.0 / r
    return (src - cs) * s, (tgt - ct) * s, s


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 20, 30, 90, 128])
def test_knn_bit_exact(ctx, orc, c1, k):
    src, _, _ = c1
    gi, gd = ctx.knn(src, k)
    oi, od = orc.knn_self(src, k)
    np.testing.assert_array_equal(gd, od)  # same non-contracted FP64 arithmetic -> identical bits
    np.testing.assert_array_equal(gi, oi)  # ties (196 exact duplicates) resolve to the smaller index on both sides


def test_knn_bunny_full(ctx, orc):
    pts = W.load_bunny()
    gi, gd = ctx.knn(pts, 90)
    oi, od = orc.knn_self(pts, 90)
    np.testing.assert_array_equal(gd, od)
    np.testing.assert_array_equal(gi, oi)


@pytest.mark.parametrize("n_dup,k", [(230, 90), (600, 90), (300, 128), (260, 20)])
def test_knn_many_coincident_points(ctx, orc, n_dup, k):
    """Hundreds of coincident points (invalid-depth pixels mapped to one xyz): every candidate ties at the search
    radius, so the bisection on the distance value cannot shrink the per-warp pool; the exact (distance, index) trim
    must keep it bounded and the result must still be the oracle's list (ties -> smaller index)."""
    rng = np.random.default_rng(n_dup)
    base = rng.normal(size=(700, 3))
    dup = np.repeat(np.array([[0.25, -0.5, 0.125]]), n_dup, axis=0)
    ring = np.array([0.25, -0.5, 0.125]) + 0.5 * np.eye(3)[rng.integers(0, 3, 200)] * rng.choice([-1.0, 1.0], (200, 1))
    pts = np.concatenate([base, dup, ring])[rng.permutation(700 + n_dup + 200)]
    gi, gd = ctx.knn(pts, k)
    oi, od = orc.knn_self(pts, k)
    np.testing.assert_array_equal(gd, od)
    np.testing.assert_array_equal(gi, oi)
    # the feature pass over such neighbourhoods must run through as well; away from the coincident cluster (whose
    # scatter matrix is exactly zero, so its eigenvectors are arbitrary) the frames agree with the oracle
    gf = ctx.lrf(pts, k)[:, :3, :3]
    of = orc.toldi(pts, k)[:, :3, :3]
    far = np.linalg.norm(pts - np.array([0.25, -0.5, 0.125]), axis=1) > 1.5
    ok = far & np.isfinite(of).all(axis=(1, 2))
    assert ok.sum() > 50
    np.testing.assert_allclose(gf[ok], of[ok], atol=1e-9)


def test_knn_edge_cases(ctx, orc, capi):
    rng = np.random.default_rng(0)
    for n in (1, 2, 31, 32, 33, 100):
        pts = rng.normal(size=(n, 3))
        gi, gd = ctx.knn(pts, 40)
        oi, od = orc.knn_self(pts, 40)
        np.testing.assert_array_equal(gi, oi)  # fewer than k points -> -1 padding on both sides
        np.testing.assert_array_equal(gd, od)
    with pytest.raises(capi.Se3IcpError) as e:
        ctx.knn(rng.normal(size=(300, 3)), 129)
    assert e.value.code == 5  # SE3ICP_ERR_UNSUPPORTED, no silent truncation
    # clustered + widely spread data (non-uniform density like a LiDAR scan)
    pts = np.concatenate([rng.normal(size=(3000, 3)) * 0.01, rng.normal(size=(3000, 3)) * 50.0])
    gi, gd = ctx.knn(pts, 90)
    oi, od = orc.knn_self(pts, 90)
    np.testing.assert_array_equal(gi, oi)


def test_lrf_matches_oracle(ctx, orc, c1):
    src, tgt, _ = c1
    ns, nt, _ = normalised(src, tgt)
    for cloud in (ns, nt):
        g = ctx.lrf(cloud, 90)
        o = orc.toldi(cloud, 90)
        np.testing.assert_allclose(g, o, atol=1e-4)  # the stated bar
        assert np.abs(g - o).max() < 1e-9            # what FP64 on both sides actually gives
    g = ctx.lrf(ns, 30)
    np.testing.assert_allclose(g, orc.toldi(ns, 30), atol=1e-9)


def test_lrf_golden(ctx, c1):
    import os
    gold = np.load(os.path.join(W.GOLDEN, "c1_se3_pt2pl_trace.npz"))
    src, tgt, _ = c1
    ns, nt, _ = normalised(src, tgt)
    np.testing.assert_allclose(ctx.lrf(ns, 90)[:, :3, :3], gold["frames_src"], atol=1e-4)
    np.testing.assert_allclose(ctx.lrf(nt, 90)[:, :3, :3], gold["frames_tgt"], atol=1e-4)


@pytest.mark.parametrize("k", [20, 30])
def test_normals_and_cov(ctx, orc, c1, k):
    src, _, _ = c1
    g = ctx.normals(src, k)
    o = orc.normals(src, k)
    sign = np.sign((g * o).sum(1))
    assert np.all(sign != 0)
    np.testing.assert_allclose(g * sign[:, None], o, atol=1e-4)
    np.testing.assert_allclose(np.linalg.norm(g, axis=1), 1.0, atol=1e-12)
    # covariance from the SAME normals must agree to rounding, including the c < -0.99 branch
    nr = np.concatenate([o, [[-1.0, 0, 0], [-0.995, 0.0998749, 0.0], [1.0, 0, 0]]])
    np.testing.assert_allclose(ctx.gicp_cov(nr, 1e-3), orc.gicp_cov(nr, 1e-3), atol=1e-12)


# ---------------------------------------------------------------------------------------------------
def se3_rows_pair(orc, c1, T_apply=None):
    src, tgt, _ = c1
    ns, nt, _ = normalised(src, tgt)
    rs = orc.se3_rows(orc.toldi(ns, 90), 3.0, 1.0)
    rt = orc.se3_rows(orc.toldi(nt, 90), 3.0, 1.0)
    return rs, rt


@pytest.mark.parametrize("mode_name", ["NN_TREE", "NN_BRUTE_F32", "NN_EXACT_F64"])
def test_nn_se3_matches_oracle(ctx, orc, capi, c1, mode_name):
    rs, rt = se3_rows_pair(orc, c1)
    gi, gd, rep = ctx.nn_se3(rs, rt, getattr(capi, mode_name))
    oi, od = orc.nn(rs, rt, brute=True)
    np.testing.assert_array_equal(gd, od)  # exact FP64 distance of the winner
    np.testing.assert_array_equal(gi, oi)  # smallest-index tie-break on both sides
    ti, td = orc.nn(rs, rt)                # kd-tree oracle (the reference's structure)
    assert_indices_match(gi, gd, ti, td, "nn_se3 vs kd-tree")
    if mode_name == "NN_BRUTE_F32":
        assert rep < 0.2 * len(rs)         # duplicates (5 %) are real ties and must go through the exact repair
        assert rep >= 50


def test_nn_se3_aligned_regime(ctx, orc, capi, c1):
    """near convergence the best and second-best are close neighbours: the certification must still hold"""
    src, tgt, T_gt = c1
    ns, nt, s = normalised(src, tgt)
    fs, ft = orc.toldi(ns, 90), orc.toldi(nt, 90)
    # move the source frames by the (normalised-space) ground truth plus a small perturbation
    Tn = np.eye(4)
    Tn[:3, :3] = T_gt[:3, :3] @ W.rot_3d(1e-3, -2e-3, 1.5e-3)
    Tn[:3, 3] = [1e-3, -2e-3, 5e-4]
    rs = orc.se3_rows(np.einsum("ij,njk->nik", Tn, fs), 3.0, 1.0)
    rt = orc.se3_rows(ft, 3.0, 1.0)
    oi, od = orc.nn(rs, rt, brute=True)
    for mode in (capi.NN_TREE, capi.NN_BRUTE_F32):
        gi, gd, rep = ctx.nn_se3(rs, rt, mode)
        np.testing.assert_array_equal(gi, oi)
        np.testing.assert_array_equal(gd, od)


def test_nn_se3_random_and_ragged(ctx, orc, capi):
    rng = np.random.default_rng(7)
    for n, m in ((1, 1), (5, 3), (257, 255), (1000, 513)):
        rs, rt = rng.normal(size=(n, 12)) * 2, rng.normal(size=(m, 12)) * 2
        if m > 2:
            rt[2] = rt[0]
        for mode in (capi.NN_TREE, capi.NN_BRUTE_F32, capi.NN_EXACT_F64):
            gi, gd, _ = ctx.nn_se3(rs, rt, mode)
            oi, od = orc.nn(rs, rt, brute=True)
            np.testing.assert_array_equal(gi, oi)
            np.testing.assert_array_equal(gd, od)


def test_nn_xyz_matches_oracle(ctx, orc, c1):
    src, tgt, T_gt = c1
    for q in (src, W.apply_T(T_gt, src) + 1e-3, tgt):
        gi, gd = ctx.nn_xyz(q, tgt)
        oi, od = orc.nn(q, tgt, brute=True)
        np.testing.assert_array_equal(gd, od)
        np.testing.assert_array_equal(gi, oi)
    rng = np.random.default_rng(1)
    for n, m in ((1, 1), (10, 33), (500, 5000)):
        q, d = rng.normal(size=(n, 3)), rng.normal(size=(m, 3)) * [5, 1, 0.1]
        gi, gd = ctx.nn_xyz(q, d)
        oi, od = orc.nn(q, d, brute=True)
        np.testing.assert_array_equal(gi, oi)
        np.testing.assert_array_equal(gd, od)


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("overlap", [1.0, 0.99, 0.8, 0.75, 0.7, 0.5, 0.01, 0.0001])
@pytest.mark.parametrize("keep_largest", [False, True])
def test_trim_matches_oracle(ctx, orc, overlap, keep_largest):
    rng = np.random.default_rng(11)
    for n in (1, 7, 1000, 4167, 100000):
        d = rng.gamma(2.0, 0.01, n).astype(np.float32)  # distinct values -> the kept set is unique
        gk, gkeep = ctx.trim(d, overlap, keep_largest)
        ok, okeep = orc.trim(d, overlap, keep_largest)
        assert gk == ok == int(gkeep.sum())
        if len(np.unique(d)) == n:
            np.testing.assert_array_equal(gkeep, okeep)
        else:  # ties at the threshold: same multiset of kept distances
            np.testing.assert_array_equal(np.sort(d[gkeep]), np.sort(d[okeep]))


def test_trim_multipass_kernels_agree():
    """The sharded pair selects the trimmed threshold with four all-reducible 8-bit histogram passes and a keep mask; one
    GPU uses the single-pass selection.  The same entry point runs the former with SE3ICP_TRIM_MULTIPASS=1 (read once
    per process, hence the subprocess): both must return the oracle's kept set."""
    import subprocess
    import sys
    code = """
import sys, numpy as np
sys.path.insert(0, %r)
import __graft_entry__ as g
capi, orc = g.load_package().capi, g.load_oracle()
rng = np.random.default_rng(5)
with capi.Context(0) as ctx:
    for n in (1, 7, 4167, 100000):
        d = rng.gamma(2.0, 0.01, n).astype(np.float32)
        d[::17] = d[0]  # ties
        for ov in (0.99, 0.7, 0.5, 0.01):
            for kl in (False, True):
                gk, gkeep = ctx.trim(d, ov, kl)
                ok, okeep = orc.trim(d, ov, kl)
                assert gk == ok == int(gkeep.sum()), (n, ov, kl)
                assert np.array_equal(np.sort(d[gkeep]), np.sort(d[okeep])), (n, ov, kl)
print("TRIM_MULTIPASS_OK")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, SE3ICP_TRIM_MULTIPASS="1"))
    assert "TRIM_MULTIPASS_OK" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


def test_trim_ties(ctx, orc):
    d = np.zeros(1000, np.float32)           # converged exact copy: every distance is 0
    d[::3] = 0.5
    for ov in (0.7, 0.5, 0.2):
        gk, gkeep = ctx.trim(d, ov)
        ok, okeep = orc.trim(d, ov)
        assert gk == ok == int(gkeep.sum())
        np.testing.assert_array_equal(np.sort(d[gkeep]), np.sort(d[okeep]))
        # deterministic: among equal distances the lowest indices survive
        eq = np.nonzero(d == d[gkeep].max())[0]
        kept_eq = np.nonzero(gkeep & (d == d[gkeep].max()))[0]
        np.testing.assert_array_equal(kept_eq, eq[:len(kept_eq)])


def test_reduce_and_solve(ctx, orc, c1):
    src, tgt, T_gt = c1
    rng = np.random.default_rng(5)
    moved = W.apply_T(T_gt, src) + rng.normal(0, 0.01, src.shape)
    corr, _ = orc.nn(moved, tgt)
    nrm = orc.normals(tgt, 30)
    cs = np.arange(len(src), dtype=np.int32)
    g = ctx.reduce_pt2pl(moved, tgt, nrm, corr)
    o = orc.reduce_pt2pl(moved, tgt, nrm, cs, corr)
    np.testing.assert_allclose(g, o, rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(ctx.solve(o), orc.solve6(o), atol=1e-12)
    Cs, Ct = orc.gicp_cov(orc.normals(moved, 20)), orc.gicp_cov(orc.normals(tgt, 20))
    g = ctx.reduce_gicp(moved, Cs, tgt, Ct, corr)
    o = orc.reduce_gicp(moved, Cs, tgt, Ct, cs, corr)
    np.testing.assert_allclose(g, o, rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(ctx.solve(g), orc.solve6(o), atol=1e-10)
    conf_s, conf_t = rng.uniform(0.1, 1, len(src)), rng.uniform(0.1, 1, len(tgt))
    g = ctx.reduce_gicp(moved, Cs, tgt, Ct, corr, conf_s, conf_t)
    o = orc.reduce_gicp(moved, Cs, tgt, Ct, cs, corr, (conf_s + conf_t[corr]) / 2)
    np.testing.assert_allclose(g, o, rtol=1e-9, atol=1e-8)
    # rejected correspondences (-1) are skipped
    corr2 = corr.copy()
    corr2[::2] = -1
    keep = corr2 >= 0
    g = ctx.reduce_pt2pl(moved, tgt, nrm, corr2)
    o = orc.reduce_pt2pl(moved, tgt, nrm, cs[keep], corr2[keep])
    np.testing.assert_allclose(g, o, rtol=1e-10, atol=1e-9)
    # Umeyama
    np.testing.assert_allclose(ctx.reduce_pt2pt(moved, tgt, corr), orc.umeyama(moved, tgt, cs, corr), atol=1e-10)
    np.testing.assert_allclose(ctx.reduce_pt2pt(src, tgt, cs), T_gt, atol=1e-10)


# ---------------------------------------------------------------------------------------------------
ENTRIES = [("RUN_SE3_ICP", "pt2pt"), ("RUN_SE3_ICP", "pt2pl"), ("RUN_SE3_ICP", "gicp"),
           ("RUN_ICP", "pt2pt"), ("RUN_ICP", "pt2pl"), ("RUN_ICP", "gicp"),
           ("RUN_SE3_PURE", "pt2pl"), ("RUN_SE3_PURE", "gicp"), ("RUN_SE3_ICP_CF", "gicp")]


def run_both(ctx, orc, capi, src, tgt, entry_name, variant, **kw):
    po = orc.default_params(variant=variant, entry=getattr(orc, entry_name), **kw)
    pg = capi.default_params(variant=variant, entry=getattr(capi, entry_name), **kw)
    To, so, _ = orc.run(src, tgt, po)
    ctx.set_cloud(capi.SOURCE, src)
    ctx.set_cloud(capi.TARGET, tgt)
    Tg, sg = ctx.run(pg)
    return Tg, sg, To, so


def assert_transform_parity(Tg, To, tgt):
    extent = np.linalg.norm(tgt.max(0) - tgt.min(0))
    assert rot_err(Tg, To) < 1e-5, "rotation differs by %.3e rad" % rot_err(Tg, To)
    assert np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < 1e-5 * extent
    np.testing.assert_array_equal(Tg[3], [0, 0, 0, 1])


@pytest.mark.parametrize("entry_name,variant", ENTRIES)
def test_registration_fixture(ctx, orc, capi, c1, entry_name, variant):
    """BASELINE.json configs[0] and its siblings: all six method names + with_cf + pure on the bundled fixture"""
    src, tgt, T_gt = c1
    Tg, sg, To, so = run_both(ctx, orc, capi, src, tgt, entry_name, variant, **RRM)
    assert_transform_parity(Tg, To, tgt)
    assert_transform_parity(Tg, T_gt, tgt)
    assert sg.num_iterations == so.num_iterations
    assert sg.num_pure_se3_iterations == so.num_pure_se3_iterations
    assert abs(sg.scaling_factor - so.scaling_factor) < 1e-12 * so.scaling_factor


@pytest.mark.parametrize("level,seed", [("easy", 1), ("moderate", 2), ("difficult", 2)])
@pytest.mark.parametrize("variant", ["pt2pt", "pt2pl", "gicp"])
def test_registration_bunny_noisy(ctx, orc, capi, level, seed, variant):
    """BASELINE.json configs[1]: noisy bunny, reference-faithful 2 % down-sample"""
    src, tgt, T_gt = W.bunny_problem(level, seed=seed, n_points=4167)
    Tg, sg, To, so = run_both(ctx, orc, capi, src, tgt, "RUN_SE3_ICP", variant, **RRM)
    assert_transform_parity(Tg, To, tgt)
    assert sg.num_iterations == so.num_iterations and sg.num_pure_se3_iterations == so.num_pure_se3_iterations


def test_registration_bunny_full(ctx, orc, capi):
    """configs[1] at the full 34 834 points"""
    src, tgt, T_gt = W.bunny_problem("easy", seed=1)
    Tg, sg, To, so = run_both(ctx, orc, capi, src, tgt, "RUN_SE3_ICP", "pt2pl", **RRM)
    assert_transform_parity(Tg, To, tgt)
    assert np.degrees(rot_err(Tg, T_gt)) <= 2.0 and np.linalg.norm(Tg[:3, 3] - T_gt[:3, 3]) <= 0.25  # .cpp:410 criterion
    assert sg.num_iterations == so.num_iterations


@pytest.mark.parametrize("overlap,keep_largest", [(0.7, 0), (0.8, 1)])
def test_registration_trimmed(ctx, orc, capi, overlap, keep_largest):
    """trimmed rejection active (KITTI / lounge style parameters), both comparator directions"""
    src, tgt, _ = W.bunny_problem("easy", seed=4, n_points=4167)
    tgt = tgt[: int(0.85 * len(tgt))]  # partial overlap
    kw = dict(RRM, estimated_overlap=overlap, trim_keep_largest=keep_largest)
    Tg, sg, To, so = run_both(ctx, orc, capi, src, tgt, "RUN_SE3_ICP", "gicp", **kw)
    assert_transform_parity(Tg, To, tgt)
    assert sg.num_iterations == so.num_iterations


@pytest.mark.parametrize("problem", ["c1", "bunny"])
def test_registration_nn_modes_agree(ctx, capi, c1, bunny4k, problem):
    """pruned traversal, FP32 sweep + certified repair and the all-FP64 sweep give identical registrations"""
    src, tgt, _ = c1 if problem == "c1" else bunny4k
    ctx.set_cloud(capi.SOURCE, src)
    ctx.set_cloud(capi.TARGET, tgt)
    res = {}
    for name in ("NN_TREE", "NN_BRUTE_F32", "NN_EXACT_F64"):
        T, st = ctx.run(capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, nn_mode=getattr(capi, name), **RRM))
        res[name] = (T, st, ctx.correspondences())
    for name in ("NN_BRUTE_F32", "NN_EXACT_F64"):
        np.testing.assert_array_equal(res["NN_TREE"][0], res[name][0])
        np.testing.assert_array_equal(res["NN_TREE"][2][0], res[name][2][0])
        assert res["NN_TREE"][1].num_iterations == res[name][1].num_iterations
    assert res["NN_BRUTE_F32"][1].exact_repairs < res["NN_EXACT_F64"][1].exact_repairs


def test_history_and_mirrors(pkg, capi, c1):
    """run_icp fills estimated_history_ (reference .cpp:491,538); SE(3) clouds are readable (hpp:59-60)"""
    src, tgt, T_gt = c1
    reg = pkg.run_registration_method("pt2pl", src, tgt)
    assert len(reg.estimated_history_) == reg.num_iterations_ + 1
    np.testing.assert_array_equal(reg.estimated_history_[0], np.eye(4))
    acc = np.eye(4)
    for Ti in reg.estimated_history_[1:]:
        acc = Ti @ acc
    np.testing.assert_allclose(acc, reg.current_estimated_T_, atol=1e-12)
    reg = pkg.run_registration_method("se3_pt2pl", src, tgt)
    fs, ft = reg.source_se3_cloud_, reg.target_se3_cloud_
    assert fs.shape == (len(src), 4, 4) and ft.shape == (len(tgt), 4, 4)
    R = ft[:, :3, :3] / 3.0
    np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.broadcast_to(np.eye(3), R.shape), atol=1e-9)
    idx, dist = reg.current_correspondences_set
    assert np.mean(idx == np.arange(len(src))) > 0.9  # exact copy: almost every point finds its twin


def test_invalid_variant_behaviour(pkg, c1, capsys):
    src, tgt, _ = c1
    reg = pkg.IterativeSE3Registration()
    reg.setSourceCloud(src)
    reg.setTargetCloud(tgt)
    T = reg.run_se3_icp("bogus")
    np.testing.assert_allclose(T[:3, :3], np.eye(3))
    np.testing.assert_allclose(T[:3, 3], tgt.mean(0) - src.mean(0))
    assert "Invalid variant name" in capsys.readouterr().err


def test_pending_run_and_stage_calls_guard_the_context(capi, c1):
    """While an asynchronous run is pending the context refuses everything that would touch its buffers, and a
    stage-level call leaves it without clouds (SE3ICP_ERR_STATE instead of a run on half-overwritten data)."""
    src, tgt, _ = c1
    p = capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, **RRM)
    with capi.Context(0) as c:
        c.set_cloud(capi.SOURCE, src)
        c.set_cloud(capi.TARGET, tgt)
        T0, s0 = c.run(p)
        c.run_async(p)
        for call in (lambda: c.set_cloud(capi.SOURCE, src), lambda: c.run_async(p), lambda: c.swap_clouds(),
                     lambda: c.knn(src, 5), lambda: c.trim(np.ones(10, np.float32), 0.5)):
            with pytest.raises(capi.Se3IcpError) as e:
                call()
            assert e.value.code == 6  # SE3ICP_ERR_STATE
        T1, s1 = c.run_finish()
        np.testing.assert_array_equal(T0, T1)
        c.knn(src, 5)  # a stage call borrows the cloud slots ...
        with pytest.raises(capi.Se3IcpError) as e:
            c.run(p)   # ... so the context holds no clouds afterwards
        assert e.value.code == 6
        c.set_cloud(capi.SOURCE, src)
        c.set_cloud(capi.TARGET, tgt)
        T2, _ = c.run(p)
        np.testing.assert_array_equal(T0, T2)


def test_set_cloud_appends(pkg, c1):
    """the PointCloud overloads of setSourceCloud/setTargetCloud push_back (reference .cpp:359-362,373-375)"""
    src, tgt, T_gt = c1
    reg = pkg.IterativeSE3Registration()
    h = len(src) // 2
    reg.setSourceCloud(src[:h])
    reg.setSourceCloud(src[h:])
    reg.setTargetCloud(tgt)
    reg.max_num_se3_iterations_, reg.mse_switch_error_, reg.number_of_nn_for_LRF_ = 10, 5e-5, 90
    T = reg.run_se3_icp("pt2pl")
    assert rot_err(T, T_gt) < 1e-5


# ---------------------------------------------------------------------------------------------------
def test_kitti_scale_properties(ctx, orc, capi):
    """BASELINE.json configs[2] at full size (~120 k points): size-independent properties.
    The oracle is too slow for a per-stage comparison at this size in a unit test, so:
      - kNN lists are sorted, start with the point itself, and a 300-query sample equals the oracle's brute force
      - the SE(3) sweep equals the all-FP64 sweep on a 500-query sample
      - the registration recovers the synthetic ground truth"""
    src, tgt, T_gt = W.lidar_pair(seed=0)
    assert 100_000 < len(src) < 140_000
    gi, gd = ctx.knn(tgt, 90)
    assert np.all(np.diff(gd, axis=1) >= 0) and np.all(gd[:, 0] == 0)
    rng = np.random.default_rng(0)
    sample = rng.choice(len(tgt), 300, replace=False)
    full = ((tgt[sample, None, :] - tgt[None, :, :]) ** 2)
    d2 = (full[..., 0] + full[..., 1]) + full[..., 2]
    order = np.lexsort((np.broadcast_to(np.arange(len(tgt)), d2.shape), d2), axis=1)[:, :90]
    np.testing.assert_array_equal(gi[sample], order)
    # registration
    pg = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, **W.KITTI_PARAMS)
    ctx.set_cloud(capi.SOURCE, src)
    ctx.set_cloud(capi.TARGET, tgt)
    Tg, sg = ctx.run(pg)
    assert np.degrees(rot_err(Tg, T_gt)) < 0.1 and np.linalg.norm(Tg[:3, 3] - T_gt[:3, 3]) < 0.05
    assert 0 < sg.num_pure_se3_iterations <= 10
    gold = np.load(os.path.join(W.GOLDEN, "fullsize_oracle.npz"))  # oracle at full size (make_golden_fullsize.py)
    assert list(gold["kitti_n"]) == [len(src), len(tgt)]
    assert_transform_parity(Tg, gold["kitti_T"], tgt)
    assert [sg.num_iterations, sg.num_pure_se3_iterations] == list(gold["kitti_it"])


@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("config", ["kitti", "lounge"])
def test_fullsize_extra_seeds(ctx, capi, config, seed):
    """configs[2] / configs[3] at full size on two more seeds, against the committed oracle goldens."""
    gold = np.load(os.path.join(W.GOLDEN, "fullsize_oracle.npz"))
    if config == "kitti":
        src, tgt, _ = W.lidar_pair(seed=seed)
        p = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, **W.KITTI_PARAMS)
    else:
        src, tgt, _ = W.rgbd_pair(seed=seed)
        p = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP_CF, **W.LOUNGE_PARAMS)
    assert list(gold["%s_n_s%d" % (config, seed)]) == [len(src), len(tgt)]
    ctx.set_cloud(capi.SOURCE, src)
    ctx.set_cloud(capi.TARGET, tgt)
    Tg, sg = ctx.run(p)
    assert_transform_parity(Tg, gold["%s_T_s%d" % (config, seed)], tgt)
    assert [sg.num_iterations, sg.num_pure_se3_iterations] == list(gold["%s_it_s%d" % (config, seed)])


def test_lounge_scale_with_cf(ctx, orc, capi):
    """BASELINE.json configs[3]: lounge-like RGB-D pair, se3_gicp_with_cf (benchmark_lounge.cpp:183-186).
    Parity against the oracle on the stride-4 image (15 k points); at full resolution (~250 k points) the
    registration must recover the synthetic ground truth and report the reference's timing fields."""
    src, tgt, T_gt = W.rgbd_pair(seed=0, stride=4)
    Tg, sg, To, so = run_both(ctx, orc, capi, src, tgt, "RUN_SE3_ICP_CF", "gicp", **W.LOUNGE_PARAMS)
    assert_transform_parity(Tg, To, tgt)
    assert (sg.num_iterations, sg.num_pure_se3_iterations) == (so.num_iterations, so.num_pure_se3_iterations)
    src, tgt, T_gt = W.rgbd_pair(seed=0)
    assert len(src) > 200_000
    ctx.set_cloud(capi.SOURCE, src)
    ctx.set_cloud(capi.TARGET, tgt)
    Tg, sg = ctx.run(capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP_CF, **W.LOUNGE_PARAMS))
    gold = np.load(os.path.join(W.GOLDEN, "fullsize_oracle.npz"))  # oracle at full size (make_golden_fullsize.py)
    assert list(gold["lounge_n"]) == [len(src), len(tgt)]
    assert_transform_parity(Tg, gold["lounge_T"], tgt)
    assert [sg.num_iterations, sg.num_pure_se3_iterations] == list(gold["lounge_it"])
    # what the method reaches here with PCL's trimmed comparator (the default); benchmark_synthetic.cpp:410 asks for 2 deg
    assert np.degrees(rot_err(Tg, T_gt)) < 2.0 and np.linalg.norm(Tg[:3, 3] - T_gt[:3, 3]) < 0.15
    assert sg.time_se3_correspondence_search_ms > 0 and sg.time_before_pure_icp_ms >= sg.time_se3_correspondence_search_ms


# ---------------------------------------------------------------------------------------------------
def test_sharded_world1_equals_plain(ctx, capi, bunny4k):
    """se3icp_run_sharded with a one-rank NCCL communicator and the full range is the plain run"""
    src, tgt, _ = bunny4k
    ctx.set_cloud(capi.SOURCE, src)
    ctx.set_cloud(capi.TARGET, tgt)
    for variant, overlap in (("pt2pl", 1.0), ("gicp", 0.7)):
        p = capi.default_params(variant=variant, entry=capi.RUN_SE3_ICP, **dict(RRM, estimated_overlap=overlap))
        T1, s1 = ctx.run(p)
        ctx.comm_init(1, 0, capi.comm_unique_id())
        T2, s2 = ctx.run_sharded(p, 0, len(src))
        ctx.comm_destroy()
        np.testing.assert_array_equal(T1, T2)
        assert s1.num_iterations == s2.num_iterations
    with pytest.raises(capi.Se3IcpError):
        ctx.run_sharded(p, 0, len(src) // 2)  # partial range without a communicator


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_multi_gpu_sharded_and_batch(p2p):
    """needs >= 2 GPUs (gpurun --gpus 2): sharded pair == single GPU, batch sharding bit-identical.  Once with the
    all-reduce inside the iteration's last kernel over peer memory (loop = one CUDA graph), once through NCCL from the
    host-driven loop (SE3ICP_SHARDED_P2P=0)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(root, "tests", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, SE3ICP_SHARDED_P2P=p2p))
    assert "MULTI_GPU_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    assert ("loop=graph" in r.stdout) == (p2p == "1"), r.stdout[-2000:]


@pytest.mark.parametrize("entry_name,variant,overlap", [("RUN_SE3_ICP", "pt2pl", 1.0), ("RUN_SE3_ICP", "gicp", 0.7),
                                                         ("RUN_ICP", "pt2pt", 1.0), ("RUN_SE3_PURE", "pt2pl", 1.0)])
def test_graph_loop_equals_host_loop(ctx, capi, bunny4k, entry_name, variant, overlap):
    """use_graph=1 runs the whole iteration loop as one CUDA graph (conditional WHILE node, device-side stop
    flag); the result must be bit-identical to the host-polled loop"""
    src, tgt, _ = bunny4k
    ctx.set_cloud(capi.SOURCE, src)
    ctx.set_cloud(capi.TARGET, tgt)
    kw = dict(RRM, estimated_overlap=overlap)
    Ta, sa = ctx.run(capi.default_params(variant=variant, entry=getattr(capi, entry_name), use_graph=0, **kw))
    Tb, sb = ctx.run(capi.default_params(variant=variant, entry=getattr(capi, entry_name), use_graph=1, **kw))
    np.testing.assert_array_equal(Ta, Tb)
    assert (sa.num_iterations, sa.num_pure_se3_iterations) == (sb.num_iterations, sb.num_pure_se3_iterations)
    assert sb.kernel_launches > 0


@pytest.mark.parametrize("n_ctx,n_pairs", [(1, 5), (3, 7), (4, 2), (2, 0)])
def test_batch_runner_equals_single_runs(capi, n_ctx, n_pairs):
    """se3icp_run_batch (two enqueue threads walking round-robin over the contexts, asynchronous runs) returns, pair
    by pair, exactly what a single blocking run returns — more pairs than contexts, fewer, and none; host buffers and
    device-resident ones"""
    import torch
    probs = [W.bunny_problem("easy", seed=20 + i, n_points=3000 + 137 * i) for i in range(n_pairs)]
    p = capi.default_params(variant="gicp", entry=capi.RUN_SE3_ICP, **dict(RRM, estimated_overlap=0.9))
    single = []
    with capi.Context(0) as c:
        for s, t, _ in probs:
            c.set_cloud(capi.SOURCE, s)
            c.set_cloud(capi.TARGET, t)
            single.append(c.run(p))
    ctxs = [capi.Context(0) for _ in range(n_ctx)]
    try:
        T, st = capi.run_batch(ctxs, [(s, t) for s, t, _ in probs], p)
        dev = [(torch.from_numpy(np.ascontiguousarray(s)).cuda(), torch.from_numpy(np.ascontiguousarray(t)).cuda()) for s, t, _ in probs]
        Td, std = capi.run_batch(ctxs, [(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0]) for a, b in dev], p, device_inputs=True)
    finally:
        for c in ctxs:
            c.close()
    assert T.shape == (n_pairs, 4, 4) and len(st) == n_pairs
    for k, (Ts, ss) in enumerate(single):
        np.testing.assert_array_equal(T[k], Ts)
        np.testing.assert_array_equal(Td[k], Ts)
        assert st[k].num_iterations == ss.num_iterations == std[k].num_iterations
        assert st[k].loop_was_graph == 1


def test_loop_graph_is_kept_and_updated_in_place(capi, c1, bunny4k):
    """The context instantiates its loop graph once and re-parameterises it for later pairs (other sizes, other
    buffers, other parameters); only another launch sequence (trimming on/off) may build a new executable.  Results
    must equal those of a fresh context."""
    problems = [c1, bunny4k, (W.bunny_problem("easy", seed=5, n_points=9000)), c1]
    fresh = []
    for src, tgt, _ in problems:
        with capi.Context(0) as f:
            f.set_cloud(capi.SOURCE, src)
            f.set_cloud(capi.TARGET, tgt)
            fresh.append(f.run(capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, **RRM)))
    with capi.Context(0) as c:
        for (src, tgt, _), (Tf, sf) in zip(problems, fresh):
            c.set_cloud(capi.SOURCE, src)
            c.set_cloud(capi.TARGET, tgt)
            T, s = c.run(capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, **RRM))
            np.testing.assert_array_equal(T, Tf)
            assert (s.num_iterations, s.num_pure_se3_iterations) == (sf.num_iterations, sf.num_pure_se3_iterations)
            assert s.graph_instantiations == 1
        # another variant and entry: same launch sequence, still the same executable
        T, s = c.run(capi.default_params(variant="gicp", entry=capi.RUN_ICP, **RRM))
        assert s.graph_instantiations == 1
        # trimming adds the two selection kernels to the iteration: a new executable, once
        for _ in range(2):
            T, s = c.run(capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, **dict(RRM, estimated_overlap=0.8)))
            assert s.graph_instantiations == 2


@pytest.mark.parametrize("problem", ["c1", "bunny", "kitti"])
def test_coherence_filter_is_exact(ctx, capi, c1, bunny4k, problem):
    """nn_coherence=1 (skip queries whose remembered match is provably still nearest) must not change a bit"""
    if problem == "kitti":
        src, tgt, _ = W.lidar_pair(seed=1)
        kw = dict(W.KITTI_PARAMS)
        variant = "gicp"
    else:
        src, tgt, _ = c1 if problem == "c1" else bunny4k
        kw = dict(RRM)
        variant = "pt2pl"
    ctx.set_cloud(capi.SOURCE, src)
    ctx.set_cloud(capi.TARGET, tgt)
    for entry in (capi.RUN_SE3_ICP, capi.RUN_ICP):
        Ta, sa = ctx.run(capi.default_params(variant=variant, entry=entry, nn_coherence=0, **kw))
        ia, da = ctx.correspondences()
        Tb, sb = ctx.run(capi.default_params(variant=variant, entry=entry, nn_coherence=1, **kw))
        ib, db = ctx.correspondences()
        np.testing.assert_array_equal(Ta, Tb)
        np.testing.assert_array_equal(ia, ib)
        np.testing.assert_array_equal(da, db)
        assert (sa.num_iterations, sa.num_pure_se3_iterations) == (sb.num_iterations, sb.num_pure_se3_iterations)


# ---------------------------------------------------------------------------------------------------
# CUDA path against the REFERENCE'S OWN SOURCE: tests/golden/reference_build.npz holds what the unmodified
# src/iterative_SE3_registration.cpp returns (oracle/Makefile `ref`, tests/golden/make_golden_reference.py).
import importlib.util  # noqa: E402
import json  # noqa: E402

_mg_spec = importlib.util.spec_from_file_location("make_golden_reference", os.path.join(W.GOLDEN, "make_golden_reference.py"))
_MG = importlib.util.module_from_spec(_mg_spec)
_mg_spec.loader.exec_module(_MG)
with open(os.path.join(W.GOLDEN, "reference_build_cases.json")) as _f:
    REF_CASES = json.load(_f)
REF_ENTRY = {"icp": "RUN_ICP", "se3": "RUN_SE3_ICP", "cf": "RUN_SE3_ICP_CF", "pure": "RUN_SE3_PURE"}


@pytest.mark.parametrize("name", sorted(REF_CASES))
def test_cuda_path_reproduces_reference_source(ctx, capi, name):
    gold = np.load(os.path.join(W.GOLDEN, "reference_build.npz"))
    c = REF_CASES[name]
    src, tgt = _MG.load_cloud(*c["cloud"])
    assert [len(src), len(tgt)] == list(gold[name + "/n"])
    ctx.set_cloud(capi.SOURCE, src)
    ctx.set_cloud(capi.TARGET, tgt)
    Tg, sg = ctx.run(capi.default_params(variant=c["variant"], entry=getattr(capi, REF_ENTRY[c["entry"]]), **c["params"]))
    assert [sg.num_iterations, sg.num_pure_se3_iterations] == list(gold[name + "/it"])
    assert_transform_parity(Tg, gold[name + "/T"], tgt)


# ---------------------------------------------------------------------------------------------------
# Sequence driver: every scan goes through the kNN-90 / LRF / normal stage once.
@pytest.mark.parametrize("variant", ["gicp", "pt2pt", "pt2pl"])
def test_sequence_reuses_features_and_matches_independent_pairs(ctx, capi, variant):
    scans, steps = W.lidar_sequence(seed=1, n_scans=4, n_rings=32, n_az=500)
    p = capi.default_params(variant=variant, entry=capi.RUN_SE3_ICP, **W.KITTI_PARAMS)
    T_seq, st_seq = ctx.run_sequence(scans, p)
    assert [s.feature_reuses for s in st_seq] == ([0, 1, 1] if variant != "pt2pl" else [0, 0, 0])  # pt2pl: target-only normals
    for i in range(3):
        ctx.set_cloud(capi.SOURCE, scans[i + 1])
        ctx.set_cloud(capi.TARGET, scans[i])
        T_ind, st_ind = ctx.run(p)
        assert st_ind.feature_reuses == 0
        assert (st_seq[i].num_iterations, st_seq[i].num_pure_se3_iterations) == (st_ind.num_iterations, st_ind.num_pure_se3_iterations)
        np.testing.assert_allclose(T_seq[i], T_ind, rtol=0, atol=1e-9)
        assert rot_err(T_seq[i], steps[i]) < np.radians(0.5)
    # the switch: no reuse when asked not to, and a run on unchanged clouds may reuse both
    p_off = capi.default_params(variant=variant, entry=capi.RUN_SE3_ICP, reuse_features=0, **W.KITTI_PARAMS)
    assert all(s.feature_reuses == 0 for s in ctx.run_sequence(scans, p_off)[1])
    if variant == "gicp":
        T_again, st_again = ctx.run(p)
        T_first, _ = ctx.run(p_off)
        assert st_again.feature_reuses == 2
        np.testing.assert_array_equal(T_again, T_first)


def test_swap_clouds_registers_the_inverse_problem(ctx, capi, c1):
    src, tgt, T_gt = c1
    p = capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, **RRM)
    ctx.set_cloud(capi.SOURCE, src)
    ctx.set_cloud(capi.TARGET, tgt)
    T_fwd, _ = ctx.run(p)
    ctx.swap_clouds()
    T_bwd, st = ctx.run(p)
    assert rot_err(T_bwd, np.linalg.inv(T_gt)) < 1e-4 and rot_err(T_fwd, T_gt) < 1e-4
    ctx.set_cloud(capi.SOURCE, tgt)
    ctx.set_cloud(capi.TARGET, src)
    T_ref, _ = ctx.run(p)
    np.testing.assert_allclose(T_bwd, T_ref, rtol=0, atol=1e-9)


# ---------------------------------------------------------------------------------------------------
# Evaluation helpers either side of the path: reference src/cc.cpp and Open3D RandomDownSample.
def test_eval_helpers_match_the_cc_library(ctx, orc, capi, c1):
    src, tgt, T_gt = c1
    rng = np.random.default_rng(3)
    T_est = W.make_T(W.rot_3d(*rng.uniform(-0.02, 0.02, 3)), rng.uniform(-0.01, 0.01, 3)) @ T_gt
    # cc::error_filterreg
    for n in (1, 33, len(src)):
        g = ctx.eval_error_filterreg(src[:n], T_gt, T_est)
        assert abs(g - orc.cc_error_filterreg(src[:n], T_gt, T_est)) <= 1e-12 * max(1.0, g)
    assert ctx.eval_error_filterreg(src, T_gt, T_gt) == 0.0
    # cc::compute_corrs_with_gt: the fixture's target is T_gt * source point by point (permuted or not): exact NN
    gi = ctx.eval_corrs_with_gt(src, tgt, T_gt)
    np.testing.assert_array_equal(gi, orc.cc_corrs_with_gt(src, tgt, T_gt))
    moved = W.apply_T(T_gt, src)
    assert np.linalg.norm(moved - tgt[gi], axis=1).max() < 1e-9
    # cc::evaluate_LRF_quality on the LRFs of both clouds over the ground-truth correspondences
    fs, ft = ctx.lrf(src, 90), ctx.lrf(tgt, 90)
    pairs = np.stack([np.arange(len(src)), gi], 1)[::7]
    g_mean, g_per = ctx.eval_lrf_quality(fs, ft, T_gt, pairs)
    o_mean, o_per = orc.cc_lrf_quality(fs, ft, T_gt, pairs)
    ok = np.isfinite(o_per)
    np.testing.assert_allclose(g_per[ok], o_per[ok], rtol=0, atol=1e-6)  # degrees; acos near 1 amplifies rounding
    assert np.array_equal(np.isfinite(g_per), ok)
    if ok.all():
        assert abs(g_mean - o_mean) < 1e-6
        # (not small even for this exact copy: the reference's TOLDI centroid sums k/3 - 1 points and divides by k/3,
        # .cpp:259-265, so the frame depends on where the origin is and is not equivariant under T_gt — the quantity
        # this metric exists to expose)
    with pytest.raises(capi.Se3IcpError):
        ctx.eval_lrf_quality(fs, ft, T_gt, [[0, len(tgt)]])  # pair out of range


@pytest.mark.parametrize("ratio", [0.02, 0.5, 1.0, 0.0])
def test_random_downsample_is_a_uniform_subset(ctx, ratio):
    """Open3D RandomDownSample as benchmark_synthetic.cpp:100,150 calls it: (size_t)(n * ratio) distinct points, every
    point equally likely, deterministic for a seed, different for another seed"""
    pts = W.load_bunny()
    n = len(pts)
    sub, idx = ctx.random_downsample(pts, ratio, seed=7)
    k = int(n * ratio)
    assert sub.shape == (k, 3) and idx.shape == (k,)
    if k == 0:
        return
    assert len(np.unique(idx)) == k and idx.min() >= 0 and idx.max() < n
    np.testing.assert_array_equal(sub, pts[idx])
    sub2, idx2 = ctx.random_downsample(pts, ratio, seed=7)
    np.testing.assert_array_equal(idx, idx2)
    if 0 < k < n:
        _, idx3 = ctx.random_downsample(pts, ratio, seed=8)
        assert not np.array_equal(idx, idx3)
        assert not np.array_equal(idx, np.sort(idx))  # shuffled order, as Open3D returns it
        # uniformity: the chosen indices spread evenly over the index range (chi-square over 16 bins, generous bound)
        hist = np.histogram(idx, bins=16, range=(0, n))[0]
        expect = k / 16.0
        assert ((hist - expect) ** 2 / expect).sum() < 60.0


# ---------------------------------------------------------------------------------------------------
# Ragged and tiny inputs: fewer points than the neighbourhood sizes, unequal cloud sizes, partial last leaves.
@pytest.mark.parametrize("n_src,n_tgt", [(40, 40), (33, 500), (500, 33), (95, 1000), (1025, 257)])
@pytest.mark.parametrize("variant", ["pt2pt", "pt2pl", "gicp"])
def test_registration_ragged_sizes(ctx, orc, capi, variant, n_src, n_tgt):
    rng = np.random.default_rng(n_src * 7919 + n_tgt)
    base = W.load_bunny()
    T = W.make_T(W.rot_3d(0.05, -0.04, 0.08), [0.4, -0.3, 0.2])
    tgt = base[rng.choice(len(base), n_tgt, replace=False)]
    src = W.apply_T(np.linalg.inv(T), base[rng.choice(len(base), n_src, replace=False)])
    kw = dict(RRM, max_num_iterations=30)
    Tg, sg, To, so = run_both(ctx, orc, capi, src, tgt, "RUN_SE3_ICP", variant, **kw)
    assert (sg.num_iterations, sg.num_pure_se3_iterations) == (so.num_iterations, so.num_pure_se3_iterations)
    if np.all(np.isfinite(To)):
        assert_transform_parity(Tg, To, tgt)
    else:  # a degenerate system may produce NaN in the reference arithmetic; the CUDA path must do the same
        assert np.array_equal(np.isfinite(Tg), np.isfinite(To))


def test_empty_cloud_is_an_error_not_a_crash(pkg, capi, c1):
    src, tgt, _ = c1
    ctx = capi.Context(0)
    ctx.set_cloud(capi.SOURCE, src)
    with pytest.raises(RuntimeError):
        ctx.run(capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, **RRM))  # no target
    ctx.set_cloud(capi.TARGET, np.zeros((0, 3)))
    with pytest.raises(RuntimeError):
        ctx.run(capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, **RRM))
    ctx.set_cloud(capi.TARGET, tgt)
    T, st = ctx.run(capi.default_params(variant="pt2pl", entry=capi.RUN_SE3_ICP, **RRM))  # and the context still works
    assert st.num_iterations == 8
    ctx.close()


# ---------------------------------------------------------------------------------------------------
# CUDA path against the reference's own source, live: oracle/_ref/libse3icp_reference.so is built where /root/reference
# exists and travels to the GPU box with the snapshot (it is not read from /root/reference at run time).
from oracle import reference_build as _RB  # noqa: E402


@pytest.mark.skipif(not os.path.exists(_RB.LIB_PATH), reason="oracle/_ref/libse3icp_reference.so did not travel")
@pytest.mark.parametrize("seed", range(10))
def test_cuda_path_matches_reference_source_on_random_problems(ctx, capi, seed):
    rng = np.random.default_rng(4321 + seed)
    base = W.load_bunny()
    n_s, n_t = int(rng.integers(300, 1500)), int(rng.integers(300, 1500))
    T = W.make_T(W.rot_3d(*rng.uniform(-0.3, 0.3, 3)), rng.uniform(-1.0, 1.0, 3))
    tgt = base[rng.choice(len(base), n_t, replace=False)] + rng.normal(0, 0.02, (n_t, 3))
    src = W.apply_T(np.linalg.inv(T), base[rng.choice(len(base), n_s, replace=False)] + rng.normal(0, 0.02, (n_s, 3)))
    entry = ["icp", "se3", "se3", "pure", "cf"][seed % 5]
    variant = "gicp" if entry == "cf" else ["pt2pt", "pt2pl", "gicp"][(seed // 5 + seed) % 3]
    params = dict(max_num_iterations=int(rng.integers(5, 40)), max_num_se3_iterations=int(rng.integers(2, 12)),
                  number_of_nn_for_LRF=int(rng.choice([12, 30, 45, 90])), mse=float(10 ** rng.uniform(-7, -4)),
                  mse_switch_error=float(10 ** rng.uniform(-5, -2)), estimated_overlap=float(rng.choice([1.0, 0.9, 0.73, 0.5])),
                  alpha_rot=float(rng.uniform(0.5, 4.0)), beta_transl=float(rng.uniform(0.5, 2.0)),
                  scale_preprocessing=float(rng.uniform(1.0, 5.0)), trim_keep_largest=int(seed % 7 == 3))
    Tr, it, it_se3 = _RB.run(_MG.ENTRY_OF[entry], variant, src, tgt, _RB.default_params(**params))
    ctx.set_cloud(capi.SOURCE, src)
    ctx.set_cloud(capi.TARGET, tgt)
    Tg, sg = ctx.run(capi.default_params(variant=variant, entry=getattr(capi, REF_ENTRY[entry]), **params))
    assert (sg.num_iterations, sg.num_pure_se3_iterations) == (it, it_se3), (entry, variant, params)
    if np.all(np.isfinite(Tr)):
        assert_transform_parity(Tg, Tr, tgt)
    else:
        assert np.array_equal(np.isfinite(Tg), np.isfinite(Tr))
