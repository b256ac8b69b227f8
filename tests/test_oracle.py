"""CPU tests: the oracle against the reference's only known answers and against
independent numpy restatements.  The reference ships no tests; the pins are the exact-copy fixture
with its ground truth, the survey's known iteration count, and self-consistency properties."""
import numpy as np
import pytest

import workloads as W
from conftest import rot_err

RRM = dict(estimated_overlap=1.0, max_num_se3_iterations=10, mse=1e-5, mse_switch_error=5e-5,
           number_of_nn_for_LRF=90)  # examples/run_registration_method.cpp:38-42


@pytest.mark.parametrize("entry_name,variant", [
    ("RUN_SE3_ICP", "pt2pt"), ("RUN_SE3_ICP", "pt2pl"), ("RUN_SE3_ICP", "gicp"),
    ("RUN_ICP", "pt2pt"), ("RUN_ICP", "pt2pl"), ("RUN_ICP", "gicp"),
    ("RUN_SE3_PURE", "pt2pl"), ("RUN_SE3_ICP_CF", "gicp")])
def test_fixture_ground_truth(orc, c1, entry_name, variant):
    """created_example_reg_problem: target = GT * source exactly, every converging method returns GT."""
    src, tgt, T_gt = c1
    p = orc.default_params(variant=variant, entry=getattr(orc, entry_name), **RRM)
    T, st, _ = orc.run(src, tgt, p)
    assert rot_err(T, T_gt) < 1e-6
    assert np.linalg.norm(T[:3, 3] - T_gt[:3, 3]) < 1e-6
    np.testing.assert_allclose(T[3], [0, 0, 0, 1])


def test_fixture_iteration_count(orc, c1):
    """SURVEY §4 known answer: se3_pt2pl converges in 8 iterations = 6 SE(3) + 2 ICP."""
    src, tgt, _ = c1
    p = orc.default_params(variant="pt2pl", entry=orc.RUN_SE3_ICP, **RRM)
    _, st, tr = orc.run(src, tgt, p, trace_iters=20)
    assert (st.num_iterations, st.num_pure_se3_iterations) == (8, 6)
    assert list(tr["se3_phase"]) == [1] * 6 + [0] * 2
    assert abs(st.scaling_factor - 3.0 / max(np.linalg.norm(src - src.mean(0), axis=1).max(),
                                               np.linalg.norm(tgt - tgt.mean(0), axis=1).max())) < 1e-12


def test_golden_trace(orc, c1):
    """Committed golden vectors (tests/golden/make_golden.py) still reproduce."""
    import os
    g = np.load(os.path.join(W.GOLDEN, "c1_se3_pt2pl_trace.npz"))
    src, tgt, _ = c1
    p = orc.default_params(variant="pt2pl", entry=orc.RUN_SE3_ICP, **RRM)
    T, st, tr = orc.run(src, tgt, p, trace_iters=20)
    assert st.num_iterations == int(g["num_iterations"]) and st.num_pure_se3_iterations == int(g["num_se3"])
    np.testing.assert_allclose(T, g["T_final"], atol=1e-9)
    np.testing.assert_allclose(tr["T_iter"], g["T_iter"], atol=1e-7)
    np.testing.assert_allclose(tr["mean_dist"], g["mean_dist"], rtol=1e-6, atol=1e-9)
    # correspondences of the first SE(3) iteration: identical except exact duplicates (equal distance)
    a, b = tr["corr_idx"][0], g["corr_idx0"]
    diff = a != b
    assert np.all(np.abs(tr["corr_dist"][0][diff] - g["corr_dist0"][diff]) <= 1e-6 * g["corr_dist0"][diff])


@pytest.mark.parametrize("dim,n,m", [(3, 500, 2000), (12, 300, 1500)])
def test_kdtree_equals_brute_force(orc, dim, n, m):
    rng = np.random.default_rng(dim)
    q, d = rng.normal(size=(n, dim)), rng.normal(size=(m, dim))
    d[7] = d[3]  # exact duplicate -> tie resolves to the smaller index in both
    i1, d1 = orc.nn(q, d)
    i2, d2 = orc.nn(q, d, brute=True)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(d1, d2)
    ref = ((q[:, None, :] - d[None, :, :]) ** 2).sum(-1)
    np.testing.assert_array_equal(i1, ref.argmin(1))


def test_knn_sorted_and_exact(orc):
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(700, 3))
    idx, d2 = orc.knn_self(pts, 25)
    assert np.all(np.diff(d2, axis=1) >= 0)
    assert np.all(idx[:, 0] == np.arange(700)) and np.all(d2[:, 0] == 0)
    full = ((pts[:, None] - pts[None]) ** 2).sum(-1)
    ref = np.argsort(full, axis=1, kind="stable")[:, :25]
    np.testing.assert_array_equal(idx, ref)


def test_toldi_is_rotation_and_equivariant(orc, c1):
    """frames of a rigidly moved cloud equal R_gt * frames."""
    src, tgt, T_gt = c1
    # The reference's centroid quirk (sum of rz-1 points divided by rz, .cpp:261-265) makes the frame depend
    # on the absolute position, so equivariance only holds for the centred clouds the algorithm really uses.
    src, tgt = (src - src.mean(0)) * 0.5, (tgt - tgt.mean(0)) * 0.5
    fs, ft = orc.toldi(src, 90), orc.toldi(tgt, 90)
    R = fs[:, :3, :3]
    np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.broadcast_to(np.eye(3), R.shape), atol=1e-9)
    np.testing.assert_allclose(np.linalg.det(R), 1.0, atol=1e-9)
    np.testing.assert_array_equal(fs[:, :3, 3], src)
    pred = T_gt[:3, :3][None] @ R
    err = np.abs(pred - ft[:, :3, :3]).max(axis=(1, 2))
    assert np.mean(err < 1e-6) > 0.95  # duplicates / near-ties in the kNN set may flip isolated frames


def test_toldi_numpy_restatement(orc):
    """Independent numpy restatement of reference .cpp:241-316 on a small cloud."""
    rng = np.random.default_rng(5)
    pts = rng.normal(size=(200, 3)) * [1.0, 0.7, 0.2]
    k = 30
    fr = orc.toldi(pts, k)
    full = ((pts[:, None] - pts[None]) ** 2).sum(-1)
    for i in (0, 17, 199):
        nb = np.argsort(full[i], kind="stable")[:k]
        P = pts[nb]
        radius = np.linalg.norm(pts[i] - P[-1])
        rz = k // 3
        c = P[1:rz].sum(0) / rz
        D = P[1:rz + 1] - c
        w, V = np.linalg.eigh(D.T @ D)
        n = V[:, 0]
        A = P[1:] - pts[i]
        wts = (radius - np.linalg.norm(A, axis=1)) ** 2 * (A @ n) ** 2
        if n @ A.sum(0) < 0:
            n = -n
        aw = (wts[:, None] * A).sum(0)
        x = aw - (aw @ n) * n
        x /= np.linalg.norm(x)
        np.testing.assert_allclose(fr[i, :3, 2], n, atol=1e-9)
        np.testing.assert_allclose(fr[i, :3, 0], x, atol=1e-9)
        np.testing.assert_allclose(fr[i, :3, 1], np.cross(n, x), atol=1e-9)


def test_normals_plane_and_cov_quirk(orc):
    rng = np.random.default_rng(1)
    pts = np.c_[rng.uniform(-1, 1, (400, 2)), np.zeros(400)]
    n = orc.normals(pts, 30)
    np.testing.assert_allclose(np.abs(n[:, 2]), 1.0, atol=1e-9)
    nr = np.array([[1.0, 0, 0], [-1.0, 0, 0], [-0.995, 0.0998749, 0], [0, 0, 1.0]])
    C = orc.gicp_cov(nr, 1e-3)
    np.testing.assert_allclose(C[0], np.diag([1e-3, 1, 1]), atol=1e-12)
    # reference .cpp:8-10: c < -0.99 -> Identity rotation, so the covariance is NOT aligned with the normal
    np.testing.assert_allclose(C[1], np.diag([1e-3, 1, 1]), atol=1e-12)
    np.testing.assert_allclose(C[2], np.diag([1e-3, 1, 1]), atol=1e-12)
    np.testing.assert_allclose(C[3], np.diag([1, 1, 1e-3]), atol=1e-12)


def test_trim_counts(orc):
    d = np.arange(10, dtype=np.float32)[::-1].copy()
    k, keep = orc.trim(d, 1.0)
    assert k == 10 and keep.all()
    k, keep = orc.trim(d, 0.5)
    assert k == 5 and set(np.nonzero(keep)[0]) == {5, 6, 7, 8, 9}
    k, keep = orc.trim(d, 0.5, keep_largest=True)
    assert k == 5 and set(np.nonzero(keep)[0]) == {0, 1, 2, 3, 4}
    # the ratio is a float (setOverlapRatio(float)): float(0.7) * 10 = 7.0000000298 -> 7, float(0.7)*1000 -> 699
    n = 1000
    k, _ = orc.trim(np.zeros(n, np.float32), 0.7)
    assert k == int(np.floor(np.float32(0.7) * np.float32(n)))


def test_solve_and_umeyama(orc):
    rng = np.random.default_rng(2)
    J = rng.normal(size=(50, 6))
    r = rng.normal(size=50) * 0.01
    JTJ, JTr = J.T @ J, J.T @ r
    in27 = np.concatenate([JTJ[np.triu_indices(6)], JTr])
    T = orc.solve6(in27)
    x = np.linalg.solve(JTJ, -JTr)
    np.testing.assert_allclose(T[:3, :3], W.rot_3d(x[0], x[1], x[2]), atol=1e-12)
    np.testing.assert_allclose(T[:3, 3], x[3:], atol=1e-12)
    src = rng.normal(size=(100, 3))
    Tg = W.make_T(W.rot_3d(0.3, -1.1, 2.0), [1, -2, 0.5])
    tgt = W.apply_T(Tg, src)
    idx = np.arange(100, dtype=np.int32)
    np.testing.assert_allclose(orc.umeyama(src, tgt, idx, idx), Tg, atol=1e-12)
    # reflection case: planar data mirrored -> still a proper rotation
    src2 = src * [1, 1, 0]
    Tm = orc.umeyama(src2, src2 * [1, -1, 0], idx, idx)
    assert abs(np.linalg.det(Tm[:3, :3]) - 1) < 1e-9


def test_reduce_matches_numpy(orc):
    rng = np.random.default_rng(3)
    n = 64
    src, tgt = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    cs = np.arange(n, dtype=np.int32)
    ct = rng.permutation(n).astype(np.int32)
    out = orc.reduce_pt2pl(src, tgt, nrm, cs, ct)
    s, t, nn = src[cs], tgt[ct], nrm[ct]
    J = np.c_[np.cross(s, nn), nn]
    r = ((s - t) * nn).sum(1)
    np.testing.assert_allclose(out[:21], (J.T @ J)[np.triu_indices(6)], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out[21:], J.T @ r, rtol=1e-12, atol=1e-12)
    # GICP: JTJ = sum w^2 A^T M^-1 A, JTr = sum w^2 A^T M^-1 d
    Cs, Ct = orc.gicp_cov(nrm, 1e-3), orc.gicp_cov(np.roll(nrm, 1, 0), 1e-3)
    w = rng.uniform(0.5, 1.5, n)
    out = orc.reduce_gicp(src, Cs, tgt, Ct, cs, ct, w)
    JTJ, JTr = np.zeros((6, 6)), np.zeros(6)
    for i in range(n):
        sv = src[cs[i]]
        S = np.array([[0, -sv[2], sv[1]], [sv[2], 0, -sv[0]], [-sv[1], sv[0], 0]])
        A = np.c_[-S, np.eye(3)]
        Minv = np.linalg.inv(Ct[ct[i]] + Cs[cs[i]])
        JTJ += w[i] ** 2 * A.T @ Minv @ A
        JTr += w[i] ** 2 * A.T @ Minv @ (sv - tgt[ct[i]])
    np.testing.assert_allclose(out[:21], JTJ[np.triu_indices(6)], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(out[21:], JTr, rtol=1e-10, atol=1e-10)


def test_eig3(orc):
    rng = np.random.default_rng(4)
    for _ in range(20):
        B = rng.normal(size=(3, 3))
        A = B @ B.T
        ev, V = orc.eig3(A)
        w, _ = np.linalg.eigh(A)
        np.testing.assert_allclose(ev, w, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(A @ V, V * ev, atol=1e-12)


def test_bunny_levels_converge(orc):
    """benchmark_synthetic.cpp:410 success criterion: SO(3) error <= 2 deg and |dt| <= 0.25."""
    for level, seed in (("easy", 1), ("moderate", 2)):  # the method is not expected to succeed on every draw
        src, tgt, T_gt = W.bunny_problem(level, seed=seed, n_points=4167)
        p = orc.default_params(variant="pt2pl", entry=orc.RUN_SE3_ICP, **RRM)
        T, st, _ = orc.run(src, tgt, p)
        assert np.degrees(rot_err(T, T_gt)) <= 2.0 and np.linalg.norm(T[:3, 3] - T_gt[:3, 3]) <= 0.25
