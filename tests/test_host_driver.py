"""The reference's own driver (examples/run_registration_method.cpp, compiled UNCHANGED against
include/iterative_SE3_registration.hpp + compat/ by se3-icp_b200/host/Makefile) linked to the CUDA path."""
import os
import subprocess

import numpy as np
import pytest

import workloads as W
from conftest import rot_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "se3-icp_b200", "run_registration_method")
BIN_DIR = os.path.join(ROOT, "se3-icp_b200", "bin")
ALL_DRIVERS = ["run_registration_method", "registration_example", "registration_example_fgr", "benchmark_synthetic",
               "benchmark_lounge", "test_se3_pure", "benchmark_kitti", "benchmark_extreme_noise_bunny",
               "create_and_save_reg_problem"]  # reference CMakeLists.txt:42-76


def write_ply(path, pts):
    """binary little-endian PLY with double x/y/z — the format of the reference's bundled fixture"""
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\ncomment Created by Open3D\nelement vertex %d\n"
                 "property double x\nproperty double y\nproperty double z\nend_header\n" % len(pts)).encode())
        f.write(np.ascontiguousarray(pts, dtype="<f8").tobytes())


def test_driver_usage_and_bad_name(tmp_path):
    assert os.path.exists(BIN), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    r = subprocess.run([BIN], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr                      # run_registration_method.cpp:10-13
    r = subprocess.run([BIN, "nonsense", "a.ply", "b.ply"], capture_output=True, text=True)
    assert r.returncode == 1 and "Not a valid algorithm name" in r.stderr  # :19-24


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["se3_pt2pl", "se3_pt2pt", "se3_gicp", "pt2pt", "pt2pl", "gicp"])
def test_driver_registers_fixture(tmp_path, method):
    """README command: ./run_registration_method se3_pt2pl source.ply target.ply -> ground truth"""
    src, tgt, T_gt = W.load_c1()
    write_ply(tmp_path / "source.ply", src)
    write_ply(tmp_path / "target.ply", tgt)
    r = subprocess.run([BIN, method, str(tmp_path / "source.ply"), str(tmp_path / "target.ply")], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "source point cloud size = 4167" in r.stdout
    lines = r.stdout.split("Estimated transformation =")[1].strip().splitlines()
    T = np.array([[float(v) for v in ln.split()] for ln in lines[:4]])
    assert rot_err(T, T_gt) < 1e-4 and np.linalg.norm(T[:3, 3] - T_gt[:3, 3]) < 1e-4  # 6 printed digits


def test_all_reference_drivers_build_unchanged():
    """every executable of the reference's CMakeLists.txt links against the drop-in library (built by
    se3-icp_b200/host/Makefile from the unmodified sources under /root/reference)"""
    for d in ALL_DRIVERS + ["libcc.so"]:
        assert os.path.exists(os.path.join(BIN_DIR, d)), d
    for d in ("benchmark_synthetic", "benchmark_kitti", "benchmark_lounge"):
        r = subprocess.run([os.path.join(BIN_DIR, d)], capture_output=True, text=True)
        assert r.returncode == 1 and "Usage" in r.stderr


def _last_float(text, label):
    line = [ln for ln in text.splitlines() if label in ln][-1]
    return float(line.split("=")[-1].split("(")[0])


@pytest.mark.gpu
def test_registration_example_driver(tmp_path):
    """examples/registration_example.cpp: bunny, 2 % random down-sample, run_se3_icp("pt2pl"), prints estimate + GT"""
    build = tmp_path / "build"
    build.mkdir()
    W.write_ply(tmp_path / "stanford_bunny.ply", np.load(os.path.join(W.GOLDEN, "bunny_unique_f32.npy")), dtype="<f4")
    r = subprocess.run([os.path.join(BIN_DIR, "registration_example")], cwd=build, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    est = r.stdout.split("Estimated transformation =")[1].split("Ground truth")[0].strip().splitlines()
    gt = r.stdout.split("Ground truth transformation =")[1].strip().splitlines()
    T = np.array([[float(v) for v in ln.split()] for ln in est[:4]])
    G = np.array([[float(v) for v in ln.split()] for ln in gt[:4]])
    assert np.degrees(rot_err(T, G)) < 2.0 and np.linalg.norm(T[:3, 3] - G[:3, 3]) < 0.25 * 0.02  # bunny not scaled here


@pytest.mark.gpu
def test_benchmark_synthetic_driver(tmp_path):
    """examples/benchmark_synthetic.cpp on a folder in its own on-disk format (gt_data + source<i>/target<i>.ply)"""
    problems = [W.bunny_problem("easy", seed=s, n_points=4167) for s in (1, 2, 3)]
    W.write_synthetic_dataset(str(tmp_path / "easy_data"), problems)
    for method in ("se3_pt2pl", "se3_gicp", "pt2pl"):
        r = subprocess.run([os.path.join(BIN_DIR, "benchmark_synthetic"), method, str(tmp_path / "easy_data")],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert "Num of fails over 3 problems is: 0" in r.stdout, r.stdout[-800:]
        assert _last_float(r.stdout, "ICP success rate") == 1.0


@pytest.mark.gpu
def test_benchmark_kitti_driver(tmp_path):
    """examples/benchmark_kitti.cpp over a synthetic Sequence_07 (551 small scans, 550 registrations)"""
    W.write_kitti_dataset(str(tmp_path / "kitti"))
    r = subprocess.run([os.path.join(BIN_DIR, "benchmark_kitti"), "se3_gicp", str(tmp_path / "kitti")], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "problem #549" in r.stdout
    rel_rot, rel_tra = _last_float(r.stdout, "REL rot error"), _last_float(r.stdout, "REL tra error")
    # identity would score ~0.49 deg / 0.18 m (the synthetic motion per pair); registration must be far below
    assert rel_rot < 0.15 and rel_tra < 0.04, r.stdout[-600:]


@pytest.mark.gpu
def test_benchmark_lounge_driver(tmp_path):
    """examples/benchmark_lounge.cpp (se3_gicp_with_cf, README.md:86) over a synthetic lounge_data folder"""
    W.write_lounge_dataset(str(tmp_path / "lounge"))
    r = subprocess.run([os.path.join(BIN_DIR, "benchmark_lounge"), "se3_gicp_with_cf", str(tmp_path / "lounge")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "=== Final results of algorithm: se3_gicp_with_cf ===" in r.stdout
    assert _last_float(r.stdout, "avg_angular_SO3_error") < 0.5 and _last_float(r.stdout, "avg_tra_error") < 0.03, r.stdout[-600:]
