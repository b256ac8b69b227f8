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
