"""Generates the committed input fixtures from the reference's bundled data files.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_inputs.py

Outputs (tests/golden/):
    c1_source.npy, c1_target.npy   created_example_reg_problem/{source,target}.ply as float64 [4167,3]
    c1_T_gt.npy                    exact ground truth rot_3d(pi/9, pi/8, -pi/7), t=(1,2,3)
                                   (reference examples/create_and_save_reg_problem.cpp:31-37, src/cc.cpp:22-30)
    bunny_unique_f32.npy           unique vertices of stanford_bunny.ply, float32 [34834,3], first-occurrence order
Only DATA is copied (point coordinates); no reference source code.
"""
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_ply_xyz(path):
    """Minimal binary-little-endian PLY reader: vertex element with leading x,y,z (float or double)."""
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply"
        fmt = None
        n_vertex = 0
        props = []
        in_vertex = False
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("format"):
                fmt = line.split()[1]
            elif line.startswith("element"):
                _, name, cnt = line.split()
                in_vertex = name == "vertex"
                if in_vertex:
                    n_vertex = int(cnt)
            elif line.startswith("property") and in_vertex:
                parts = line.split()
                props.append((parts[-1], parts[1]))
            elif line == "end_header":
                break
        assert fmt == "binary_little_endian"
        tmap = {"float": "<f4", "double": "<f8", "uchar": "u1", "int": "<i4", "float32": "<f4", "float64": "<f8"}
        dt = np.dtype([(n, tmap[t]) for n, t in props])
        v = np.frombuffer(f.read(n_vertex * dt.itemsize), dtype=dt, count=n_vertex)
    return np.stack([v["x"], v["y"], v["z"]], axis=1), dt


def rot_3d(roll, pitch, yaw):
    """cc::rot_3d: q = yaw(Z) * pitch(Y) * roll(X)."""
    cx, sx, cy, sy, cz, sz = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def main():
    src, _ = read_ply_xyz(os.path.join(REF, "created_example_reg_problem/source.ply"))
    tgt, _ = read_ply_xyz(os.path.join(REF, "created_example_reg_problem/target.ply"))
    src = src.astype(np.float64)
    tgt = tgt.astype(np.float64)
    T = np.eye(4)
    T[:3, :3] = rot_3d(np.pi / 9, np.pi / 8, -np.pi / 7)
    T[:3, 3] = [1.0, 2.0, 3.0]
    err = np.abs(src @ T[:3, :3].T + T[:3, 3] - tgt).max()
    txt = np.loadtxt(os.path.join(REF, "created_example_reg_problem/transformation_gt.txt"))
    print("C1: %d / %d points, |T_gt*src - tgt|_max = %.3e, |T_gt - txt|_max = %.3e" %
          (len(src), len(tgt), err, np.abs(T - txt).max()))
    assert err < 1e-12 and np.abs(T - txt).max() < 1e-6
    np.save(os.path.join(OUT, "c1_source.npy"), src)
    np.save(os.path.join(OUT, "c1_target.npy"), tgt)
    np.save(os.path.join(OUT, "c1_T_gt.npy"), T)

    bunny, _ = read_ply_xyz(os.path.join(REF, "stanford_bunny.ply"))
    _, first = np.unique(bunny, axis=0, return_index=True)
    uniq = bunny[np.sort(first)].astype(np.float32)
    print("bunny: %d vertices, %d unique" % (len(bunny), len(uniq)))
    np.save(os.path.join(OUT, "bunny_unique_f32.npy"), uniq)


if __name__ == "__main__":
    sys.exit(main())
