"""Synthetic workloads for tests and bench.py (the reference's gdrive datasets are not available offline).

Shapes follow BASELINE.json `configs` / SURVEY.md §8(d):
  C1  bundled fixture (tests/golden/c1_*.npy)
  C2  stanford bunny x50 with Gaussian noise and easy/moderate/difficult ground truth
      (reference examples/benchmark_synthetic.cpp:13-56,91-116)
  C3  KITTI-like spinning-LiDAR scan pairs, ~120 k points (reference examples/benchmark_kitti.cpp:120-148)
  C4  lounge-like RGB-D frame pairs, ~300 k points (reference examples/benchmark_lounge.cpp:154-186)
Nothing here is product code.
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rot_3d(roll, pitch, yaw):
    """reference src/cc.cpp:22-30: q = yaw(Z) * pitch(Y) * roll(X)."""
    cx, sx, cy, sy, cz, sz = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def make_T(R, t):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def apply_T(T, pts):
    return pts @ T[:3, :3].T + T[:3, 3]


def rotation_error(A, B):
    """reference src/cc.cpp:49-61 angularErrorSO3_alt, in radians."""
    c = (np.trace(A[:3, :3].T @ B[:3, :3]) - 1.0) / 2.0
    return float(np.arccos(np.clip(c, -1.0, 1.0)))


def load_c1():
    return (np.load(os.path.join(GOLDEN, "c1_source.npy")), np.load(os.path.join(GOLDEN, "c1_target.npy")),
            np.load(os.path.join(GOLDEN, "c1_T_gt.npy")))


def load_bunny(scale=50.0):
    """unique vertices of stanford_bunny.ply (34 834), scaled by 50 as benchmark_synthetic.cpp:95."""
    return np.load(os.path.join(GOLDEN, "bunny_unique_f32.npy")).astype(np.float64) * scale


LEVELS = {"easy": (np.pi / 4, 5.0), "moderate": (np.pi / 2, 10.0), "difficult": (np.pi, 15.0)}


def bunny_problem(level="easy", seed=1, n_points=None, noise=0.005):
    """benchmark_synthetic.cpp:104-116,150-160: GT angles U(+-a), translation U(+-t); independent
    N(0, noise*I) on source and target.  n_points: random down-sample (the reference uses 2 %)."""
    rng = np.random.default_rng(seed)
    pts = load_bunny()
    if n_points is not None and n_points < len(pts):
        pts = pts[np.sort(rng.choice(len(pts), n_points, replace=False))]
    a, t = LEVELS[level]
    T = make_T(rot_3d(*rng.uniform(-a, a, 3)), rng.uniform(-t, t, 3))
    sd = np.sqrt(noise)
    src = pts + (rng.normal(0, sd, pts.shape) if noise > 0 else 0)
    tgt = apply_T(T, pts) + (rng.normal(0, sd, pts.shape) if noise > 0 else 0)
    return src, tgt, T


# --------------------------------------------------------------------------------------------------
# KITTI-like spinning LiDAR
# --------------------------------------------------------------------------------------------------
def _street_scene(rng):
    """ground plane + boxes (buildings, cars) + vertical cylinders (poles, trunks) along a street."""
    boxes = []
    for side in (-1.0, 1.0):
        x = -60.0
        while x < 90.0:
            w, d, h = rng.uniform(6, 18), rng.uniform(6, 14), rng.uniform(4, 14)
            off = side * rng.uniform(9, 16)
            y0, y1 = (off, off + side * d) if side > 0 else (off + side * d, off)
            boxes.append((x, min(y0, y1), -1.73, x + w, max(y0, y1), -1.73 + h))
            x += w + rng.uniform(1, 8)
    for _ in range(14):  # parked cars
        cx, side = rng.uniform(-50, 80), rng.choice([-1.0, 1.0])
        cy = side * rng.uniform(3.5, 6.5)
        boxes.append((cx, cy - 0.9, -1.73, cx + rng.uniform(3.8, 4.8), cy + 0.9, -1.73 + rng.uniform(1.3, 1.9)))
    cyl = []
    for _ in range(30):
        cyl.append((rng.uniform(-55, 85), rng.choice([-1.0, 1.0]) * rng.uniform(6.5, 8.5), rng.uniform(0.08, 0.35),
                    -1.73 + rng.uniform(3, 8)))
    return np.array(boxes), np.array(cyl)


def _ray_cast(origin, dirs, boxes, cyl, max_range):
    """nearest hit distance for rays origin + s*dirs (world frame); inf where nothing within max_range."""
    n = dirs.shape[0]
    best = np.full(n, np.inf)
    # ground z = -1.73
    dz = dirs[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        s = (-1.73 - origin[2]) / dz
    ok = (dz < 0) & (s > 0)
    best = np.where(ok, np.minimum(best, s), best)
    # axis-aligned boxes, slab method
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dirs
    for b in boxes:
        t0 = (b[:3] - origin) * inv
        t1 = (b[3:] - origin) * inv
        tmin = np.nanmax(np.minimum(t0, t1), axis=1)
        tmax = np.nanmin(np.maximum(t0, t1), axis=1)
        hit = (tmax >= tmin) & (tmax > 0)
        s = np.where(tmin > 0, tmin, tmax)
        best = np.where(hit & (s < best), s, best)
    # vertical cylinders (x0, y0, r, ztop), from the ground up
    dxy = dirs[:, :2]
    a = (dxy ** 2).sum(1)
    for c in cyl:
        oc = origin[:2] - c[:2]
        bq = 2 * (dxy @ oc)
        cq = oc @ oc - c[2] ** 2
        disc = bq * bq - 4 * a * cq
        with np.errstate(divide="ignore", invalid="ignore"):
            s = (-bq - np.sqrt(np.maximum(disc, 0))) / (2 * a)
        z = origin[2] + s * dirs[:, 2]
        hit = (disc > 0) & (s > 0) & (z < c[3]) & (z > -1.73)
        best = np.where(hit & (s < best), s, best)
    best[best > max_range] = np.inf
    return best


def _lidar_scan(pose, boxes, cyl, rng, n_rings, n_az, max_range, sigma):
    elev = np.deg2rad(np.linspace(-24.8, 2.0, n_rings))
    az = np.linspace(0, 2 * np.pi, n_az, endpoint=False) + rng.uniform(0, 2 * np.pi / n_az)
    ce, se_ = np.cos(elev)[:, None], np.sin(elev)[:, None]
    d_s = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se_, (n_rings, n_az))], -1)
    d_s = d_s.reshape(-1, 3)
    d_w = d_s @ pose[:3, :3].T
    rng_hit = _ray_cast(pose[:3, 3], d_w, boxes, cyl, max_range)
    ok = np.isfinite(rng_hit)
    r = rng_hit[ok] + rng.normal(0, sigma, ok.sum())
    return d_s[ok] * r[:, None]  # points in the sensor frame


def lidar_pair(seed=0, n_rings=64, n_az=1900, max_range=80.0, sigma=0.02):
    """Two consecutive scans of a static street scene.  Returns (source, target, T_gt) with
    T_gt * source ~ target, i.e. scan i+1 registered onto scan i as benchmark_kitti.cpp:130-131."""
    rng = np.random.default_rng(1000 + seed)
    boxes, cyl = _street_scene(rng)
    pose0 = make_T(rot_3d(0, 0, rng.uniform(-0.05, 0.05)), [rng.uniform(-5, 5), rng.uniform(-1, 1), 0.0])
    step = make_T(rot_3d(rng.uniform(-0.005, 0.005), rng.uniform(-0.005, 0.005), np.deg2rad(rng.uniform(-3, 3))),
                  [rng.uniform(1.0, 1.5), rng.uniform(-0.05, 0.05), rng.uniform(-0.02, 0.02)])
    pose1 = pose0 @ step
    tgt = _lidar_scan(pose0, boxes, cyl, rng, n_rings, n_az, max_range, sigma)
    src = _lidar_scan(pose1, boxes, cyl, rng, n_rings, n_az, max_range, sigma)
    return src, tgt, step  # p_0 = step * p_1


def lidar_sequence(seed=0, n_scans=4, n_rings=64, n_az=1900, max_range=80.0, sigma=0.02):
    """n_scans consecutive scans of one static street scene along a gently curving drive (the shape of KITTI
    odometry: benchmark_kitti.cpp:120-131 registers scan i+1 onto scan i).  Returns (scans, steps) with
    steps[i] * scans[i+1] ~ scans[i]."""
    rng = np.random.default_rng(4000 + seed)
    boxes, cyl = _street_scene(rng)
    pose = make_T(rot_3d(0, 0, rng.uniform(-0.05, 0.05)), [rng.uniform(-20, -10), rng.uniform(-1, 1), 0.0])
    scans, steps = [], []
    for k in range(n_scans):
        scans.append(_lidar_scan(pose, boxes, cyl, rng, n_rings, n_az, max_range, sigma))
        step = make_T(rot_3d(rng.uniform(-0.005, 0.005), rng.uniform(-0.005, 0.005), np.deg2rad(rng.uniform(-3, 3))),
                      [rng.uniform(1.0, 1.5), rng.uniform(-0.05, 0.05), rng.uniform(-0.02, 0.02)])
        if k + 1 < n_scans:
            steps.append(step)
        pose = pose @ step
    return scans, steps


KITTI_PARAMS = dict(estimated_overlap=0.7, mse=1e-7, mse_switch_error=5e-7, max_num_se3_iterations=10,
                    number_of_nn_for_LRF=90, alpha_rot=3.0)  # benchmark_kitti.cpp:133-148


# --------------------------------------------------------------------------------------------------
# lounge-like RGB-D
# --------------------------------------------------------------------------------------------------
def rgbd_pair(seed=0, width=640, height=480, f=525.0, stride=1):
    """Depth images of a box room with furniture boxes from two nearby camera poses (depth 0.4-4 m,
    depth-dependent noise following the model in reference .cpp:25-27).  Returns (source, target, T_gt)."""
    rng = np.random.default_rng(2000 + seed)
    room = np.array([[-2.5, -1.4, -0.5, 2.5, 1.4, 3.7]])
    boxes = []
    for _ in range(10):
        cx, cz = rng.uniform(-2.0, 2.0), rng.uniform(1.2, 3.2)
        w, h, d = rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.2), rng.uniform(0.3, 0.9)
        boxes.append((cx - w / 2, 1.4 - h, cz - d / 2, cx + w / 2, 1.4, cz + d / 2))  # standing on the floor (y down)
    boxes = np.array(boxes)
    u, v = np.meshgrid(np.arange(0, width, stride), np.arange(0, height, stride))
    d_c = np.stack([(u - width / 2 + 0.5) / f, (v - height / 2 + 0.5) / f, np.ones_like(u, dtype=float)], -1).reshape(-1, 3)

    def cast(pose):
        o = pose[:3, 3]
        d_w = d_c @ pose[:3, :3].T
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d_w
        # inside the room: exit distance of the room box
        t0 = (room[0, :3] - o) * inv
        t1 = (room[0, 3:] - o) * inv
        best = np.nanmin(np.maximum(t0, t1), axis=1)
        for b in boxes:
            t0 = (b[:3] - o) * inv
            t1 = (b[3:] - o) * inv
            tmin = np.nanmax(np.minimum(t0, t1), axis=1)
            tmax = np.nanmin(np.maximum(t0, t1), axis=1)
            hit = (tmax >= tmin) & (tmin > 0)
            best = np.where(hit & (tmin < best), tmin, best)
        z = best  # d_c has unit z, so the ray parameter is the depth
        sig = 0.002203 * z * z - 0.001028 * z + 0.0005351
        z = z + rng.normal(0, 1, z.shape) * sig
        ok = (z > 0.4) & (z < 4.0)
        return d_c[ok] * z[ok, None]

    pose0 = make_T(rot_3d(rng.uniform(-0.05, 0.05), rng.uniform(-0.2, 0.2), rng.uniform(-0.03, 0.03)),
                   [rng.uniform(-0.5, 0.5), rng.uniform(-0.2, 0.2), rng.uniform(-0.3, 0.2)])
    step = make_T(rot_3d(rng.uniform(-0.03, 0.03), rng.uniform(-0.08, 0.08), rng.uniform(-0.03, 0.03)),
                  rng.uniform(-0.08, 0.08, 3))
    pose1 = pose0 @ step
    tgt = cast(pose0)
    src = cast(pose1)
    return src, tgt, step


def rgbd_pair_device(seed=0, width=4096, height=3072, f=525.0 * 6.4, device="cuda"):
    """rgbd_pair() rendered on the GPU with torch, for the one configuration whose host-side generation takes a
    minute (BASELINE.json configs[4]: ~10 M points per cloud).  Same scene, poses and noise model; the noise samples
    come from torch's generator, so the clouds are not those of rgbd_pair().  Returns (src, tgt, T_gt) with the
    clouds as contiguous (n, 3) float64 CUDA tensors."""
    import torch
    rng = np.random.default_rng(2000 + seed)
    room = np.array([-2.5, -1.4, -0.5, 2.5, 1.4, 3.7])
    boxes = []
    for _ in range(10):
        cx, cz = rng.uniform(-2.0, 2.0), rng.uniform(1.2, 3.2)
        w, h, d = rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.2), rng.uniform(0.3, 0.9)
        boxes.append((cx - w / 2, 1.4 - h, cz - d / 2, cx + w / 2, 1.4, cz + d / 2))
    pose0 = make_T(rot_3d(rng.uniform(-0.05, 0.05), rng.uniform(-0.2, 0.2), rng.uniform(-0.03, 0.03)),
                   [rng.uniform(-0.5, 0.5), rng.uniform(-0.2, 0.2), rng.uniform(-0.3, 0.2)])
    step = make_T(rot_3d(rng.uniform(-0.03, 0.03), rng.uniform(-0.08, 0.08), rng.uniform(-0.03, 0.03)),
                  rng.uniform(-0.08, 0.08, 3))
    gen = torch.Generator(device=device)
    gen.manual_seed(2000 + seed)
    f64 = dict(dtype=torch.float64, device=device)
    u = (torch.arange(width, **f64) - width / 2 + 0.5) / f
    v = (torch.arange(height, **f64) - height / 2 + 0.5) / f
    d_c = torch.stack([u[None, :].expand(height, width), v[:, None].expand(height, width),
                       torch.ones(height, width, **f64)], -1).reshape(-1, 3)

    def slab(box, o, inv):
        lo = (torch.tensor(box[:3], **f64) - o) * inv
        hi = (torch.tensor(box[3:], **f64) - o) * inv
        return torch.minimum(lo, hi).amax(1), torch.maximum(lo, hi).amin(1)

    def cast(pose):
        o = torch.tensor(pose[:3, 3], **f64)
        inv = 1.0 / (d_c @ torch.tensor(pose[:3, :3].T.copy(), **f64))
        _, best = slab(room, o, inv)  # inside the room: exit distance of the room box
        for b in boxes:
            tmin, tmax = slab(b, o, inv)
            hit = (tmax >= tmin) & (tmin > 0) & (tmin < best)
            best = torch.where(hit, tmin, best)
        z = best
        z = z + torch.randn(z.shape, generator=gen, **f64) * (0.002203 * z * z - 0.001028 * z + 0.0005351)
        ok = (z > 0.4) & (z < 4.0)
        return (d_c[ok] * z[ok, None]).contiguous()

    tgt = cast(pose0)
    src = cast(pose0 @ step)
    return src, tgt, step


LOUNGE_PARAMS = dict(estimated_overlap=0.75, mse_switch_error=5e-5, max_num_se3_iterations=10,
                     number_of_nn_for_LRF=90)  # benchmark_lounge.cpp:183-186


# --------------------------------------------------------------------------------------------------
# on-disk datasets in the layouts the reference's benchmark drivers read (tests of the unchanged drivers)
# --------------------------------------------------------------------------------------------------
def write_ply(path, pts, dtype="<f8"):
    """binary little-endian PLY with x/y/z (double by default, as Open3D writes the bundled fixture)"""
    name = {"<f8": "double", "<f4": "float"}[dtype]
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\ncomment Created by Open3D\nelement vertex %d\n"
                 "property %s x\nproperty %s y\nproperty %s z\nend_header\n" % (len(pts), name, name, name)).encode())
        f.write(np.ascontiguousarray(pts, dtype=dtype).tobytes())


def read_ply_xyz(path):
    """reads back what write_ply wrote"""
    with open(path, "rb") as f:
        n, dt = 0, "<f8"
        while True:
            line = f.readline().decode().strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line.startswith("property float"):
                dt = "<f4"
            if line == "end_header":
                break
        return np.frombuffer(f.read(n * 3 * np.dtype(dt).itemsize), dtype=dt).reshape(n, 3).astype(np.float64)


def _row12(T):
    return " ".join("%.12e" % v for v in T[:3].reshape(-1))


def write_synthetic_dataset(folder, problems):
    """examples/benchmark_synthetic.cpp:300-345: <folder>/gt_data (12 numbers per line), source<i>.ply, target<i>.ply"""
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, "gt_data"), "w") as f:
        for i, (src, tgt, T) in enumerate(problems):
            f.write(_row12(T) + "\n")
            write_ply(os.path.join(folder, "source%d.ply" % i), src)
            write_ply(os.path.join(folder, "target%d.ply" % i), tgt)


def write_kitti_dataset(folder, seed=0, n_rings=16, n_az=300):
    """examples/benchmark_kitti.cpp:70-108: <folder>/Sequence_07/07.txt (poses on every other line) and
    Sequence_07/Downsampled/000000.ply ... 001100.ply (every 2nd scan, 551 files).  Small synthetic scans of one
    static street scene along a gently curving trajectory.  Returns the list of poses."""
    rng = np.random.default_rng(3000 + seed)
    boxes, cyl = _street_scene(rng)
    seq = os.path.join(folder, "Sequence_07")
    os.makedirs(os.path.join(seq, "Downsampled"), exist_ok=True)
    poses = []
    pose = make_T(rot_3d(0, 0, 0), [-50.0, 0.0, 0.0])
    with open(os.path.join(seq, "07.txt"), "w") as f:
        for k in range(551):
            poses.append(pose.copy())
            f.write(_row12(pose) + "\n")
            f.write("unused line (the driver reads every other line)\n")
            pts = _lidar_scan(pose, boxes, cyl, rng, n_rings, n_az, 80.0, 0.01)
            write_ply(os.path.join(seq, "Downsampled", "%06d.ply" % (2 * k)), pts)
            step = make_T(rot_3d(0, 0, np.deg2rad(rng.uniform(-1.0, 1.0))), [rng.uniform(0.15, 0.22), 0.0, 0.0])
            pose = pose @ step
    return poses


def _rgbd_scene(rng):
    room = np.array([[-2.5, -1.4, -0.5, 2.5, 1.4, 3.7]])
    boxes = []
    for _ in range(10):
        cx, cz = rng.uniform(-2.0, 2.0), rng.uniform(1.2, 3.2)
        w, h, d = rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.2), rng.uniform(0.3, 0.9)
        boxes.append((cx - w / 2, 1.4 - h, cz - d / 2, cx + w / 2, 1.4, cz + d / 2))
    return room, np.array(boxes)


def _rgbd_render(pose, room, boxes, rng, width=640, height=480, f=525.0, stride=8):
    u, v = np.meshgrid(np.arange(0, width, stride), np.arange(0, height, stride))
    d_c = np.stack([(u - width / 2 + 0.5) / f, (v - height / 2 + 0.5) / f, np.ones_like(u, dtype=float)], -1).reshape(-1, 3)
    o = pose[:3, 3]
    d_w = d_c @ pose[:3, :3].T
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d_w
    t0 = (room[0, :3] - o) * inv
    t1 = (room[0, 3:] - o) * inv
    best = np.nanmin(np.maximum(t0, t1), axis=1)
    for b in boxes:
        t0 = (b[:3] - o) * inv
        t1 = (b[3:] - o) * inv
        tmin = np.nanmax(np.minimum(t0, t1), axis=1)
        tmax = np.nanmin(np.maximum(t0, t1), axis=1)
        hit = (tmax >= tmin) & (tmin > 0)
        best = np.where(hit & (tmin < best), tmin, best)
    z = best
    sig = 0.002203 * z * z - 0.001028 * z + 0.0005351
    z = z + rng.normal(0, 1, z.shape) * sig
    ok = (z > 0.4) & (z < 4.0)
    return d_c[ok] * z[ok, None]


def write_lounge_dataset(folder, seed=0, stride=8):
    """examples/benchmark_lounge.cpp:142-175: <folder>/lounge_data/lounge_trajectory.log (redwood format: a line
    with three integers, then the 4x4 camera-to-world matrix) and 00000i.ply for i = 1, 6, ..., 396."""
    rng = np.random.default_rng(4000 + seed)
    room, boxes = _rgbd_scene(rng)
    d = os.path.join(folder, "lounge_data")
    os.makedirs(d, exist_ok=True)
    poses = []
    pose = make_T(rot_3d(0.0, -0.25, 0.0), [-0.8, 0.0, -0.2])
    with open(os.path.join(d, "lounge_trajectory.log"), "w") as f:
        for k in range(401):  # frame k+1 of the driver's numbering
            poses.append(pose.copy())
            f.write("%d %d %d\n" % (k, k, k + 1))
            for r in range(4):
                f.write(" ".join("%.10f" % v for v in pose[r]) + "\n")
            if k % 5 == 0 and k <= 395:
                write_ply(os.path.join(d, "%06d.ply" % (k + 1)), _rgbd_render(pose, room, boxes, rng, stride=stride))
            step = make_T(rot_3d(rng.uniform(-0.002, 0.002), 0.0025 + rng.uniform(-0.001, 0.001), rng.uniform(-0.002, 0.002)),
                          [0.004, rng.uniform(-0.001, 0.001), 0.001])
            pose = pose @ step
    return poses
