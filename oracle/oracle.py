"""ctypes binding of the CPU oracle (oracle/libse3icp_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libse3icp_oracle.so")

PT2PT, PT2PL, GICP = 0, 1, 2
RUN_ICP, RUN_SE3_ICP, RUN_SE3_ICP_CF, RUN_SE3_PURE = 0, 1, 2, 3
VARIANTS = {"pt2pt": PT2PT, "pt2pl": PT2PL, "gicp": GICP}


class Params(C.Structure):
    _fields_ = [
        ("variant", C.c_int32),
        ("entry", C.c_int32),
        ("max_num_iterations", C.c_int32),
        ("max_num_se3_iterations", C.c_int32),
        ("number_of_nn_for_LRF", C.c_int32),
        ("knn_normals_pt2pl", C.c_int32),
        ("knn_normals_gicp", C.c_int32),
        ("trim_keep_largest", C.c_int32),
        ("mse", C.c_double),
        ("mse_switch_error", C.c_double),
        ("estimated_overlap", C.c_double),
        ("alpha_rot", C.c_double),
        ("beta_transl", C.c_double),
        ("scale_preprocessing", C.c_double),
        ("gicp_epsilon", C.c_double),
        ("lrf_method", C.c_int32),
        ("reserved0", C.c_int32),
        ("lrf_radius", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("num_iterations", C.c_int32),
        ("num_pure_se3_iterations", C.c_int32),
        ("scaling_factor", C.c_double),
        ("time_total_ms", C.c_double),
        ("time_setup_ms", C.c_double),
        ("time_corr_ms", C.c_double),
        ("time_opt_ms", C.c_double),
        ("time_before_pure_icp_ms", C.c_double),
    ]


class Trace(C.Structure):
    _fields_ = [
        ("max_iters", C.c_int32),
        ("n_iters", C.c_int32),
        ("corr_idx", C.POINTER(C.c_int32)),
        ("corr_dist", C.POINTER(C.c_float)),
        ("T_iter", C.POINTER(C.c_double)),
        ("mean_dist", C.POINTER(C.c_double)),
        ("se3_phase", C.POINTER(C.c_int32)),
        ("n_kept", C.POINTER(C.c_int32)),
    ]


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def default_params(**kw):
    p = Params()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "variant" and isinstance(v, str):
            v = VARIANTS[v]
        setattr(p, k, v)
    return p


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def run(src, tgt, params, trace_iters=0):
    """Returns (T 4x4, Stats, trace dict or None)."""
    src, tgt = _c64(src), _c64(tgt)
    n, m = src.shape[0], tgt.shape[0]
    T = np.zeros((4, 4))
    st = Stats()
    tr = None
    bufs = None
    if trace_iters > 0:
        bufs = dict(
            corr_idx=np.full((trace_iters, n), -1, np.int32),
            corr_dist=np.zeros((trace_iters, n), np.float32),
            T_iter=np.zeros((trace_iters, 4, 4)),
            mean_dist=np.zeros(trace_iters),
            se3_phase=np.zeros(trace_iters, np.int32),
            n_kept=np.zeros(trace_iters, np.int32),
        )
        tr = Trace()
        tr.max_iters = trace_iters
        tr.corr_idx = _ip(bufs["corr_idx"])
        tr.corr_dist = bufs["corr_dist"].ctypes.data_as(C.POINTER(C.c_float))
        tr.T_iter = _dp(bufs["T_iter"])
        tr.mean_dist = _dp(bufs["mean_dist"])
        tr.se3_phase = _ip(bufs["se3_phase"])
        tr.n_kept = _ip(bufs["n_kept"])
    rc = lib().orc_run(_dp(src), C.c_size_t(n), _dp(tgt), C.c_size_t(m), C.byref(params), _dp(T), C.byref(st),
                       C.byref(tr) if tr is not None else None)
    if rc != 0:
        raise RuntimeError("orc_run failed: %d" % rc)
    if bufs is not None:
        k = tr.n_iters
        bufs = {key: val[:k] for key, val in bufs.items()}
    return T, st, bufs


def knn_self(xyz, k):
    xyz = _c64(xyz)
    n = xyz.shape[0]
    idx = np.zeros((n, k), np.int32)
    d2 = np.zeros((n, k))
    lib().orc_knn_self(_dp(xyz), C.c_size_t(n), int(k), _ip(idx), _dp(d2))
    return idx, d2


def toldi(xyz, k):
    xyz = _c64(xyz)
    n = xyz.shape[0]
    fr = np.zeros((n, 4, 4))
    lib().orc_toldi(_dp(xyz), C.c_size_t(n), int(k), _dp(fr))
    return fr


def shot(xyz, radius):
    """SHOT LRF with radius support (reference .cpp:121-239): n x 4x4 frames [x y z p]"""
    xyz = _c64(xyz)
    n = xyz.shape[0]
    fr = np.zeros((n, 4, 4))
    lib().orc_shot(_dp(xyz), C.c_size_t(n), C.c_double(radius), _dp(fr))
    return fr


def normals(xyz, k):
    xyz = _c64(xyz)
    n = xyz.shape[0]
    out = np.zeros((n, 3))
    lib().orc_normals(_dp(xyz), C.c_size_t(n), int(k), _dp(out))
    return out


def gicp_cov(nrm, eps=1e-3):
    nrm = _c64(nrm)
    n = nrm.shape[0]
    out = np.zeros((n, 3, 3))
    lib().orc_gicp_cov(_dp(nrm), C.c_size_t(n), C.c_double(eps), _dp(out))
    return out


def se3_rows(frames, alpha, beta):
    frames = _c64(frames)
    n = frames.shape[0]
    out = np.zeros((n, 12))
    lib().orc_se3_rows(_dp(frames), C.c_size_t(n), C.c_double(alpha), C.c_double(beta), _dp(out))
    return out


def nn(queries, data, brute=False):
    queries, data = _c64(queries), _c64(data)
    nq, dim = queries.shape
    idx = np.zeros(nq, np.int32)
    d2 = np.zeros(nq)
    fn = lib().orc_nn_brute if brute else lib().orc_nn
    fn(_dp(queries), C.c_size_t(nq), _dp(data), C.c_size_t(data.shape[0]), int(dim), _ip(idx), _dp(d2))
    return idx, d2


def trim(dist, overlap, keep_largest=False):
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    keep = np.zeros(dist.shape[0], np.uint8)
    k = lib().orc_trim(dist.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(dist.shape[0]), C.c_double(overlap),
                       int(keep_largest), keep.ctypes.data_as(C.POINTER(C.c_uint8)))
    return k, keep.astype(bool)


def reduce_pt2pl(src, tgt, tgt_normals, cs, ct):
    src, tgt, tgt_normals = _c64(src), _c64(tgt), _c64(tgt_normals)
    cs = np.ascontiguousarray(cs, np.int32)
    ct = np.ascontiguousarray(ct, np.int32)
    out = np.zeros(27)
    lib().orc_reduce_pt2pl(_dp(src), _dp(tgt), _dp(tgt_normals), _ip(cs), _ip(ct), C.c_size_t(cs.shape[0]), _dp(out))
    return out


def reduce_gicp(src, src_cov, tgt, tgt_cov, cs, ct, weights=None):
    src, tgt, src_cov, tgt_cov = _c64(src), _c64(tgt), _c64(src_cov), _c64(tgt_cov)
    cs = np.ascontiguousarray(cs, np.int32)
    ct = np.ascontiguousarray(ct, np.int32)
    w = _dp(_c64(weights)) if weights is not None else None
    out = np.zeros(27)
    lib().orc_reduce_gicp(_dp(src), _dp(src_cov), _dp(tgt), _dp(tgt_cov), _ip(cs), _ip(ct), w,
                          C.c_size_t(cs.shape[0]), _dp(out))
    return out


def solve6(in27):
    in27 = _c64(in27)
    T = np.zeros((4, 4))
    lib().orc_solve6(_dp(in27), _dp(T))
    return T


def umeyama(src, tgt, cs, ct):
    src, tgt = _c64(src), _c64(tgt)
    cs = np.ascontiguousarray(cs, np.int32)
    ct = np.ascontiguousarray(ct, np.int32)
    T = np.zeros((4, 4))
    lib().orc_umeyama(_dp(src), _dp(tgt), _ip(cs), _ip(ct), C.c_size_t(cs.shape[0]), _dp(T))
    return T


def eig3(A):
    A = _c64(A)
    ev = np.zeros(3)
    V = np.zeros((3, 3))
    lib().orc_eig3(_dp(A), _dp(ev), _dp(V))
    return ev, V


# --------------------------------------------------------------------------------------------------
# reference src/cc.cpp evaluation helpers, restated in numpy (test infrastructure for se3icp_eval_*)
# --------------------------------------------------------------------------------------------------
def cc_error_filterreg(src, T_gt, T_est):
    """cc::error_filterreg, src/cc.cpp:4-20: mean || T_gt p - T_est p ||"""
    src = np.asarray(src, dtype=np.float64)
    a = src @ T_gt[:3, :3].T + T_gt[:3, 3]
    b = src @ T_est[:3, :3].T + T_est[:3, 3]
    return float(np.linalg.norm(a - b, axis=1).sum() / len(src))


def cc_angular_error_alt(R1, R2):
    """cc::angularErrorSO3_alt with safe_acos, src/cc.cpp:39-61, in degrees"""
    a = (np.trace(R1.T @ R2) - 1.0) / 2.0
    ang = np.pi if a <= -1.0 else (0.0 if a >= 1.0 else np.arccos(a))
    return abs(ang) * (180.0 / np.pi)


def cc_lrf_quality(src_frames, tgt_frames, T_gt, pairs):
    """cc::evaluate_LRF_quality, src/cc.cpp:63-88: (mean, per-pair) angular error between T_gt * source frame and target frame"""
    err = np.array([cc_angular_error_alt((T_gt @ src_frames[i])[:3, :3], tgt_frames[j][:3, :3]) for i, j in pairs])
    return float(err.sum() / len(pairs)), err


def cc_corrs_with_gt(src, tgt, T_gt):
    """cc::compute_corrs_with_gt, src/cc.cpp:116-143: nearest target point of every T_gt-moved source point (exact;
    ties to the smaller index like the oracle's kd-tree)"""
    moved = np.asarray(src, dtype=np.float64) @ T_gt[:3, :3].T + T_gt[:3, 3]
    idx, _ = nn(moved, np.asarray(tgt, dtype=np.float64))
    return idx
