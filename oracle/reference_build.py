"""ctypes binding of oracle/_ref/libse3icp_reference.so — TEST INFRASTRUCTURE ONLY.

The library is the reference's own src/iterative_SE3_registration.cpp (compiled unmodified from /root/reference by
oracle/Makefile) behind the C shim oracle/ref_shim.cpp; Open3D / PCL / Eigen are replaced by compat/ + oracle/refdeps/.
It exists to pin oracle/se3icp_oracle.cpp against the reference's own control flow.  Nothing in the product imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libse3icp_reference.so")
REFERENCE_SRC = "/root/reference/src/iterative_SE3_registration.cpp"
ENTRIES = {"run_icp": 0, "run_se3_icp": 1, "run_se3_icp_with_cf": 2, "run_se3_pure": 3}


class RefParams(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int),
        ("max_num_se3_iterations", C.c_int),
        ("number_of_nn_for_LRF", C.c_int),
        ("trim_keep_largest", C.c_int),
        ("mse", C.c_double),
        ("mse_switch_error", C.c_double),
        ("estimated_overlap", C.c_double),
        ("alpha_rot", C.c_double),
        ("beta_transl", C.c_double),
        ("scale_preprocessing", C.c_double),
    ]


_lib = None


def available():
    """True when the library exists (it travels to the GPU box) or can be built here (reference tree present)."""
    return os.path.exists(LIB_PATH) or os.path.exists(REFERENCE_SRC)


def build():
    if not os.path.exists(REFERENCE_SRC):
        return os.path.exists(LIB_PATH)
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True, env=env)
    return True


def lib():
    global _lib
    if _lib is None:
        if not build():
            raise RuntimeError("oracle/_ref/libse3icp_reference.so is missing and /root/reference is not present")
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def default_params(**kw):
    p = RefParams()
    lib().ref_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def run(entry, variant, src, tgt, params=None, want_corr=False):
    """Runs the reference class; returns (T 4x4, num_iterations, num_pure_se3_iterations[, corr])."""
    src, tgt = _c64(src), _c64(tgt)
    p = params if params is not None else default_params()
    T = np.zeros((4, 4))
    it = (C.c_int * 2)()
    corr = np.zeros(len(src), np.int32) if want_corr else None
    rc = lib().ref_run(ENTRIES[entry] if isinstance(entry, str) else int(entry), variant.encode(), _dp(src),
                       C.c_size_t(len(src)), _dp(tgt), C.c_size_t(len(tgt)), C.byref(p), _dp(T), it,
                       corr.ctypes.data_as(C.POINTER(C.c_int)) if want_corr else None)
    if rc != 0:
        raise RuntimeError("ref_run failed: %d" % rc)
    out = (T, it[0], it[1])
    return out + (corr,) if want_corr else out


def toldi(xyz, knn):
    xyz = _c64(xyz)
    fr = np.zeros((len(xyz), 4, 4))
    lib().ref_toldi(_dp(xyz), C.c_size_t(len(xyz)), int(knn), _dp(fr))
    return fr


def shot(xyz, radius):
    """computeAllSHOTSE3FramesOMP of the reference source (.cpp:226-239)"""
    xyz = _c64(xyz)
    fr = np.zeros((len(xyz), 4, 4))
    lib().ref_shot(_dp(xyz), C.c_size_t(len(xyz)), C.c_double(radius), _dp(fr))
    return fr


def gicp_cov(xyz, eps=1e-3):
    xyz = _c64(xyz)
    nrm = np.zeros((len(xyz), 3))
    cov = np.zeros((len(xyz), 3, 3))
    lib().ref_gicp_cov(_dp(xyz), C.c_size_t(len(xyz)), C.c_double(eps), _dp(nrm), _dp(cov))
    return nrm, cov


def nn_se3(src_frames, tgt_frames):
    s, t = _c64(src_frames), _c64(tgt_frames)
    idx = np.zeros(len(s), np.int32)
    dist = np.zeros(len(s))
    lib().ref_nn_se3(_dp(s), C.c_size_t(len(s)), _dp(t), C.c_size_t(len(t)), idx.ctypes.data_as(C.POINTER(C.c_int)), _dp(dist))
    return idx, dist
