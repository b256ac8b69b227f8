"""ctypes binding of the C ABI declared in include/se3icp.h (libse3icp_cuda.so).

The shared library is the product; this module only marshals numpy arrays across the boundary.
There is no CPU fallback: if the library is missing or no sm_100 device is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libse3icp_cuda.so")
# development hook: A/B timing of kernel variants built by profiles/experiments/build_variant.sh
if os.environ.get("SE3ICP_LIB"):
    LIB_PATH = os.path.abspath(os.environ["SE3ICP_LIB"])

PT2PT, PT2PL, GICP = 0, 1, 2
RUN_ICP, RUN_SE3_ICP, RUN_SE3_ICP_CF, RUN_SE3_PURE = 0, 1, 2, 3
NN_AUTO, NN_BRUTE_F32, NN_EXACT_F64, NN_TREE = 0, 1, 2, 3
LRF_TOLDI, LRF_SHOT = 0, 1
SOURCE, TARGET = 0, 1
STAGE_NN_SE3, STAGE_NN_XYZ, STAGE_REDUCE, STAGE_KNN_TARGET = 0, 1, 2, 3
VARIANTS = {"pt2pt": PT2PT, "pt2pl": PT2PL, "gicp": GICP}
MAX_KNN = 128
COMM_ID_BYTES = 128

STATUS = {0: "OK", 1: "ERR_ARG", 2: "ERR_NO_DEVICE", 3: "ERR_CUDA", 4: "ERR_NCCL", 5: "ERR_UNSUPPORTED", 6: "ERR_STATE"}


class Se3IcpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("se3icp status %d (%s): %s" % (code, STATUS.get(code, "?"), msg))
        self.code = code


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
        ("nn_mode", C.c_int32),
        ("use_graph", C.c_int32),
        ("record_history", C.c_int32),
        ("nn_coherence", C.c_int32),
        ("reuse_features", C.c_int32),
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
        ("time_se3_correspondence_search_ms", C.c_double),
        ("time_before_pure_icp_ms", C.c_double),
        ("exact_repairs", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("time_se3_phase_search_ms", C.c_double),
        ("feature_reuses", C.c_int64),
        ("queries_searched", C.c_int64),
        ("graph_instantiations", C.c_int64),
        ("loop_was_graph", C.c_int64),
    ]


EXPORTED_SYMBOLS = [
    "se3icp_abi_version", "se3icp_last_error", "se3icp_default_params", "se3icp_create", "se3icp_destroy",
    "se3icp_synchronize", "se3icp_set_cloud", "se3icp_set_cloud_device", "se3icp_run", "se3icp_run_async",
    "se3icp_run_finish", "se3icp_get_history", "se3icp_get_correspondences", "se3icp_get_se3_cloud",
    "se3icp_run_batch", "se3icp_run_batch_device", "se3icp_swap_clouds", "se3icp_run_sequence", "se3icp_run_sharded", "se3icp_comm_unique_id", "se3icp_comm_init",
    "se3icp_comm_destroy", "se3icp_comm_info", "se3icp_time_stage", "se3icp_knn", "se3icp_lrf", "se3icp_shot_lrf",
    "se3icp_normals", "se3icp_gicp_cov", "se3icp_nn_se3", "se3icp_nn_xyz", "se3icp_trim", "se3icp_reduce_pt2pt",
    "se3icp_reduce_pt2pl", "se3icp_reduce_gicp", "se3icp_solve",
    "se3icp_eval_error_filterreg", "se3icp_eval_corrs_with_gt", "se3icp_eval_lrf_quality", "se3icp_random_downsample",
]

_lib = None


def lib():
    """Loads libse3icp_cuda.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                "%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C se3-icp_b200/csrc)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.se3icp_last_error.restype = C.c_char_p
        _lib.se3icp_abi_version.restype = C.c_int
    return _lib


_nccl_preloaded = False


def _preload_nccl():
    """The C library resolves NCCL with dlopen("libnccl.so.2").  In a Python process that will also import torch,
    the copy bundled with torch (nvidia-nccl-cu12) must be the one that gets loaded: once the older system
    libnccl.so.2 is mapped, the loader reuses it for torch and libtorch_cuda.so fails on missing symbols."""
    global _nccl_preloaded
    if _nccl_preloaded:
        return
    _nccl_preloaded = True
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec is not None and spec.submodule_search_locations:
            path = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(path):
                C.CDLL(path, mode=C.RTLD_GLOBAL)
    except Exception:
        pass  # fall back to whatever libnccl.so.2 the loader finds


def _check(rc):
    if rc != 0:
        raise Se3IcpError(rc, lib().se3icp_last_error().decode("utf-8", "replace"))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def default_params(**kw):
    p = Params()
    lib().se3icp_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "variant" and isinstance(v, str):
            v = VARIANTS[v]
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class Context:
    """One registration context = one CUDA stream + the device memory of one source/target pair."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        _check(lib().se3icp_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.n = [0, 0]

    def close(self):
        if self._h:
            lib().se3icp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._h

    def synchronize(self):
        _check(lib().se3icp_synchronize(self._h))

    def set_cloud(self, which, xyz, append=False):
        xyz = _f64(xyz)
        assert xyz.ndim == 2 and xyz.shape[1] == 3
        _check(lib().se3icp_set_cloud(self._h, int(which), _dp(xyz), C.c_size_t(xyz.shape[0]), int(append)))
        self.n[which] = xyz.shape[0] + (self.n[which] if append else 0)

    def set_cloud_device(self, which, dev_ptr, n):
        _check(lib().se3icp_set_cloud_device(self._h, int(which), C.c_void_p(int(dev_ptr)), C.c_size_t(n)))
        self.n[which] = n

    def swap_clouds(self):
        """source <-> target together with their indices and neighbourhood features (se3icp_swap_clouds)"""
        _check(lib().se3icp_swap_clouds(self._h))
        self.n[SOURCE], self.n[TARGET] = self.n[TARGET], self.n[SOURCE]

    def run_sequence(self, scans, params, device_inputs=False):
        """Registers scans[i+1] onto scans[i] for consecutive scans, computing each scan's features once.
        scans: list of (n, 3) arrays (host) or (device_ptr, n) tuples.  Returns (T [n-1, 4, 4], [Stats])."""
        m = len(scans)
        ptr = (C.c_void_p * m)()
        cnt = (C.c_size_t * m)()
        keep = []
        for i, sc in enumerate(scans):
            if device_inputs:
                ptr[i], cnt[i] = int(sc[0]), int(sc[1])
            else:
                a = _f64(sc)
                keep.append(a)
                ptr[i], cnt[i] = a.ctypes.data, a.shape[0]
        T = np.zeros((m - 1, 4, 4))
        stats = (Stats * (m - 1))()
        _check(lib().se3icp_run_sequence(self._h, ptr, cnt, m, C.byref(params), int(bool(device_inputs)), _dp(T), stats))
        self.n[SOURCE], self.n[TARGET] = int(cnt[m - 1]), int(cnt[m - 2])
        return T, list(stats)

    def run(self, params):
        T = np.zeros((4, 4))
        st = Stats()
        _check(lib().se3icp_run(self._h, C.byref(params), _dp(T), C.byref(st)))
        return T, st

    def run_async(self, params):
        _check(lib().se3icp_run_async(self._h, C.byref(params)))

    def run_finish(self):
        T = np.zeros((4, 4))
        st = Stats()
        _check(lib().se3icp_run_finish(self._h, _dp(T), C.byref(st)))
        return T, st

    def history(self, max_entries=1024):
        buf = np.zeros((max_entries, 4, 4))
        n = C.c_int(0)
        _check(lib().se3icp_get_history(self._h, _dp(buf), int(max_entries), C.byref(n)))
        return buf[:min(n.value, max_entries)]

    def correspondences(self):
        n = self.n[SOURCE]
        idx = np.zeros(n, np.int32)
        dist = np.zeros(n)
        _check(lib().se3icp_get_correspondences(self._h, _ip(idx), _dp(dist), C.c_size_t(n)))
        return idx, dist

    def se3_cloud(self, which):
        n = self.n[which]
        fr = np.zeros((n, 4, 4))
        _check(lib().se3icp_get_se3_cloud(self._h, int(which), _dp(fr), C.c_size_t(n)))
        return fr

    # ---- one large pair sharded over ranks --------------------------------------------------------
    def comm_init(self, n_ranks, rank, comm_id):
        """comm_id: the COMM_ID_BYTES produced by comm_unique_id() on one rank and broadcast to all"""
        _preload_nccl()
        buf = (C.c_char * COMM_ID_BYTES).from_buffer_copy(bytes(comm_id))
        _check(lib().se3icp_comm_init(self._h, int(n_ranks), int(rank), buf))

    def comm_destroy(self):
        _check(lib().se3icp_comm_destroy(self._h))

    def comm_info(self):
        """(rank, n_ranks) of the communicator the library holds for this context"""
        r, n = C.c_int(0), C.c_int(0)
        _check(lib().se3icp_comm_info(self._h, C.byref(r), C.byref(n)))
        return r.value, n.value

    def run_sharded(self, params, src_begin, src_end, nccl_comm=None, rank=0, n_ranks=1):
        T = np.zeros((4, 4))
        st = Stats()
        _check(lib().se3icp_run_sharded(self._h, C.byref(params), C.c_size_t(src_begin), C.c_size_t(src_end),
                                        C.c_void_p(nccl_comm) if nccl_comm else None, int(rank), int(n_ranks), _dp(T),
                                        C.byref(st)))
        return T, st

    # ---- evaluation helpers (reference src/cc.cpp, Open3D RandomDownSample) -------------------------
    def eval_error_filterreg(self, src, T_gt, T_est):
        src, Tg, Te = _f64(src), _f64(T_gt), _f64(T_est)
        out = C.c_double(0)
        _check(lib().se3icp_eval_error_filterreg(self._h, _dp(src), C.c_size_t(src.shape[0]), _dp(Tg), _dp(Te), C.byref(out)))
        return out.value

    def eval_corrs_with_gt(self, src, tgt, T_gt):
        src, tgt, Tg = _f64(src), _f64(tgt), _f64(T_gt)
        idx = np.zeros(src.shape[0], np.int32)
        _check(lib().se3icp_eval_corrs_with_gt(self._h, _dp(src), C.c_size_t(src.shape[0]), _dp(tgt), C.c_size_t(tgt.shape[0]),
                                               _dp(Tg), _ip(idx)))
        return idx

    def eval_lrf_quality(self, src_frames, tgt_frames, T_gt, pairs):
        sf, tf, Tg = _f64(src_frames).reshape(-1, 16), _f64(tgt_frames).reshape(-1, 16), _f64(T_gt)
        pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        mean = C.c_double(0)
        per = np.zeros(pr.shape[0])
        _check(lib().se3icp_eval_lrf_quality(self._h, _dp(sf), C.c_size_t(sf.shape[0]), _dp(tf), C.c_size_t(tf.shape[0]), _dp(Tg),
                                             _ip(pr), C.c_size_t(pr.shape[0]), C.byref(mean), _dp(per)))
        return mean.value, per

    def random_downsample(self, xyz, ratio, seed=0):
        xyz = _f64(xyz)
        n = xyz.shape[0]
        out = np.zeros((int(n * ratio) + 1, 3))
        idx = np.zeros(int(n * ratio) + 1, np.int32)
        k = C.c_size_t(0)
        _check(lib().se3icp_random_downsample(self._h, _dp(xyz), C.c_size_t(n), C.c_double(ratio), C.c_uint64(seed), _dp(out),
                                              _ip(idx), C.byref(k)))
        return out[:k.value], idx[:k.value]

    def time_stage(self, stage, repeats=10):
        """average launch duration (ms) of one hot-path kernel on the data of the last run (CUDA events)"""
        ms = C.c_double(0)
        _check(lib().se3icp_time_stage(self._h, int(stage), int(repeats), C.byref(ms)))
        return ms.value

    # ---- stage-level entry points -------------------------------------------------------------
    def knn(self, xyz, k):
        xyz = _f64(xyz)
        n = xyz.shape[0]
        idx = np.zeros((n, k), np.int32)
        d2 = np.zeros((n, k))
        _check(lib().se3icp_knn(self._h, _dp(xyz), C.c_size_t(n), int(k), _ip(idx), _dp(d2)))
        return idx, d2

    def lrf(self, xyz, k):
        xyz = _f64(xyz)
        n = xyz.shape[0]
        fr = np.zeros((n, 4, 4))
        _check(lib().se3icp_lrf(self._h, _dp(xyz), C.c_size_t(n), int(k), _dp(fr)))
        return fr

    def shot_lrf(self, xyz, radius, return_unresolved=False):
        """SHOT frames with radius support (reference .cpp:121-239): n x 4x4 [x y z p]"""
        xyz = _f64(xyz)
        n = xyz.shape[0]
        fr = np.zeros((n, 4, 4))
        unresolved = C.c_int64(0)
        _check(lib().se3icp_shot_lrf(self._h, _dp(xyz), C.c_size_t(n), C.c_double(radius), _dp(fr), C.byref(unresolved)))
        return (fr, int(unresolved.value)) if return_unresolved else fr

    def normals(self, xyz, k):
        xyz = _f64(xyz)
        n = xyz.shape[0]
        out = np.zeros((n, 3))
        _check(lib().se3icp_normals(self._h, _dp(xyz), C.c_size_t(n), int(k), _dp(out)))
        return out

    def gicp_cov(self, normals, eps=1e-3):
        normals = _f64(normals)
        n = normals.shape[0]
        out = np.zeros((n, 3, 3))
        _check(lib().se3icp_gicp_cov(self._h, _dp(normals), C.c_size_t(n), C.c_double(eps), _dp(out)))
        return out

    def nn_se3(self, src_rows, tgt_rows, nn_mode=NN_AUTO):
        src_rows, tgt_rows = _f64(src_rows), _f64(tgt_rows)
        n, m = src_rows.shape[0], tgt_rows.shape[0]
        idx = np.zeros(n, np.int32)
        d2 = np.zeros(n)
        rep = C.c_int64(0)
        _check(lib().se3icp_nn_se3(self._h, _dp(src_rows), C.c_size_t(n), _dp(tgt_rows), C.c_size_t(m), int(nn_mode),
                                   _ip(idx), _dp(d2), C.byref(rep)))
        return idx, d2, rep.value

    def nn_xyz(self, queries, tgt):
        queries, tgt = _f64(queries), _f64(tgt)
        n, m = queries.shape[0], tgt.shape[0]
        idx = np.zeros(n, np.int32)
        d2 = np.zeros(n)
        _check(lib().se3icp_nn_xyz(self._h, _dp(queries), C.c_size_t(n), _dp(tgt), C.c_size_t(m), _ip(idx), _dp(d2)))
        return idx, d2

    def trim(self, dist, overlap, keep_largest=False):
        dist = np.ascontiguousarray(dist, dtype=np.float32)
        n = dist.shape[0]
        keep = np.zeros(n, np.uint8)
        nk = C.c_int64(0)
        _check(lib().se3icp_trim(self._h, dist.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(n), C.c_double(overlap),
                                 int(keep_largest), keep.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nk)))
        return nk.value, keep.astype(bool)

    def reduce_pt2pt(self, src, tgt, corr_tgt):
        src, tgt = _f64(src), _f64(tgt)
        corr = np.ascontiguousarray(corr_tgt, np.int32)
        T = np.zeros((4, 4))
        _check(lib().se3icp_reduce_pt2pt(self._h, _dp(src), C.c_size_t(src.shape[0]), _dp(tgt), C.c_size_t(tgt.shape[0]),
                                         _ip(corr), _dp(T)))
        return T

    def reduce_pt2pl(self, src, tgt, tgt_normals, corr_tgt):
        src, tgt, tgt_normals = _f64(src), _f64(tgt), _f64(tgt_normals)
        corr = np.ascontiguousarray(corr_tgt, np.int32)
        out = np.zeros(27)
        _check(lib().se3icp_reduce_pt2pl(self._h, _dp(src), C.c_size_t(src.shape[0]), _dp(tgt), _dp(tgt_normals),
                                         C.c_size_t(tgt.shape[0]), _ip(corr), _dp(out)))
        return out

    def reduce_gicp(self, src, src_cov, tgt, tgt_cov, corr_tgt, conf_src=None, conf_tgt=None):
        src, tgt, src_cov, tgt_cov = _f64(src), _f64(tgt), _f64(src_cov), _f64(tgt_cov)
        corr = np.ascontiguousarray(corr_tgt, np.int32)
        cs = _dp(_f64(conf_src)) if conf_src is not None else None
        ct = _dp(_f64(conf_tgt)) if conf_tgt is not None else None
        out = np.zeros(27)
        _check(lib().se3icp_reduce_gicp(self._h, _dp(src), _dp(src_cov), C.c_size_t(src.shape[0]), _dp(tgt), _dp(tgt_cov),
                                        C.c_size_t(tgt.shape[0]), _ip(corr), cs, ct, _dp(out)))
        return out

    def solve(self, in27):
        in27 = _f64(in27)
        T = np.zeros((4, 4))
        _check(lib().se3icp_solve(self._h, _dp(in27), _dp(T)))
        return T


def comm_unique_id():
    _preload_nccl()
    buf = (C.c_char * COMM_ID_BYTES)()
    _check(lib().se3icp_comm_unique_id(buf))
    return bytes(buf)


def run_batch(ctxs, pairs, params, device_inputs=False):
    """pairs: list of (src, tgt) numpy arrays (host) or (src_ptr, n_src, tgt_ptr, n_tgt) device tuples."""
    n_pairs = len(pairs)
    n_ctx = len(ctxs)
    handles = (C.c_void_p * n_ctx)(*[c.handle for c in ctxs])
    srcp = (C.c_void_p * n_pairs)()
    tgtp = (C.c_void_p * n_pairs)()
    ns = (C.c_size_t * n_pairs)()
    nt = (C.c_size_t * n_pairs)()
    keep = []
    for i, pr in enumerate(pairs):
        if device_inputs:
            srcp[i], ns[i], tgtp[i], nt[i] = int(pr[0]), int(pr[1]), int(pr[2]), int(pr[3])
        else:
            s, t = _f64(pr[0]), _f64(pr[1])
            keep.append((s, t))
            srcp[i], ns[i], tgtp[i], nt[i] = s.ctypes.data, s.shape[0], t.ctypes.data, t.shape[0]
    T = np.zeros((n_pairs, 4, 4))
    stats = (Stats * n_pairs)()
    fn = lib().se3icp_run_batch_device if device_inputs else lib().se3icp_run_batch
    _check(fn(handles, n_ctx, n_pairs, srcp, ns, tgtp, nt, C.byref(params), _dp(T), stats))
    return T, list(stats)
