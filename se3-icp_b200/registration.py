"""Python mirror of the reference class IterativeSE3Registration
(reference include/iterative_SE3_registration.hpp:27-99, src/iterative_SE3_registration.cpp:334-1128).

Same method names, public field names, defaults and error behaviour as the C++ class, so scripts and
tests read like the reference's drivers (examples/run_registration_method.cpp:35-60).  All numeric
work happens in libse3icp_cuda.so through the C ABI; nothing here computes on the CPU.
"""
import sys

import numpy as np

from . import capi


class IterativeSE3Registration:
    def __init__(self, device=0, stream=None):
        # defaults: reference .cpp:334-348
        self.max_num_iterations_ = 150
        self.max_num_se3_iterations_ = 20
        self.num_iterations_ = 0
        self.num_pure_se3_iterations_ = -1
        self.mse_ = 0.00001
        self.lrf_radius_ = 0.8  # SHOT frame radius; used only when use_shot_lrf_ is set (reference: dormant, .cpp:593-594)
        self.mse_switch_error_ = 0.001
        self.number_of_nn_for_LRF_ = 30
        self.estimated_overlap_ = 1.0
        self.alpha_rot = 3.0
        self.beta_transl = 1.0
        self.scale_preprocessing = 3.0
        self.time_before_pure_icp_ = 0.0
        self.time_se3_correspondence_search_ = 0.0
        self.current_estimated_T_ = np.eye(4)
        self.estimated_history_ = []
        # extensions (not in the reference): comparator direction of the trimmed rejector and NN strategy
        self.trim_keep_largest_ = True  # PCL 1.14's comparator (include/se3icp.h)
        self.nn_mode_ = capi.NN_AUTO
        self.use_shot_lrf_ = False  # True: SHOT frames (reference .cpp:121-239) instead of TOLDI
        self._source = np.zeros((0, 3))
        self._target = np.zeros((0, 3))
        self._ctx = capi.Context(device, stream)
        self.last_stats = None

    # reference .cpp:358-366 / :372-376: the cloud overloads append
    def setSourceCloud(self, cloud):
        self._source = np.concatenate([self._source, np.asarray(cloud, dtype=np.float64).reshape(-1, 3)])

    def setTargetCloud(self, cloud):
        self._target = np.concatenate([self._target, np.asarray(cloud, dtype=np.float64).reshape(-1, 3)])

    def _params(self, entry, variant):
        return capi.default_params(
            variant=variant, entry=entry, max_num_iterations=self.max_num_iterations_,
            max_num_se3_iterations=self.max_num_se3_iterations_, number_of_nn_for_LRF=self.number_of_nn_for_LRF_,
            trim_keep_largest=int(self.trim_keep_largest_), mse=self.mse_, mse_switch_error=self.mse_switch_error_,
            estimated_overlap=self.estimated_overlap_, alpha_rot=self.alpha_rot, beta_transl=self.beta_transl,
            scale_preprocessing=self.scale_preprocessing, nn_mode=self.nn_mode_, record_history=int(entry == capi.RUN_ICP),
            lrf_method=int(self.use_shot_lrf_), lrf_radius=self.lrf_radius_)

    def _run(self, entry, variant_name):
        if variant_name not in capi.VARIANTS:
            return self._invalid_variant(entry)
        self._ctx.set_cloud(capi.SOURCE, self._source)
        self._ctx.set_cloud(capi.TARGET, self._target)
        T, st = self._ctx.run(self._params(entry, capi.VARIANTS[variant_name]))
        self.current_estimated_T_ = T
        self.num_iterations_ = st.num_iterations
        if entry != capi.RUN_ICP:
            self.num_pure_se3_iterations_ = st.num_pure_se3_iterations
            self.time_se3_correspondence_search_ = st.time_se3_correspondence_search_ms if entry == capi.RUN_SE3_ICP_CF else 0.0
        if entry == capi.RUN_SE3_ICP_CF:
            self.time_before_pure_icp_ = st.time_before_pure_icp_ms
        if entry == capi.RUN_ICP:  # .cpp:491,538
            self.estimated_history_.append(np.eye(4))
            self.estimated_history_.extend(list(self._ctx.history(max(self.max_num_iterations_, 1))))
        self.last_stats = st
        return T

    def _invalid_variant(self, entry):
        # reference: message on stderr; run_se3_icp/pure break at the first optimisation (.cpp:700-703) leaving
        # R = I and, through .cpp:735-738, t = c_tgt - c_src; run_icp's behaviour is undefined -> identity here.
        if entry == capi.RUN_ICP:
            sys.stderr.write("Invalid ICP variant name. Valid names are pt2pt, pt2pl and gicp.\n")
            self.current_estimated_T_ = np.eye(4)
            return self.current_estimated_T_
        sys.stderr.write("Invalid variant name. Choose one of: pt2pt, pt2pl, gicp \n")
        T = np.eye(4)
        T[:3, 3] = self._target.mean(axis=0) - self._source.mean(axis=0)
        self.current_estimated_T_ = T
        self.num_iterations_ = 1
        self.num_pure_se3_iterations_ = 1
        return T

    def run_icp(self, variant_name):
        return self._run(capi.RUN_ICP, variant_name)

    def run_se3_icp(self, variant_name):
        return self._run(capi.RUN_SE3_ICP, variant_name)

    def run_se3_icp_with_cf(self):
        return self._run(capi.RUN_SE3_ICP_CF, "gicp")

    def run_se3_pure(self, variant_name):
        return self._run(capi.RUN_SE3_PURE, variant_name)

    # state the reference keeps as public members (hpp:59-60,74)
    @property
    def source_se3_cloud_(self):
        return self._ctx.se3_cloud(capi.SOURCE)

    @property
    def target_se3_cloud_(self):
        return self._ctx.se3_cloud(capi.TARGET)

    @property
    def current_correspondences_set(self):
        return self._ctx.correspondences()


METHODS = ("pt2pt", "pt2pl", "gicp", "se3_pt2pt", "se3_pt2pl", "se3_gicp", "se3_gicp_with_cf")


def run_registration_method(method, source, target, **fields):
    """examples/run_registration_method.cpp:35-57 as a function: method in METHODS; returns the object."""
    reg = IterativeSE3Registration(device=fields.pop("device", 0))
    reg.setSourceCloud(source)
    reg.setTargetCloud(target)
    reg.estimated_overlap_ = 1.0
    reg.max_num_se3_iterations_ = 10
    reg.mse_ = 0.00001
    reg.mse_switch_error_ = 5 * reg.mse_
    reg.number_of_nn_for_LRF_ = 90
    for k, v in fields.items():
        if not hasattr(reg, k):
            raise AttributeError(k)
        setattr(reg, k, v)
    if method in ("pt2pt", "pt2pl", "gicp"):
        reg.run_icp(method)
    elif method in ("se3_pt2pt", "se3_pt2pl", "se3_gicp"):
        reg.run_se3_icp(method[4:])
    elif method == "se3_gicp_with_cf":
        reg.run_se3_icp_with_cf()
    else:
        raise ValueError("Not a valid algorithm name; available: %s" % ", ".join(METHODS))
    return reg


def register_sequence(method, scans, device=0, **params):
    """Odometry-style driver (examples/benchmark_kitti.cpp:120-131): registers scans[i+1] onto scans[i] for every
    consecutive pair on one GPU context, computing each scan's neighbourhood features once (se3icp_run_sequence;
    extension of the reference API).  method: "se3_pt2pt" | "se3_pt2pl" | "se3_gicp" | "se3_gicp_with_cf" or a plain
    ICP name; params: fields of se3icp_params (e.g. the KITTI values of benchmark_kitti.cpp:133-148).
    Returns (T [n-1, 4, 4], [Stats])."""
    if method in ("pt2pt", "pt2pl", "gicp"):
        entry, variant = capi.RUN_ICP, method
    elif method in ("se3_pt2pt", "se3_pt2pl", "se3_gicp"):
        entry, variant = capi.RUN_SE3_ICP, method[4:]
    elif method == "se3_gicp_with_cf":
        entry, variant = capi.RUN_SE3_ICP_CF, "gicp"
    else:
        raise ValueError("Not a valid algorithm name; available: %s" % ", ".join(METHODS))
    ctx = capi.Context(device)
    return ctx.run_sequence(list(scans), capi.default_params(variant=variant, entry=entry, **params))


def make_hybrid_alpha_grid():
    """examples/benchmark_kitti.cpp:354-384 (makeHybridLGrid): the rotation-scale values of the reference's alpha sweep —
    0, 0.01..0.10 step 0.01, 0.2..1.0 step 0.1, 1.0..5.0 step 0.5, then a geometric tail up to 1000; sorted, unique."""
    grid = [0.0]
    grid += [i * 0.01 for i in range(1, 11)]
    grid += [i * 0.1 for i in range(2, 11)]
    grid += [1.0 + i * 0.5 for i in range(0, 9)]
    grid += [5, 7, 10, 15, 25, 50, 60, 70, 80, 90, 100, 200, 300, 400, 500, 600, 700, 800, 900, 1000]
    return sorted(set(float(a) for a in grid))


def benchmark_different_rot_scales(method, source, target, alphas=None, device=0, **params):
    """The reference's alpha-sweep harness (examples/benchmark_kitti.cpp:387-393, benchmark_different_rot_scales) for one
    pair: the same registration once per rotation scale alpha_rot.  The frames, normals and covariances of a cloud do not
    depend on alpha, so they are computed for the first value only and reused by the others (params.reuse_features; the
    12-D rows and their search structure are rebuilt per value).  Returns [(alpha, T 4x4, Stats)]."""
    if method in ("se3_pt2pt", "se3_pt2pl", "se3_gicp"):
        entry, variant = capi.RUN_SE3_ICP, method[4:]
    elif method == "se3_gicp_with_cf":
        entry, variant = capi.RUN_SE3_ICP_CF, "gicp"
    else:
        raise ValueError("the rotation scale only enters the SE(3) methods; available: %s" % ", ".join(METHODS[3:]))
    ctx = capi.Context(device)
    ctx.set_cloud(capi.SOURCE, source)
    ctx.set_cloud(capi.TARGET, target)
    out = []
    for alpha in (make_hybrid_alpha_grid() if alphas is None else alphas):
        p = capi.default_params(variant=variant, entry=entry, reuse_features=1, **dict(params, alpha_rot=float(alpha)))
        T, st = ctx.run(p)
        out.append((float(alpha), T, st))
    ctx.close()
    return out
