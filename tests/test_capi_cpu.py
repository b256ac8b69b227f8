"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/se3icp.h declares,
and refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "se3icp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(se3icp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(capi):
    lib = capi.lib()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libse3icp_cuda.so does not export %s" % n
    assert sorted(capi.EXPORTED_SYMBOLS) == names
    assert lib.se3icp_abi_version() == 4


def test_default_params_match_reference_ctor(capi):
    p = capi.default_params()  # reference .cpp:334-348
    assert (p.max_num_iterations, p.max_num_se3_iterations, p.number_of_nn_for_LRF) == (150, 20, 30)
    assert (p.mse, p.mse_switch_error, p.estimated_overlap) == (1e-5, 1e-3, 1.0)
    assert (p.alpha_rot, p.beta_transl, p.scale_preprocessing) == (3.0, 1.0, 3.0)
    assert (p.knn_normals_pt2pl, p.knn_normals_gicp, p.gicp_epsilon) == (30, 20, 1e-3)
    assert (p.lrf_method, p.lrf_radius) == (capi.LRF_TOLDI, 0.8)  # .cpp:340; TOLDI is the reference's active frame


def test_params_layout_matches_oracle(capi, orc):
    """the algorithmic fields of se3icp_params and orc_params have identical order and types: the reference's 15, then
    the execution switches that exist only on the CUDA side, then the frame choice both sides share"""
    a = [(n, t) for n, t in capi.Params._fields_]
    b = [(n, t) for n, t in orc.Params._fields_]
    assert a[:15] == b[:15]
    assert a[-3:] == b[-3:] and [n for n, _ in b[-3:]] == ["lrf_method", "reserved0", "lrf_radius"]


def test_no_cpu_fallback(capi):
    """without a usable sm_100 device the product must fail loudly"""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(capi.Se3IcpError) as e:
        capi.Context(0)
    assert e.value.code == 2  # SE3ICP_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "se3-icp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.replace("the oracle", "").replace("oracle's", "").replace("oracle is", "") or \
                    "import" not in txt or not re.search(r"(import|include|dlopen).*oracle", txt), f
