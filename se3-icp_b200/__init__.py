"""se3-icp_b200: B200-native (sm_100a) SE(3)-ICP registration path behind the reference's class API.

csrc/          hand-written CUDA kernels + the C ABI (include/se3icp.h) -> libse3icp_cuda.so
host/          C++17 host class with the reference's declaration, calling the C ABI
capi.py        ctypes binding of the C ABI
registration.py  Python mirror of IterativeSE3Registration (same names / defaults / error behaviour)
"""
from . import capi, sharding  # noqa: F401
from .registration import (IterativeSE3Registration, run_registration_method, register_sequence, METHODS,  # noqa: F401
                           make_hybrid_alpha_grid, benchmark_different_rot_scales)
