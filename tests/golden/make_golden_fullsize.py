"""Full-size golden transforms from the CPU oracle for BASELINE.json configs[2] and [3]
(KITTI-like se3_gicp, lounge-like se3_gicp_with_cf) so the GPU tests have a parity check at full size
without running the oracle for tens of seconds inside the test suite.

    python tests/golden/make_golden_fullsize.py      # ~1 min on 8 cores
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import workloads as W  # noqa: E402
from oracle import oracle as orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    res = {}
    s, t, _ = W.lidar_pair(seed=0)
    T, st, _ = orc.run(s, t, orc.default_params(variant="gicp", entry=orc.RUN_SE3_ICP, **W.KITTI_PARAMS))
    res.update(kitti_T=T, kitti_it=[st.num_iterations, st.num_pure_se3_iterations], kitti_n=[len(s), len(t)])
    s, t, _ = W.rgbd_pair(seed=0)
    T, st, _ = orc.run(s, t, orc.default_params(variant="gicp", entry=orc.RUN_SE3_ICP_CF, **W.LOUNGE_PARAMS))
    res.update(lounge_T=T, lounge_it=[st.num_iterations, st.num_pure_se3_iterations], lounge_n=[len(s), len(t)])
    for seed in (1, 2):  # two more seeds per configuration (keys suffixed _s<seed>)
        s, t, _ = W.lidar_pair(seed=seed)
        T, st, _ = orc.run(s, t, orc.default_params(variant="gicp", entry=orc.RUN_SE3_ICP, **W.KITTI_PARAMS))
        res.update({"kitti_T_s%d" % seed: T, "kitti_it_s%d" % seed: [st.num_iterations, st.num_pure_se3_iterations],
                    "kitti_n_s%d" % seed: [len(s), len(t)]})
        s, t, _ = W.rgbd_pair(seed=seed)
        T, st, _ = orc.run(s, t, orc.default_params(variant="gicp", entry=orc.RUN_SE3_ICP_CF, **W.LOUNGE_PARAMS))
        res.update({"lounge_T_s%d" % seed: T, "lounge_it_s%d" % seed: [st.num_iterations, st.num_pure_se3_iterations],
                    "lounge_n_s%d" % seed: [len(s), len(t)]})
    np.savez(os.path.join(OUT, "fullsize_oracle.npz"), **res)
    for k, v in res.items():
        print(k, v if np.size(v) < 4 else "\n%s" % v)


if __name__ == "__main__":
    main()
