"""Golden vectors from the oracle on BASELINE.json configs[0] (se3_pt2pl on the bundled fixture).

    python tests/golden/make_golden.py

Writes tests/golden/c1_se3_pt2pl_trace.npz: per-iteration estimates, mean distances, phase flags,
iteration counts, the final transform, the first iteration's correspondences and the TOLDI frames of
both clouds.  The reference itself cannot be executed here (Open3D/PCL/Eigen absent), so these pin
the oracle against regressions; the only reference-derived answers in them are the ground truth and
the 8 = 6 + 2 iteration count of SURVEY.md §4.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    src, tgt = np.load(OUT + "/c1_source.npy"), np.load(OUT + "/c1_target.npy")
    p = orc.default_params(variant="pt2pl", entry=orc.RUN_SE3_ICP, estimated_overlap=1.0, max_num_se3_iterations=10,
                           mse=1e-5, mse_switch_error=5e-5, number_of_nn_for_LRF=90)
    T, st, tr = orc.run(src, tgt, p, trace_iters=20)
    # frames of the NORMALISED clouds, as the registration sees them
    c_s, c_t = src.mean(0), tgt.mean(0)
    s = st.scaling_factor
    fs, ft = orc.toldi((src - c_s) * s, 90), orc.toldi((tgt - c_t) * s, 90)
    np.savez_compressed(OUT + "/c1_se3_pt2pl_trace.npz", T_final=T, num_iterations=st.num_iterations,
                        num_se3=st.num_pure_se3_iterations, scaling_factor=s, T_iter=tr["T_iter"],
                        mean_dist=tr["mean_dist"], se3_phase=tr["se3_phase"], corr_idx0=tr["corr_idx"][0],
                        corr_dist0=tr["corr_dist"][0], corr_idx_last=tr["corr_idx"][-1],
                        frames_src=fs[:, :3, :3].astype(np.float32), frames_tgt=ft[:, :3, :3].astype(np.float32))
    print("iterations", st.num_iterations, st.num_pure_se3_iterations)


if __name__ == "__main__":
    main()
