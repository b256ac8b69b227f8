"""Golden SHOT frames from the REFERENCE'S OWN SOURCE (computeAllSHOTSE3FramesOMP, .cpp:226-239, through
oracle/_ref/libse3icp_reference.so; see make_golden_reference.py for how that library is built):

    python tests/golden/make_golden_shot.py        # needs /root/reference

Cases: the normalised source of the bundled fixture at the class default radius (`lrf_radius_` = 0.8, .cpp:340) and at 0.4,
and a 3 000-point noisy bunny sample at 0.5.  Only the 3x3 rotations are stored (the translation column is the point).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import workloads as W  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def clouds():
    src, tgt, _ = W.load_c1()
    cs, ct = src.mean(0), tgt.mean(0)
    s = 3.0 / max(np.linalg.norm(src - cs, axis=1).max(), np.linalg.norm(tgt - ct, axis=1).max())
    c1 = (src - cs) * s
    b, _, _ = W.bunny_problem("easy", seed=5, n_points=3000)
    b = (b - b.mean(0)) * (3.0 / np.linalg.norm(b - b.mean(0), axis=1).max())
    return {"c1_r0.8": (c1, 0.8), "c1_r0.4": (c1, 0.4), "bunny3k_r0.5": (b, 0.5)}


def main():
    from oracle import reference_build as RB
    out = {}
    for name, (xyz, r) in clouds().items():
        out[name] = RB.shot(xyz, r)[:, :3, :3]
        assert np.isfinite(out[name]).all(), name
    np.savez_compressed(os.path.join(OUT, "shot_reference.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
