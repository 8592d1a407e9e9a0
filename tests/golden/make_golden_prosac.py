#!/usr/bin/env python
"""Golden vectors for the PROSAC branch of the sampler (progressive_sampling=True), generated from the reference
binary in THIS container (oracle/_ref, unpacked from /root/reference/demo/*.whl by oracle/build_ref.py):

  * RandomSampler::initialize_prosac so@0x4f8a20 growth tables and RandomSampler::generate_sample so@0x4f8970
    sample sequences, called through ctypes on a hand-built RandomSampler (oracle/ref_wheel.py);
  * end-to-end outputs of the three estimate_monodepth_* entry points with
    ransac_opt = {progressive_sampling: True, max_prosac_iterations: ...} on correspondences sorted by a
    noisy quality score (inliers tend to come first), the situation PROSAC is made for.

    python tests/golden/make_golden_prosac.py      ->  tests/golden/prosac.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from mdrp_b200 import synth  # noqa: E402
from oracle import build_ref, ref_wheel as rw  # noqa: E402

SAMPLER_CASES = [(3, 0, 100000, 64), (7, 5, 100000, 400), (50, 1, 100000, 3000), (2000, 0, 100000, 4000),
                 (300, 3, 200, 600), (10000, 123456789, 100000, 2000), (4, 9, 1000, 300)]
E2E_CASES = [("calib", "cfg1_calib_scale", 300, 300, 100000), ("calib_shift", "cfg2_calib_shift", 300, 300, 100000),
             ("shared", "cfg3_shared_focal", 300, 300, 100000), ("varying", "cfg4_varying_focal", 300, 300, 100000),
             ("calib", "hard_calib", 400, 500, 250), ("calib_shift", "cfg2_calib_shift", 200, 400, 150)]


def quality_sorted(sc, rng):
    """Permutation that puts likely inliers first (a matcher's confidence ordering)."""
    r = np.linalg.norm(sc.x2 - sc.x2_clean, axis=1) if hasattr(sc, "x2_clean") else None
    if r is None:
        # no ground-truth reprojection stored with the scene: use the Sampson-free proxy |x2 - pi(R X1 + t)|
        z1 = sc.d1 + 0.0
        K1i = np.array([[1 / sc.f1, 0, -640 / sc.f1], [0, 1 / sc.f1, -480 / sc.f1], [0, 0, 1]])
        X = (K1i @ np.c_[sc.x1, np.ones(len(z1))].T) * (z1 + sc.shift1 if hasattr(sc, "shift1") else z1)
        Y = sc.R @ X + sc.t[:, None]
        p = np.c_[sc.f2 * Y[0] / Y[2] + 640, sc.f2 * Y[1] / Y[2] + 480]
        r = np.linalg.norm(p - sc.x2, axis=1)
    score = r * np.exp(0.5 * rng.standard_normal(len(r)))
    return np.argsort(score, kind="stable")


def main():
    assert build_ref.build(), "reference wheel not available"
    pl = rw.poselib()
    out = {}
    for n, seed, mp, iters in SAMPLER_CASES:
        s, growth = rw.generate_samples(n, 3, seed, True, mp, iters)
        out[f"sampler_n{n}_s{seed}_m{mp}"] = s.astype(np.int32)
        out[f"growth_n{n}_m{mp}"] = growth
    rng = np.random.default_rng(77)
    for name, cfg, n, iters, mp in E2E_CASES:
        for idx in range(2):
            sc = synth.scene_for(cfg, 300 + idx, n=n)
            c = synth.CONFIGS[cfg]
            perm = quality_sorted(sc, rng)
            ro = {"max_epipolar_error": 2.0, "max_reproj_error": 16.0, "seed": idx, "max_iterations": iters,
                  "min_iterations": iters, "monodepth_estimate_shift": c["shift"], "progressive_sampling": True,
                  "max_prosac_iterations": mp}
            bo = {"loss_type": "TRUNCATED_CAUCHY"}
            d1, d2 = sc.d1[perm], sc.d2[perm]
            if c["variant"] == "calib":
                c1, c2 = sc.camera_dicts()
                x1, x2 = sc.x1[perm], sc.x2[perm]
                g, info = pl.estimate_monodepth_relative_pose(x1, x2, d1, d2, c1, c2, ro, bo)
                model = np.r_[np.array(g.pose.q), np.array(g.pose.t).ravel(), g.scale, g.shift1, g.shift2, 1.0, 1.0]
            else:
                x1, x2 = (a[perm] for a in sc.centred())
                fn = (pl.estimate_monodepth_shared_focal_relative_pose if c["variant"] == "shared"
                      else pl.estimate_monodepth_varying_focal_relative_pose)
                p, info = fn(x1, x2, d1, d2, ro, bo)
                g = p.geometry
                model = np.r_[np.array(g.pose.q), np.array(g.pose.t).ravel(), g.scale, g.shift1, g.shift2,
                              p.camera1.focal(), p.camera2.focal()]
            key = f"{name}_{cfg}_n{n}_{idx}"
            out[key + "_x1"], out[key + "_x2"], out[key + "_d1"], out[key + "_d2"] = x1, x2, d1, d2
            out[key + "_f"] = np.array([sc.f1, sc.f2])
            out[key + "_opts"] = np.array([iters, mp], dtype=np.int64)
            out[key + "_model"] = model
            out[key + "_stats"] = np.array([info["refinements"], info["iterations"], info["num_inliers"]], dtype=np.int64)
            out[key + "_fstats"] = np.array([info["inlier_ratio"], info["model_score"]])
            out[key + "_mask"] = np.array(info["inliers"], dtype=np.uint8)
    path = os.path.join(HERE, "prosac.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
