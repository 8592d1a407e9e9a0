#!/usr/bin/env python
"""h5py/madpose-free clone of the reference's eval_experiment (eval.py:93-160, eval_shared_f.py:111-183)
on synthetic scenes: same experiment strings -> same option dicts -> same fork-API calls, then the
reference's metrics (mAA(10 deg) of max(R_err, t_err), eval_utils.py:41-67).

    python tools/eval_synth.py --pairs 200 --matches 1000                 # B200 path, batched
    python tools/eval_synth.py --pairs 50 --backend reference             # the reference wheel (CPU)
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdrp_b200 import synth  # noqa: E402

EXPERIMENTS = ["p3p_hybrid_ctruncated", "3p_ours_shift_scale_hybrid-s_ctruncated", "3p_ours_scale_hybrid_ctruncated"]


def dicts_for(experiment, iters, t=2.0, r=16.0):
    """eval.py:99-129."""
    lo_iterations = 0 if "nLO" in experiment else 25
    ransac = {"max_iterations": iters, "max_epipolar_error": t, "progressive_sampling": False, "min_iterations": iters,
              "lo_iterations": lo_iterations, "max_reproj_error": r, "all_permutations": True,
              "use_reldepth": "reldepth" in experiment, "use_p3p": "p3p" in experiment, "use_ours": "ours" in experiment,
              "use_madpose": "mad_poselib" in experiment, "solver_shift": "shift" in experiment,
              "solver_scale": "scale" in experiment, "optimize_hybrid": "hybrid" in experiment,
              "optimize_shift": "reproj-s" in experiment or "hybrid-s" in experiment, "weight_sampson": 1.0}
    bundle = {"max_iterations": 0 if lo_iterations == 0 else 100, "verbose": False}
    if "truncated" in experiment:
        bundle["loss_type"] = "TRUNCATED"
    if "ctruncated" in experiment:
        bundle["loss_type"] = "TRUNCATED_CAUCHY"
    return ransac, bundle


def pose_error(R, t, R_gt, t_gt):
    return max(synth.rotation_error_deg(R, R_gt), synth.translation_error_deg(t, t_gt))


def maa(errs):
    e = np.array([180.0 if not np.isfinite(x) else x for x in errs])
    return 100.0 * np.mean([np.mean(e < th) for th in range(1, 11)])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=100)
    ap.add_argument("--matches", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--config", default="hard_calib")
    ap.add_argument("--backend", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    scenes = [synth.scene_for(args.config, 5000 + i, n=args.matches) for i in range(args.pairs)]
    if args.backend == "b200":
        sys.path.insert(0, os.path.join(ROOT, "mdrp_b200", "compat"))
        import poselib
    else:
        from oracle import ref_wheel
        poselib = ref_wheel.poselib()
    print(f"{'experiment':45s} {'median':>8s} {'mAA(10)':>8s} {'ms/pair':>9s}")
    for exp in EXPERIMENTS:
        ransac, bundle = dicts_for(exp, args.iters)
        cams = [{"model": "PINHOLE", "width": -1, "height": -1, "params": [s.f1, s.f1, 640.0, 480.0]} for s in scenes]
        shift = "shift" in exp and "ours" in exp and "p3p" not in exp
        t0 = time.perf_counter()
        if args.backend == "b200":
            from mdrp_b200 import api
            res = poselib.estimate_monodepth_relative_pose_batch(
                [s.x1 for s in scenes], [s.x2 for s in scenes], [s.d1 for s in scenes], [s.d2 for s in scenes],
                cams, cams, api._fork_ransac(ransac), bundle)
            poses = [g.pose for g, _ in res]
        else:
            ro = dict(ransac, monodepth_estimate_shift=shift)
            poses = [poselib.estimate_monodepth_relative_pose(s.x1, s.x2, s.d1, s.d2, c, c, ro, bundle)[0].pose
                     for s, c in zip(scenes, cams)]
        dt = time.perf_counter() - t0
        errs = [pose_error(np.array(p.R), np.array(p.t).ravel(), s.R, s.t) for p, s in zip(poses, scenes)]
        print(f"{exp:45s} {np.median(errs):8.3f} {maa(errs):8.2f} {1000 * dt / len(scenes):9.3f}")


if __name__ == "__main__":
    main()
