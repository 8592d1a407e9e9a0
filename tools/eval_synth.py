#!/usr/bin/env python
"""h5py/madpose-free clone of the reference's eval_experiment (eval.py:93-160, eval_shared_f.py:110-183,
eval_varying_f.py:110-178) on synthetic scenes: same experiment strings -> same option dicts -> the fork-API
calls of the three drivers, then the reference's metrics (median / mAA(10 deg) of max(R_err, t_err) and, for the
focal drivers, the median focal error and mAA_f, utils/eval_utils.py:8-67).

    python tools/eval_synth.py --pairs 200 --matches 1000                 # B200 path, batched
    python tools/eval_synth.py --pairs 50 --backend reference             # the reference wheel (CPU)
    python tools/eval_synth.py --driver shared --config cfg3_shared_focal # eval_shared_f.py's experiments
    python tools/eval_synth.py --driver varying --config cfg4_varying_focal
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdrp_b200 import synth  # noqa: E402

EXPERIMENTS = {"calib": ["p3p_hybrid_ctruncated", "3p_ours_shift_scale_hybrid-s_ctruncated", "3p_ours_scale_hybrid_ctruncated"],
               "shared": ["3p_ours_scale_hybrid_truncated", "3p_ours_scale_hybrid_ctruncated"],     # eval_shared_f.py:276-277
               "varying": ["3p_ours_scale_hybrid_truncated", "3p_ours_scale_hybrid_ctruncated"]}    # eval_varying_f.py:265


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


def run(driver, config, pairs, matches, iters, backend, out=print):
    """Returns {experiment: (median pose err, mAA, median f err or None)}."""
    scenes = [synth.scene_for(config, 5000 + i, n=matches) for i in range(pairs)]
    if backend == "b200":
        sys.path.insert(0, os.path.join(ROOT, "mdrp_b200", "compat"))
        import poselib
        from mdrp_b200 import api
    else:
        from oracle import ref_wheel
        poselib = ref_wheel.poselib()
    out(f"{'experiment':45s} {'median':>8s} {'mAA(10)':>8s} {'f err':>8s} {'ms/pair':>9s}")
    results = {}
    for exp in EXPERIMENTS[driver]:
        ransac, bundle = dicts_for(exp, iters)
        t0 = time.perf_counter()
        f_errs = None
        if driver == "calib":
            cams = [{"model": "PINHOLE", "width": -1, "height": -1, "params": [s.f1, s.f1, 640.0, 480.0]} for s in scenes]
            shift = "shift" in exp and "ours" in exp and "p3p" not in exp
            if backend == "b200":
                res = poselib.estimate_monodepth_relative_pose_batch(
                    [s.x1 for s in scenes], [s.x2 for s in scenes], [s.d1 for s in scenes], [s.d2 for s in scenes],
                    cams, cams, api._fork_ransac(ransac), bundle)
                poses = [g.pose for g, _ in res]
            else:
                ro = dict(ransac, monodepth_estimate_shift=shift)
                poses = [poselib.estimate_monodepth_relative_pose(s.x1, s.x2, s.d1, s.d2, c, c, ro, bundle)[0].pose
                         for s, c in zip(scenes, cams)]
        else:
            cent = [s.centred() for s in scenes]
            if backend == "b200":
                fn = (poselib.estimate_monodepth_shared_focal_relative_pose_batch if driver == "shared"
                      else poselib.estimate_monodepth_varying_focal_relative_pose_batch)
                res = fn([c[0] for c in cent], [c[1] for c in cent], [s.d1 for s in scenes], [s.d2 for s in scenes],
                         api._fork_ransac(ransac), bundle)
            else:
                fn = (poselib.estimate_monodepth_shared_focal_relative_pose if driver == "shared"
                      else poselib.estimate_monodepth_varying_focal_relative_pose)
                res = [fn(c[0], c[1], s.d1, s.d2, ransac, bundle) for s, c in zip(scenes, cent)]
            poses = [ip.geometry.pose for ip, _ in res]
            f_errs = [np.sqrt(abs(ip.camera1.focal() - s.f1) / s.f1 * abs(ip.camera2.focal() - s.f2) / s.f2)
                      for (ip, _), s in zip(res, scenes)]
        dt = time.perf_counter() - t0
        errs = [pose_error(np.array(p.R), np.array(p.t).ravel(), s.R, s.t) for p, s in zip(poses, scenes)]
        fe = float(np.median(f_errs)) if f_errs is not None else None
        results[exp] = (float(np.median(errs)), float(maa(errs)), fe)
        out(f"{exp:45s} {np.median(errs):8.3f} {maa(errs):8.2f} {('%8.4f' % fe) if fe is not None else '       -'} "
            f"{1000 * dt / len(scenes):9.3f}")
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=100)
    ap.add_argument("--matches", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--driver", default="calib", choices=["calib", "shared", "varying"])
    ap.add_argument("--config", default=None)
    ap.add_argument("--backend", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    config = args.config or {"calib": "hard_calib", "shared": "cfg3_shared_focal", "varying": "cfg4_varying_focal"}[args.driver]
    run(args.driver, config, args.pairs, args.matches, args.iters, args.backend)


if __name__ == "__main__":
    main()
