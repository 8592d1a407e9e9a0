"""Generates tests/golden/*.npz from the REFERENCE binary (oracle/_ref = the PoseLib wheel the
reference ships, /root/reference/demo/poselib-2.0.5-cp312-cp312-linux_x86_64.whl).

Run in the build container (where /root/reference is mounted): python tests/golden/make_golden.py
The .npz files are committed; the GPU box never needs the reference tree.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mdrp_b200 import synth  # noqa: E402
from oracle import build_ref, ref_wheel as rw  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def normalised(sc, variant):
    if variant in ("calib", "calib_shift"):
        return (sc.x1 - synth.PP) / sc.f1, (sc.x2 - synth.PP) / sc.f2, sc.f1
    x1, x2 = sc.centred()
    ns = (np.linalg.norm(x1, axis=1) + np.linalg.norm(x2, axis=1)).sum() / (np.sqrt(2) * len(x1))
    return x1 / ns, x2 / ns, ns


def flat_geom(g):
    return np.r_[np.array(g.pose.q).ravel(), np.array(g.pose.t).ravel(), g.scale, g.shift1, g.shift2, 1.0, 1.0]


def flat_pair(p):
    g = p.geometry
    return np.r_[np.array(g.pose.q).ravel(), np.array(g.pose.t).ravel(), g.scale, g.shift1, g.shift2,
                 p.camera1.focal(), p.camera2.focal()]


def main():
    assert build_ref.build(), "reference wheel not available"
    pl = rw.poselib()
    rng = np.random.default_rng(2024)
    out = {}

    # R2 sampler ---------------------------------------------------------------------------
    for n, seed, iters in ((3, 0, 64), (7, 5, 200), (2000, 0, 300), (10000, 123456789, 300)):
        st = seed
        rows = []
        for _ in range(iters):
            s, st = rw.draw_sample(3, n, st)
            rows.append(s)
        out[f"sampler_n{n}_s{seed}"] = np.array(rows, dtype=np.int32)
        out[f"sampler_n{n}_s{seed}_state"] = np.array([st], dtype=np.uint64)

    cfg_of = {"calib": "cfg1_calib_scale", "calib_shift": "cfg2_calib_shift", "shared": "cfg3_shared_focal",
              "varying": "cfg4_varying_focal"}
    # S1-S4 solvers ------------------------------------------------------------------------
    for variant, cfg in cfg_of.items():
        sc = synth.scene_for(cfg, 7, n=400)
        x1, x2, _ = normalised(sc, variant)
        n_prob = 150
        idx = np.stack([rng.choice(len(x1), 3, replace=False) for _ in range(n_prob)])
        x1h = np.concatenate([x1[idx], np.ones((n_prob, 3, 1))], axis=2)
        x2h = np.concatenate([x2[idx], np.ones((n_prob, 3, 1))], axis=2)
        d1, d2 = sc.d1[idx], sc.d2[idx]
        sols = np.full((n_prob, 4, 12), np.nan)
        cnt = np.zeros(n_prob, dtype=np.int32)
        for i in range(n_prob):
            if variant == "calib":
                X = x1h[i] * d1[i][:, None]
                b = x2h[i] / np.linalg.norm(x2h[i], axis=1)[:, None]
                res = []
                for p in pl.p3p(b, X):
                    R, t = np.array(p.R), np.array(p.t).ravel()
                    res.append(np.r_[np.array(p.q).ravel(), t, (R @ X[0] + t)[0] / (d2[i, 0] * x2h[i, 0, 0]), 0, 0, 1, 1])
            elif variant == "calib_shift":
                res = [flat_geom(g) for g in pl.monodepth_pose_3pt(x1h[i], x2h[i], list(d1[i]), list(d2[i]))]
            elif variant == "shared":
                res = [flat_pair(g) for g in pl.shared_focal_monodepth_pose_3pt(x1h[i], x2h[i], list(d1[i]), list(d2[i]))]
            else:
                res = [flat_pair(g) for g in pl.varying_focal_monodepth_pose_4pt(x1h[i], x2h[i], list(d1[i]), list(d2[i]))]
            cnt[i] = len(res)
            for k, r in enumerate(res[:4]):
                sols[i, k] = r
        out[f"solve_{variant}_x1h"], out[f"solve_{variant}_x2h"] = x1h, x2h
        out[f"solve_{variant}_d1"], out[f"solve_{variant}_d2"] = d1, d2
        out[f"solve_{variant}_models"], out[f"solve_{variant}_counts"] = sols, cnt

    # SC / SF scorer + I1 masks --------------------------------------------------------------
    for variant, cfg in cfg_of.items():
        sc = synth.scene_for(cfg, 11, n=500)
        x1, x2, ns = normalised(sc, variant)
        thr2 = (2.0 / ns) ** 2
        models = out[f"solve_{variant}_models"]
        # models solved on another scene are (mostly) bad hypotheses here; add near-GT ones
        cand = [m for m in models.reshape(-1, 12) if np.isfinite(m).all()][:60]
        from scipy.spatial.transform import Rotation as Rot
        qq = Rot.from_matrix(sc.R).as_quat()
        qgt = np.array([qq[3], qq[0], qq[1], qq[2]])
        for j in range(40):
            q = qgt + 2e-3 * rng.normal(size=4)
            q /= np.linalg.norm(q)
            t = sc.t + 2e-3 * rng.normal(size=3)
            f1 = sc.f1 / ns if variant in ("shared", "varying") else 1.0
            f2 = sc.f2 / ns if variant in ("shared", "varying") else 1.0
            cand.append(np.r_[q, t, 1.7, 0, 0, f1, f2])
        cand = np.array(cand)
        scores, counts, masks = [], [], []
        for m in cand:
            if variant in ("calib", "calib_shift"):
                s, c = rw.msac_score_pose(m[:4], m[4:7], x1, x2, thr2)
                mk = rw.get_inliers_pose(m[:4], m[4:7], x1, x2, thr2)
            else:
                E = rw.essential_from_motion(m[:4], m[4:7])
                F = np.diag([1, 1, m[11]]) @ E @ np.diag([1, 1, m[10]])
                # build F exactly as the estimators do: diag(1,1,f2) * E * diag(1,1,f1), elementwise
                F = E.copy()
                F[2, :] = m[11] * F[2, :]
                F[:, 2] = F[:, 2] * m[10]
                s, c = rw.msac_score_F(F, x1, x2, thr2)
                mk = rw.get_inliers_F(F, x1, x2, thr2)
            scores.append(s); counts.append(c); masks.append(mk)
        out[f"score_{variant}_x1"], out[f"score_{variant}_x2"] = x1, x2
        out[f"score_{variant}_thr2"] = np.array([thr2])
        out[f"score_{variant}_models"] = cand
        out[f"score_{variant}_scores"] = np.array(scores)
        out[f"score_{variant}_counts"] = np.array(counts, dtype=np.int64)
        out[f"score_{variant}_masks"] = np.array(masks, dtype=np.uint8)

    # L2 refinement --------------------------------------------------------------------------------
    for variant, cfg in cfg_of.items():
        sc = synth.scene_for(cfg, 13, n=400)
        x1, x2, ns = normalised(sc, variant)
        thr = 2.0 / ns
        sr = (2.0 / 16.0) ** 2
        from scipy.spatial.transform import Rotation as Rot
        qq = Rot.from_matrix(sc.R).as_quat()
        qgt = np.array([qq[3], qq[0], qq[1], qq[2]])
        starts, ends, costs, its = [], [], [], []
        for loss, iters in (("TRUNCATED", 25), ("TRUNCATED_CAUCHY", 100)):
            for j in range(6):
                q = qgt + 5e-3 * rng.normal(size=4)
                q /= np.linalg.norm(q)
                t = sc.t + 1e-2 * rng.normal(size=3)
                s = 1.7 * (1 + 0.01 * rng.normal())
                bo = rw.bundle_options(max_iterations=iters, loss_type=loss, loss_scale=thr)
                if variant == "calib":
                    (g, st) = rw.refine_calib(x1, x2, sc.d1, sc.d2, q, t, s, 0, 0, sr, 1.0, bo, False)
                    start = np.r_[q, t, s, 0, 0, 1, 1]; end = np.r_[g[0], g[1], g[2], g[3], g[4], 1, 1]
                elif variant == "calib_shift":
                    (g, st) = rw.refine_calib(x1, x2, sc.d1, sc.d2, q, t, s, 0.02, -0.03, sr, 1.0, bo, True)
                    start = np.r_[q, t, s, 0.02, -0.03, 1, 1]; end = np.r_[g[0], g[1], g[2], g[3], g[4], 1, 1]
                elif variant == "shared":
                    f = sc.f1 / ns * 1.02
                    (g, st) = rw.refine_shared(x1, x2, sc.d1, sc.d2, q, t, s, f, sr, 1.0, bo)
                    start = np.r_[q, t, s, 0, 0, f, f]; end = np.r_[g[0], g[1], g[2], 0, 0, g[3], g[4]]
                else:
                    f1, f2 = sc.f1 / ns * 1.02, sc.f2 / ns * 0.98
                    (g, st) = rw.refine_varying(x1, x2, sc.d1, sc.d2, q, t, s, f1, f2, sr, 1.0, bo)
                    start = np.r_[q, t, s, 0, 0, f1, f2]; end = np.r_[g[0], g[1], g[2], 0, 0, g[3], g[4]]
                starts.append(start); ends.append(end); costs.append([st.initial_cost, st.cost]); its.append(st.iterations)
        out[f"refine_{variant}_x1"], out[f"refine_{variant}_x2"] = x1, x2
        out[f"refine_{variant}_d1"], out[f"refine_{variant}_d2"] = sc.d1, sc.d2
        out[f"refine_{variant}_thr"] = np.array([thr])
        out[f"refine_{variant}_start"], out[f"refine_{variant}_end"] = np.array(starts), np.array(ends)
        out[f"refine_{variant}_cost"], out[f"refine_{variant}_iters"] = np.array(costs), np.array(its)

    np.savez_compressed(os.path.join(HERE, "stages.npz"), **out)

    # end-to-end -------------------------------------------------------------------------------------
    e2e = {}
    cases = [("calib", "cfg1_calib_scale", 300, 300), ("calib_shift", "cfg2_calib_shift", 300, 300),
             ("shared", "cfg3_shared_focal", 300, 300), ("varying", "cfg4_varying_focal", 300, 300),
             ("calib", "hard_calib", 400, 500), ("calib_default_iters", "cfg1_calib_scale", 150, None)]
    for name, cfg, n, iters in cases:
        for idx in range(3):
            sc = synth.scene_for(cfg, 100 + idx, n=n)
            c = synth.CONFIGS[cfg]
            ro = {"max_epipolar_error": 2.0, "max_reproj_error": 16.0, "seed": idx,
                  "monodepth_estimate_shift": c["shift"]}
            if iters is not None:
                ro["max_iterations"] = ro["min_iterations"] = iters
            else:
                ro["min_iterations"] = 100
                ro["max_iterations"] = 5000
            bo = {"loss_type": "TRUNCATED_CAUCHY"}
            if c["variant"] == "calib":
                c1, c2 = sc.camera_dicts()
                g, info = pl.estimate_monodepth_relative_pose(sc.x1, sc.x2, sc.d1, sc.d2, c1, c2, ro, bo)
                model = flat_geom(g)
                x1, x2 = sc.x1, sc.x2
            else:
                x1, x2 = sc.centred()
                fn = (pl.estimate_monodepth_shared_focal_relative_pose if c["variant"] == "shared"
                      else pl.estimate_monodepth_varying_focal_relative_pose)
                g, info = fn(x1, x2, sc.d1, sc.d2, ro, bo)
                model = flat_pair(g)
            key = f"{name}_{cfg}_{idx}"
            e2e[key + "_x1"], e2e[key + "_x2"], e2e[key + "_d1"], e2e[key + "_d2"] = x1, x2, sc.d1, sc.d2
            e2e[key + "_f"] = np.array([sc.f1, sc.f2])
            e2e[key + "_iters"] = np.array([-1 if iters is None else iters])
            e2e[key + "_model"] = model
            e2e[key + "_stats"] = np.array([info["refinements"], info["iterations"], info["num_inliers"]], dtype=np.int64)
            e2e[key + "_fstats"] = np.array([info["inlier_ratio"], info["model_score"]])
            e2e[key + "_mask"] = np.array(info["inliers"], dtype=np.uint8)
            e2e[key + "_Rt_gt"] = np.c_[sc.R, sc.t]
    np.savez_compressed(os.path.join(HERE, "e2e.npz"), **e2e)
    print("wrote", os.path.join(HERE, "stages.npz"), os.path.join(HERE, "e2e.npz"))


if __name__ == "__main__":
    main()
