"""Runs the reference's own estimators on a packed batch and compares a candidate result with them.

TEST INFRASTRUCTURE ONLY: used by tests/ and by bench.py's cpu_baseline / --impl reference legs
(the checker, never the thing measured or shipped).  The reference here is the PoseLib wheel
unpacked into oracle/_ref (whl:_core.pyi:446-501, call site /root/reference/make_video.py:284);
where it is absent the C restatement (oracle/port.py) stands in and `kind` says so.
"""
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

MODEL_FIELDS = ("q", "t", "scale", "shift1", "shift2", "f1", "f2")


def _model_row(variant, obj):
    """12-vector (q4, t3, scale, shift1, shift2, f1, f2) of a wheel result object."""
    if variant in ("shared", "varying"):
        g = obj.geometry
        f1, f2 = obj.camera1.focal(), obj.camera2.focal()
    else:
        g, f1, f2 = obj, 1.0, 1.0
    return np.r_[np.array(g.pose.q), np.array(g.pose.t), g.scale, g.shift1, g.shift2, f1, f2]


def run_reference(variant, shift, batch, ransac_opt, bundle_opt, threads=None):
    """variant: 'calib' | 'shared' | 'varying'; batch: dict(offsets, x1, x2, d1, d2, cams).
    Returns dict(models [P,12], stats [P,3] (refinements, iterations, num_inliers), score [P], masks [N] bool,
    seconds, cores, kind)."""
    from oracle import build_ref, port, ref_wheel
    offs = batch["offsets"]
    P = len(offs) - 1
    threads = threads or (os.cpu_count() or 1)
    models = np.zeros((P, 12))
    stats = np.zeros((P, 3), dtype=np.int64)
    score = np.zeros(P)
    masks = np.zeros(int(offs[-1]), dtype=bool)
    use_ref = build_ref.have_ref()
    if use_ref:
        pl = ref_wheel.poselib()
        ro = dict(ransac_opt)
        ro["monodepth_estimate_shift"] = bool(shift)

        def run(i):
            sl = slice(offs[i], offs[i + 1])
            if variant == "calib":
                k = batch["cams"][i]
                cam1 = {"model": "PINHOLE", "width": -1, "height": -1, "params": list(k[:4])}
                cam2 = {"model": "PINHOLE", "width": -1, "height": -1, "params": list(k[4:])}
                m, info = pl.estimate_monodepth_relative_pose(batch["x1"][sl], batch["x2"][sl], batch["d1"][sl],
                                                              batch["d2"][sl], cam1, cam2, ro, bundle_opt)
            elif variant == "shared":
                m, info = pl.estimate_monodepth_shared_focal_relative_pose(batch["x1"][sl], batch["x2"][sl],
                                                                           batch["d1"][sl], batch["d2"][sl], ro, bundle_opt)
            else:
                m, info = pl.estimate_monodepth_varying_focal_relative_pose(batch["x1"][sl], batch["x2"][sl],
                                                                            batch["d1"][sl], batch["d2"][sl], ro, bundle_opt)
            models[i] = _model_row(variant, m)
            stats[i] = (info["refinements"], info["iterations"], info["num_inliers"])
            score[i] = info["model_score"]
            masks[sl] = np.asarray(info["inliers"], dtype=bool)
    else:
        port.build()
        vid = {"calib": 1 if shift else 0, "shared": 2, "varying": 3}[variant]
        keys = ("max_iterations", "min_iterations", "max_epipolar_error", "max_reproj_error", "seed")
        rop = port.ransac_opt(estimate_shift=bool(shift), **{k: ransac_opt[k] for k in keys if k in ransac_opt})
        bop = port.bundle_opt(**bundle_opt)

        def run(i):
            sl = slice(offs[i], offs[i + 1])
            k = batch["cams"][i] if batch.get("cams") is not None else None
            m, st, mk = port.estimate(vid, batch["x1"][sl], batch["x2"][sl], batch["d1"][sl], batch["d2"][sl],
                                      None if k is None else k[:4], None if k is None else k[4:], rop, bop)
            models[i] = np.r_[np.array(m.q), np.array(m.t), m.scale, m.shift1, m.shift2, m.f1, m.f2]
            stats[i] = (st.refinements, st.iterations, st.num_inliers)
            score[i] = st.model_score
            masks[sl] = mk

    run(0)  # warm (page in the library)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(run, range(P)))
    dt = time.perf_counter() - t0
    return dict(models=models, stats=stats, score=score, masks=masks, seconds=dt, cores=threads,
                kind="reference" if use_ref else "port")


def _canon(v):
    v = np.array(v, dtype=np.float64)
    if v[0] < 0 or (v[0] == 0 and v[1] < 0):
        v[:4] = -v[:4]
    return v


def struct_to_rows(models):
    """MODEL_DTYPE structured array -> [P,12]."""
    return np.concatenate([np.asarray(models[f], dtype=np.float64).reshape(len(models), -1) for f in MODEL_FIELDS], axis=1)


def compare(ref, offsets, models, stats, masks, rtol=1e-6, atol=1e-8):
    """Candidate (structured models / stats arrays, uint8 masks) against run_reference() output.
    Returns the parity record bench.py prints and the per-pair classification the tests assert on:
      identical       (refinements, iterations, num_inliers) and mask equal, model within rtol, score within 1e-9
      refinements_only  only the `refinements` counter differs, everything else as above (an LO more or less was
                      triggered on the way — DESIGN.md §3 tie rule, S2/S3 solver quirks — and led to the same result)
      different       anything else
    """
    P = len(offsets) - 1
    rows = struct_to_rows(models)
    cls = np.zeros(P, dtype=np.int8)  # 0 identical, 1 refinements only, 2 different
    n_stats = n_mask = n_model = n_score = n_tie = 0
    worst_model = 0.0
    for i in range(P):
        sl = slice(offsets[i], offsets[i + 1])
        st = (int(stats[i]["refinements"]), int(stats[i]["iterations"]), int(stats[i]["num_inliers"]))
        rs = tuple(int(v) for v in ref["stats"][i])
        stats_eq = st == rs
        ref_only = (not stats_eq) and st[1:] == rs[1:]
        tie = ref_only and abs(st[0] - rs[0]) == 1
        mask_eq = np.array_equal(np.asarray(masks[sl]).astype(bool), ref["masks"][sl])
        a, b = _canon(rows[i]), _canon(ref["models"][i])
        model_eq = bool(np.allclose(a, b, rtol=rtol, atol=atol, equal_nan=True))
        with np.errstate(all="ignore"):
            d = np.abs(a - b) / np.maximum(np.abs(b), 1e-2)
            if np.isfinite(d).all():
                worst_model = max(worst_model, float(d.max()))
        s, r = float(stats[i]["model_score"]), float(ref["score"][i])
        score_eq = s == r or abs(s - r) <= 1e-9 * abs(r)
        n_stats += stats_eq
        n_mask += mask_eq
        n_model += model_eq
        n_score += score_eq
        n_tie += tie
        if stats_eq and mask_eq and model_eq and score_eq:
            cls[i] = 0
        elif ref_only and mask_eq and model_eq and score_eq:
            cls[i] = 1
        else:
            cls[i] = 2
    rec = {"pairs": P, "stats_equal": int(n_stats), "mask_equal": int(n_mask), "model_1e-6": int(n_model),
           "score_1e-9": int(n_score), "refinements_off_by_one": int(n_tie), "identical": int((cls == 0).sum()),
           "refinements_only": int((cls == 1).sum()),
           "different": int((cls == 2).sum()), "max_model_rel_diff": worst_model, "against": ref["kind"]}
    return rec, cls
