"""Reader of the reference's H5 benchmark files -> packed batches for the batched estimators.

The reference's eval drivers (eval.py:305-349, eval_shared_f.py:326-360, eval_varying_f.py:325-360) iterate an
`h5py.File` whose keys are

    corr_{img1}_{img2}   [N, >=32] f64   columns 0:2 = keypoints in image 1, 2:4 = keypoints in image 2 (pixels),
                                          then (depth in image 1, depth in image 2) column pairs, one pair per
                                          monocular depth network (column map: utils/data.py:22-46)
    pose_{img1}_{img2}   [3, 4]          ground truth [R | t]
    K_{img}              [3, 3]          intrinsics

and call the estimator once per pair.  Here the same file (any mapping with those keys: an `h5py.File`, or a dict of
arrays in the tests — `h5py` is not part of this image and is imported lazily by `open_h5`) is turned into ONE packed
batch per experiment, which is what `rp_estimate_batch_host` / `api.*_batch` consume.

Driver rules reproduced (`driver=` of read_pairs / evaluate):
  calib    eval.py:333-341          pairs with < 5 matches skipped; rows whose depth is inf / NaN / negative in either
                                    image get depth 1.0 in both (utils/data.py:14-20); PINHOLE cameras from K
  shared   eval_shared_f.py:340-358 < 6 matches skipped; keypoints minus the principal point (pp / 2 with ppbug); when
                                    the two focal sums differ, image-2 keypoints AND K2 are rescaled to image 1's focal;
                                    depths taken as they are (no sanitising)
  varying  eval_varying_f.py:340-352 < 7 matches skipped; keypoints minus the principal point; depths as they are
"""
from collections import namedtuple

import numpy as np

BenchPair = namedtuple("BenchPair", "name1 name2 kp1 kp2 d R_gt t_gt K1 K2")

# utils/data.py:22-46: depth id (the `+k` suffix of an experiment string) -> columns of corr_*
DEPTH_NAMES = {1: "real", 2: "midas", 3: "dpt", 4: "zoe", 5: "depth-anything-v1", 6: "depth-anything-v2",
               7: "depth-pro", 8: "metric3d", 9: "marigold-e2e", 10: "moge", 11: "marigold", 12: "unidepth"}


def depth_indices(depth: int):
    """Columns of `corr_*` holding (depth in image 1, depth in image 2) of depth source `depth` (1..12)."""
    if depth not in DEPTH_NAMES:
        raise ValueError(f"unknown depth id {depth}")
    return (6 + 2 * depth, 7 + 2 * depth)


def invalid_depth_mask(d):
    """utils/data.py:14-20 (`get_valid_depth_mask` returns the INVALID rows): inf, NaN or negative in either column."""
    d = np.asarray(d)
    return np.isinf(d).any(axis=1) | np.isnan(d).any(axis=1) | (d < 0).any(axis=1)


def open_h5(path):
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - h5py is absent from this image
        raise ImportError("reading benchmark files needs h5py; pass any mapping with the same keys instead") from e
    return h5py.File(path, "r")


def pair_names(h5, first=None):
    """eval.py:307-312: pair list from the `corr_` keys; image names end in `_o`."""
    prelim = [k.split("corr_")[1] for k in h5.keys() if "corr_" in k]
    pairs = [(p.split("_o_")[0] + "_o", p.split("_o_")[1]) for p in prelim]
    return pairs[:first] if first is not None else pairs


_MIN_MATCHES = {"calib": 5, "shared": 6, "varying": 7}


def read_pairs(h5, depth=None, first=None, min_matches=None, ppbug=False, driver="calib"):
    """Yields one BenchPair per usable pair, with the depth columns of source `depth` (None: all ones, as the
    reference does for experiments without a `+k` suffix), under the rules of the named eval driver (module
    docstring).  For the focal drivers the keypoints come back already centred."""
    if driver not in _MIN_MATCHES:
        raise ValueError(f"unknown driver {driver!r}")
    min_matches = _MIN_MATCHES[driver] if min_matches is None else min_matches
    cols = depth_indices(depth) if depth is not None else None
    for n1, n2 in pair_names(h5, first):
        data = np.array(h5[f"corr_{n1}_{n2}"], dtype=np.float64)
        if len(data) < min_matches:
            continue
        Rt = np.array(h5[f"pose_{n1}_{n2}"], dtype=np.float64)
        K1, K2 = np.array(h5[f"K_{n1}"], dtype=np.float64), np.array(h5[f"K_{n2}"], dtype=np.float64)
        kp1, kp2 = data[:, :2].copy(), data[:, 2:4].copy()
        if driver == "calib":
            if ppbug:  # eval.py:324-326
                K1, K2 = K1.copy(), K2.copy()
                K1[:2, 2] /= 2
                K2[:2, 2] /= 2
        else:
            pp1, pp2 = K1[:2, 2], K2[:2, 2]
            kp1 -= pp1 / 2 if ppbug else pp1
            kp2 -= pp2 / 2 if ppbug else pp2
            if driver == "shared" and (K1[0, 0] + K1[1, 1]) != (K2[0, 0] + K2[1, 1]):   # eval_shared_f.py:349-351
                ratio = (K1[0, 0] + K1[1, 1]) / (K2[0, 0] + K2[1, 1])
                kp2 *= ratio
                K2 = ratio * K2
        if cols is not None:
            d = data[:, list(cols)].copy()
            if driver == "calib":
                d[invalid_depth_mask(d)] = 1.0
        else:
            d = np.ones_like(kp1)
        yield BenchPair(n1, n2, kp1, kp2, d, Rt[:3, :3], Rt[:, 3], K1, K2)


PackedBatch = namedtuple("PackedBatch", "offsets x1 x2 d1 d2 cams pairs")


def pack(pairs, centre=False):
    """Packed ragged arrays for `Context.estimate_batch_host` / `rp_estimate_batch_host`.

    centre=False: calibrated variants, `cams[p] = (fx1, fy1, cx1, cy1, fx2, fy2, cx2, cy2)` from K (PINHOLE, eval.py:131-132).
    centre=True:  focal variants: pairs read with driver="shared" / "varying" are already centred, no cameras."""
    pairs = list(pairs)
    offsets = np.zeros(len(pairs) + 1, dtype=np.int64)
    for i, p in enumerate(pairs):
        offsets[i + 1] = offsets[i] + len(p.kp1)
    n = int(offsets[-1])
    x1, x2 = np.empty((n, 2)), np.empty((n, 2))
    d1, d2 = np.empty(n), np.empty(n)
    cams = None if centre else np.empty((len(pairs), 8))
    for i, p in enumerate(pairs):
        a, b = offsets[i], offsets[i + 1]
        x1[a:b] = p.kp1
        x2[a:b] = p.kp2
        d1[a:b], d2[a:b] = p.d[:, 0], p.d[:, 1]
        if not centre:
            cams[i] = [p.K1[0, 0], p.K1[1, 1], p.K1[0, 2], p.K1[1, 2], p.K2[0, 0], p.K2[1, 1], p.K2[0, 2], p.K2[1, 2]]
    return PackedBatch(offsets, x1, x2, d1, d2, cams, pairs)


# ---- metrics of utils/eval_utils.py / utils/data.py ------------------------------------------------------------
def rotation_error_deg(R, R_gt):
    """utils/data.py R_err_fun: 2 asin(|R_gt - R|_F / (2 sqrt 2))."""
    s = np.linalg.norm(np.asarray(R_gt) - np.asarray(R)) / (2 * np.sqrt(2))
    return float(np.rad2deg(2 * np.arcsin(max(min(1.0, s), -1.0))))


def translation_error_deg(t, t_gt):
    """utils/data.py t_err_fun: angle between the directions, sign-agnostic."""
    eps = 1e-15
    t = np.asarray(t, dtype=np.float64).ravel()
    t_gt = np.asarray(t_gt, dtype=np.float64).ravel()
    t = t / (np.linalg.norm(t) + eps)
    t_gt = t_gt / (np.linalg.norm(t_gt) + eps)
    loss = max(eps, 1.0 - float(np.sum(t * t_gt)) ** 2)
    return float(np.rad2deg(np.arccos(np.sqrt(1 - loss))))


def pose_maa(pose_errs, max_deg=10):
    """utils/eval_utils.py:41-67: mean over thresholds 1..10 deg of the fraction of pairs below; NaN counts as 180."""
    e = np.array(pose_errs, dtype=np.float64)
    e[np.isnan(e)] = 180.0
    return float(np.mean([np.sum(e < t) / len(e) for t in range(1, max_deg + 1)]))


# Experiment strings of eval.py:93-129 this build can run: the monodepth estimators with the hybrid
# (Sampson + reprojection) LO.  Everything that would silently run as something else is refused:
#   nLO (no local optimisation), GLO (graduated LO), reproj / sym_reproj (non-hybrid cost), reldepth,
#   mad_poselib / madpose, 5p (no depth at all), noshift.
_UNSUPPORTED = ("nLO", "GLO", "sym_reproj", "reldepth", "mad_poselib", "madpose", "5p", "noshift")


def check_experiment(experiment: str) -> None:
    """Raises ValueError for an experiment string outside SURVEY.md §8f row 1 (the fork keys that
    api.make_options would ignore would otherwise report numbers under the wrong name)."""
    name = experiment.split("+")[0]
    bad = [k for k in _UNSUPPORTED if k in name]
    if "reproj" in name and "hybrid" not in name:
        bad.append("reproj (without hybrid)")
    if "hybrid" not in name:
        bad.append("no 'hybrid' LO cost")
    if not ("p3p" in name or "ours" in name):
        bad.append("neither p3p nor ours")
    if bad:
        raise ValueError(f"experiment {experiment!r} is outside this build ({', '.join(bad)}); supported: "
                         "p3p_hybrid_*, 3p_ours_scale_hybrid_*, 3p_ours_shift_scale_hybrid-s_* [+depth]")


def evaluate(h5, experiment, iterations=1000, threshold=2.0, reproj_threshold=16.0, first=None, device=0, driver="calib",
             ppbug=False):
    """One experiment string of eval.py:93-160 / eval_shared_f.py:110-183 / eval_varying_f.py:110-178 over a whole
    benchmark file in ONE batched call.

    Supported strings: the monodepth ones of SURVEY.md §8f row 1 — calib: `p3p_hybrid_ctruncated+k`,
    `3p_ours_scale_hybrid_ctruncated+k`, `3p_ours_shift_scale_hybrid-s_ctruncated+k`; shared / varying:
    `3p_ours_scale_hybrid_[c]truncated+k` (eval_shared_f.py:276-277, eval_varying_f.py:265); `+k` selects the depth
    source.  Returns {"median", "mAA", "errs", "inlier_ratio", "stats"} plus, for the focal drivers, "f_err" (geometric
    mean of the two relative focal errors, eval_shared_f.py:100-102) and "mAA_f"."""
    from . import api
    from . import _native as nv
    check_experiment(experiment)
    depth = int(experiment.split("+")[1]) if "+" in experiment else None
    lo_iterations = 25
    ransac = {"max_iterations": iterations, "min_iterations": iterations, "max_epipolar_error": threshold,
              "max_reproj_error": reproj_threshold, "progressive_sampling": False, "lo_iterations": lo_iterations,
              "use_p3p": "p3p" in experiment, "use_ours": "ours" in experiment, "solver_shift": "shift" in experiment,
              "solver_scale": "scale" in experiment, "optimize_hybrid": "hybrid" in experiment,
              "optimize_shift": "reproj-s" in experiment or "hybrid-s" in experiment, "weight_sampson": 1.0}
    bundle = {"max_iterations": 100, "verbose": False}
    if "truncated" in experiment:
        bundle["loss_type"] = "TRUNCATED"
    if "ctruncated" in experiment:
        bundle["loss_type"] = "TRUNCATED_CAUCHY"
    batch = pack(read_pairs(h5, depth=depth, first=first, driver=driver, ppbug=ppbug), centre=driver != "calib")
    if driver != "calib" and ("shift" in experiment or "p3p" in experiment):
        raise ValueError(f"experiment {experiment!r}: the focal drivers of this build run the scale-only 3-point solver")
    opt = api.make_options(api._fork_ransac(ransac), bundle, focal_variant=driver != "calib")
    variant = {"calib": nv.CALIB_SHIFT if opt.estimate_shift else nv.CALIB, "shared": nv.SHARED, "varying": nv.VARYING}[driver]
    models, stats, _ = api.context(device).estimate_batch_host(variant, batch.offsets, batch.x1, batch.x2, batch.d1,
                                                               batch.d2, batch.cams, opt)
    errs, f_errs = [], []
    for m, p in zip(models, batch.pairs):
        pose = api.CameraPose(q=np.array(m["q"]), t=np.array(m["t"]))
        errs.append(max(rotation_error_deg(pose.R, p.R_gt), translation_error_deg(pose.t, p.t_gt)))
        if driver != "calib":
            f1_gt, f2_gt = (p.K1[0, 0] + p.K1[1, 1]) / 2, (p.K2[0, 0] + p.K2[1, 1]) / 2
            f_errs.append(float(np.sqrt(abs(m["f1"] - f1_gt) / f1_gt * abs(m["f2"] - f2_gt) / f2_gt)))
    out = {"median": float(np.median(errs)) if errs else float("nan"), "mAA": pose_maa(errs) if errs else float("nan"),
           "errs": errs, "inlier_ratio": float(np.mean(stats["inlier_ratio"])) if len(errs) else float("nan"),
           "stats": stats}
    if driver != "calib":
        out["f_err"] = float(np.median(f_errs)) if f_errs else float("nan")
        # utils/eval_utils.py:54-56: mAA_f over thresholds 1..10 % of the focal error (as a fraction * 100)
        fe = np.array(f_errs)
        fe[np.isnan(fe)] = 1.0
        fe = 100.0 * fe
        out["mAA_f"] = float(np.mean([np.sum(fe < t) / len(fe) for t in range(1, 11)])) if len(fe) else float("nan")
    return out
