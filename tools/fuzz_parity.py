#!/usr/bin/env python
"""Randomised end-to-end parity: many small scenes with random sizes, outlier ratios, noise levels, seeds,
iteration counts, thresholds and bundle options (plus deliberately degenerate data: duplicated correspondences,
non-positive depths, all-outlier pairs) through the batched GPU path and, pair by pair, through the oracle.

    python tools/fuzz_parity.py --cases 400            # on a GPU box; prints the mismatches and a summary

`run(ctx, port, cases, seed)` is what tests/test_gpu_fuzz.py calls.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdrp_b200 import _native as nv, synth  # noqa: E402

LOSSES = ["TRIVIAL", "TRUNCATED", "HUBER", "CAUCHY", "TRUNCATED_CAUCHY"]


def random_case(rng, variant, regime="regular"):
    """regime "regular": N >= 21 and a reprojection term (max_reproj_error > 0).  Regime "all" adds what the
    reference itself does not determine uniquely: N of 3..13 (the same triplets are drawn again and again, so
    exactly tied scores decide `refinements` by summation order) and max_reproj_error = 0 (Sampson-only cost: |t|,
    scale and shifts are gauge directions of J^T J + lambda I, rounding noise / lambda decides them)."""
    n = int(rng.choice([21, 34, 60, 150] if regime == "regular" else [3, 4, 5, 8, 13, 21, 34, 60, 150]))
    kw = dict(outlier_ratio=float(rng.choice([0.0, 0.2, 0.5, 0.8])), sigma_px=float(rng.choice([0.0, 0.5, 2.0])),
              depth_noise=float(rng.choice([0.0, 0.01, 0.1])))
    if variant == 1:
        kw.update(shift1=float(rng.uniform(-0.5, 0.5)), shift2=float(rng.uniform(-0.5, 0.5)))
    if variant == 3:
        kw.update(f1=float(rng.uniform(500, 1200)), f2=float(rng.uniform(500, 1200)))
    if variant == 2:
        f = float(rng.uniform(500, 1200))
        kw.update(f1=f, f2=f)
    sc = synth.make_scene(int(rng.integers(0, 10 ** 6)), n, **kw)
    x1, x2, d1, d2 = sc.x1.copy(), sc.x2.copy(), sc.d1.copy(), sc.d2.copy()
    kind = rng.integers(0, 8)
    if kind == 0 and n >= 4:      # duplicated correspondences
        x1[1], x2[1], d1[1], d2[1] = x1[0], x2[0], d1[0], d2[0]
        x1[3], x2[3], d1[3], d2[3] = x1[2], x2[2], d1[2], d2[2]
    elif kind == 1:               # a few non-positive depths
        d1[rng.integers(0, n)] = 0.0
        d2[rng.integers(0, n)] = -1.0
    elif kind == 2:               # no geometry at all
        x2 = np.stack([rng.uniform(0, 1280, n), rng.uniform(0, 960, n)], axis=1)
    opts = dict(iters=int(rng.choice([1, 7, 50, 200])), seed=int(rng.integers(0, 2 ** 31)),
                t_epi=float(rng.choice([0.5, 2.0, 5.0])),
                t_rep=float(rng.choice([8.0, 16.0] if regime == "regular" else [0.0, 8.0, 16.0])),
                loss=str(rng.choice(LOSSES)), bundle_iters=int(rng.choice([0, 5, 100])),
                weight_sampson=float(rng.choice([1.0, 0.5])),
                # a quarter of the cases each: PROSAC sampling (switching back to uniform after 40 samples or never),
                # early termination (min_iterations < max_iterations)
                prosac=bool(rng.integers(0, 4) == 0), max_prosac=int(rng.choice([40, 100000])),
                min_iters=None)
    if rng.integers(0, 4) == 0:
        opts["min_iters"] = int(rng.choice([1, 10, 30]))
        opts["iters"] = int(rng.choice([100, 400]))
    return sc, x1, x2, d1, d2, opts


def run(ctx, port, cases=200, seed=0, verbose=False, regime="regular"):
    """Returns (n_cases, mismatches) where a mismatch is a dict describing the case."""
    rng = np.random.default_rng(seed)
    bad = []
    total = 0
    for variant in (0, 1, 2, 3):
        for _ in range(cases // 4):
            sc, x1, x2, d1, d2, o = random_case(rng, variant, regime)
            n = len(d1)
            opt = nv.default_options()
            opt.max_iterations = o["iters"]
            opt.min_iterations = o["iters"] if o["min_iters"] is None else o["min_iters"]
            opt.progressive_sampling, opt.max_prosac_iterations = int(o["prosac"]), o["max_prosac"]
            opt.max_epipolar_error, opt.max_reproj_error, opt.seed = o["t_epi"], o["t_rep"], o["seed"]
            opt.estimate_shift = int(variant == 1)
            opt.weight_sampson = o["weight_sampson"]
            opt.loss_type = nv.LOSS[o["loss"]]
            opt.loss_scale = 0.5 * o["t_epi"]
            opt.bundle_max_iterations = o["bundle_iters"]
            if variant < 2:
                a1, a2 = x1, x2
                cams = np.array([[sc.f1, sc.f1, 640, 480, sc.f2, sc.f2, 640, 480]])
                cam = ([sc.f1, sc.f1, 640, 480], [sc.f2, sc.f2, 640, 480])
            else:
                a1, a2 = x1 - [640.0, 480.0], x2 - [640.0, 480.0]
                cams, cam = None, (None, None)
            models, stats, masks = ctx.estimate_batch_host(variant, [0, n], a1, a2, d1, d2, cams, opt)
            ro = port.ransac_opt(max_iterations=o["iters"], min_iterations=int(opt.min_iterations),
                                 max_epipolar_error=o["t_epi"], max_reproj_error=o["t_rep"], seed=o["seed"],
                                 estimate_shift=variant == 1, weight_sampson=o["weight_sampson"],
                                 progressive_sampling=o["prosac"], max_prosac_iterations=o["max_prosac"])
            bo = port.bundle_opt(max_iterations=o["bundle_iters"], loss_type=o["loss"], loss_scale=0.5 * o["t_epi"])
            m, st, mask = port.estimate(variant, a1, a2, d1, d2, cam[0], cam[1], ro, bo)
            total += 1
            same_stats = (stats[0]["refinements"], stats[0]["iterations"], stats[0]["num_inliers"]) == \
                (st.refinements, st.iterations, st.num_inliers)
            same_mask = np.array_equal(masks.astype(bool), mask)
            ref = np.r_[np.array(m.q), np.array(m.t), m.scale, m.shift1, m.shift2, m.f1, m.f2]
            got = np.r_[models[0]["q"], models[0]["t"], models[0]["scale"], models[0]["shift1"], models[0]["shift2"],
                        models[0]["f1"], models[0]["f2"]]
            if got[:4] @ ref[:4] < 0:
                got[:4] = -got[:4]
            finite = np.isfinite(ref).all() and np.isfinite(got).all()
            same_model = bool(np.allclose(got, ref, rtol=1e-6, atol=1e-8)) if finite else bool(
                np.array_equal(np.isnan(got), np.isnan(ref)))
            if not (same_stats and same_mask and same_model):  # noqa: E501
                bad.append(dict(variant=variant, n=n, opts=o, stats_gpu=(int(stats[0]["refinements"]), int(stats[0]["iterations"]),
                                                                            int(stats[0]["num_inliers"])),
                                stats_ref=(st.refinements, st.iterations, st.num_inliers), same_mask=same_mask,
                                same_model=same_model, max_diff=float(np.nanmax(np.abs(got - ref))) if finite else float("nan")))
                if verbose:
                    print(bad[-1])
    return total, bad


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=400)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--regime", default="regular", choices=["regular", "all"])
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args()
    from oracle import port
    port.build()
    total, bad = run(nv.Context(0), port, a.cases, a.seed, verbose=not a.quiet, regime=a.regime)
    print(f"{total} cases, {len(bad)} mismatches")
    import collections
    kinds = collections.Counter()
    for b in bad:
        if (b["stats_gpu"][2] != b["stats_ref"][2] or not b["same_mask"]) and a.quiet:
            print(b)
        kind = []
        if b["stats_gpu"][1] != b["stats_ref"][1]:
            kind.append("iterations")
        if b["stats_gpu"][0] != b["stats_ref"][0]:
            kind.append("refinements")
        if b["stats_gpu"][2] != b["stats_ref"][2] or not b["same_mask"]:
            kind.append("inliers/mask")
        if not b["same_model"]:
            kind.append("model")
        kinds[(b["variant"], "+".join(kind), "early-term" if b["opts"]["min_iters"] is not None else "fixed-iters")] += 1
    for k, v in sorted(kinds.items()):
        print(k, v)
