"""GPU tier: every stage of the hot path, called through the C ABI (librepose_b200.so), against the
oracle on the same seeded inputs and against the golden vectors of the reference binary."""
import numpy as np
import pytest

from mdrp_b200 import synth
from util import VARIANT_ID, dedup, model_vec, models_close, same_set, struct_models

pytestmark = pytest.mark.gpu
VARIANTS = ["calib", "calib_shift", "shared", "varying"]


def test_sampler_bit_exact(ctx, port, stages):
    for n, seed in ((3, 0), (7, 5), (2000, 0), (10000, 123456789)):
        ref = stages[f"sampler_n{n}_s{seed}"]
        got = ctx.sample(n, seed, len(ref))
        assert np.array_equal(got, ref)
    # long run against the oracle, including many rejections (tiny n)
    for n, iters in ((4, 3000), (2000, 10000)):
        got = ctx.sample(n, 42, iters)
        st = 42
        for k in range(iters):
            s, st = port.draw_sample(3, n, st)
            assert (s == got[k]).all(), (n, k)


def test_prosac_sampler_bit_exact(ctx, port, prosac_golden):
    """PROSAC branch of the sampler against the reference binary's sequences and, on longer runs (including the
    switch back to uniform sampling after max_prosac_iterations), against the oracle."""
    g = prosac_golden
    for k in [k for k in g.files if k.startswith("sampler_")]:
        n, seed, mp = (int(p[1:]) for p in k.split("_")[1:])
        got = ctx.sample(n, seed, len(g[k]), progressive_sampling=True, max_prosac_iterations=mp)
        assert np.array_equal(got, g[k]), k
    for n, seed, mp, iters in ((3, 1, 100000, 200), (5, 2, 100000, 4000), (2000, 42, 100000, 10000), (2000, 42, 777, 3000),
                               (40, 3, 33, 100), (1000, 8, 1, 64), (1000, 8, 0, 64), (100000, 4, 5000, 6000)):
        got = ctx.sample(n, seed, iters, progressive_sampling=True, max_prosac_iterations=mp)
        ref = port.generate_samples(n, 3, seed, True, mp, iters)
        assert np.array_equal(got, ref), (n, seed, mp)


@pytest.mark.parametrize("variant", VARIANTS)
def test_solvers_vs_golden_and_oracle(ctx, port, stages, variant):
    v = VARIANT_ID[variant]
    x1h, x2h = stages[f"solve_{variant}_x1h"], stages[f"solve_{variant}_x2h"]
    d1, d2 = stages[f"solve_{variant}_d1"], stages[f"solve_{variant}_d2"]
    ref, cnt = stages[f"solve_{variant}_models"], stages[f"solve_{variant}_counts"]
    models, counts = ctx.solve(v, x1h, x2h, d1, d2)
    fn = {"calib": port.solve_calib_scale, "calib_shift": port.solve_calib_shift,
          "shared": port.solve_shared_focal, "varying": port.solve_varying_focal}[variant]
    bad_ref = 0
    for i in range(len(x1h)):
        got = [models[i, k] for k in range(counts[i])]
        # reference (golden): same solution sets to 1e-6 relative, quirk cases tolerated (S2/S3)
        r = [ref[i, k] for k in range(min(cnt[i], 4))]
        bad_ref += not same_set(dedup(r), dedup(got))
        # oracle: same count, same order, 1e-9
        o = fn(x1h[i], x2h[i], d1[i], d2[i])
        assert len(o) == counts[i]
        for a, b in zip(o, got):
            assert np.allclose(model_vec(a), model_vec(b), rtol=1e-9, atol=1e-12, equal_nan=True)
    # S1 / S4: none.  S2: the binary's raw roots are imprecise and its 5-step polish then lands on another root, a
    # duplicate, or a negative-depth solution in 0.45 % of the samples (measured over 40 000 triplets);
    # S3: duplicated pseudo-solutions from complex eigenvalue pairs in 0.14 % (30 000 triplets).  Budgets: 3x those rates.
    limit = {"calib": 0, "varying": 0, "calib_shift": max(2, len(x1h) // 75), "shared": max(1, len(x1h) // 250)}[variant]
    assert bad_ref <= limit


@pytest.mark.parametrize("variant", VARIANTS)
def test_scorer_counts_bit_exact(ctx, port, stages, variant):
    v = VARIANT_ID[variant]
    x1, x2 = stages[f"score_{variant}_x1"], stages[f"score_{variant}_x2"]
    thr2 = float(stages[f"score_{variant}_thr2"][0])
    models = struct_models(stages[f"score_{variant}_models"])
    scores, counts, masks = ctx.score(v, models, x1, x2, thr2, want_masks=True)
    assert np.array_equal(counts, stages[f"score_{variant}_counts"])          # bit exact inlier counts
    assert np.array_equal(masks, stages[f"score_{variant}_masks"])            # and the same inliers
    ref = stages[f"score_{variant}_scores"]
    ok = np.isfinite(ref)
    assert np.all(np.abs(scores[ok] - ref[ok]) <= 1e-12 * np.abs(ref[ok]))


@pytest.mark.parametrize("variant,n_points", [("calib", 1000), ("calib_shift", 2000), ("shared", 2000),
                                              ("varying", 2000), ("calib", 10000), ("calib", 37)])
def test_scorer_many_hypotheses_vs_oracle(ctx, port, variant, n_points):
    """thousands of hypotheses x {1k,2k,10k} points, counts exact."""
    v = VARIANT_ID[variant]
    cfg = {"calib": "cfg1_calib_scale", "calib_shift": "cfg2_calib_shift", "shared": "cfg3_shared_focal",
           "varying": "cfg4_varying_focal"}[variant]
    sc = synth.scene_for(cfg, 21, n=n_points)
    focal = variant in ("shared", "varying")
    if focal:
        x1, x2 = sc.centred()
        ns = (np.linalg.norm(x1, axis=1) + np.linalg.norm(x2, axis=1)).sum() / (np.sqrt(2) * len(x1))
        x1, x2 = x1 / ns, x2 / ns
    else:
        x1, x2, ns = (sc.x1 - synth.PP) / sc.f1, (sc.x2 - synth.PP) / sc.f2, sc.f1
    thr2 = (2.0 / ns) ** 2
    rng = np.random.default_rng(8)
    n_prob = 3000
    idx = np.stack([rng.choice(n_points, 3, replace=False) for _ in range(n_prob)])
    x1h = np.concatenate([x1[idx], np.ones((n_prob, 3, 1))], axis=2)
    x2h = np.concatenate([x2[idx], np.ones((n_prob, 3, 1))], axis=2)
    models, counts = ctx.solve(v, x1h, x2h, sc.d1[idx], sc.d2[idx])
    hyps = np.array([models[i, k] for i in range(n_prob) for k in range(counts[i])], dtype=models.dtype)
    assert len(hyps) > 500
    scores, cnts = ctx.score(v, hyps, x1, x2, thr2)
    step = max(1, len(hyps) // 300)  # the CPU oracle checks a strided subset to stay within seconds
    for i in range(0, len(hyps), step):
        g = hyps[i]
        if focal:
            F = port.fundamental_from_model(port.make_model(g["q"], g["t"], g["scale"], 0, 0, g["f1"], g["f2"]))
            s, c = port.msac_score_F(F, x1, x2, thr2)
        else:
            s, c = port.msac_score_pose(g["q"], g["t"], x1, x2, thr2)
        assert c == cnts[i]
        if s == s:
            assert abs(s - scores[i]) <= 1e-12 * abs(s)
    # size-independent property: the MSAC score is bounded by N*thr^2 and consistent with the count
    assert np.all(scores <= n_points * thr2 * (1 + 1e-12))
    assert np.all(scores >= (n_points - cnts) * thr2 * (1 - 1e-12))


@pytest.mark.parametrize("variant", VARIANTS)
def test_refine_vs_golden(ctx, stages, variant):
    from mdrp_b200 import _native as nv
    v = VARIANT_ID[variant]
    x1, x2 = stages[f"refine_{variant}_x1"], stages[f"refine_{variant}_x2"]
    d1, d2 = stages[f"refine_{variant}_d1"], stages[f"refine_{variant}_d2"]
    thr = float(stages[f"refine_{variant}_thr"][0])
    starts, ends, costs = stages[f"refine_{variant}_start"], stages[f"refine_{variant}_end"], stages[f"refine_{variant}_cost"]
    for sl, loss, iters in ((slice(0, 6), "TRUNCATED", 25), (slice(6, 12), "TRUNCATED_CAUCHY", 100)):
        out, st = ctx.refine(v, struct_models(starts[sl]), x1, x2, d1, d2, (2.0 / 16.0) ** 2, 1.0,
                             nv.bundle_options(max_iterations=iters, loss_type=loss, loss_scale=thr))
        for o, s, e, c in zip(out, st, ends[sl], costs[sl]):
            assert abs(s["initial_cost"] - c[0]) <= 1e-11 * abs(c[0])
            assert s["cost"] <= c[1] * (1 + 1e-9)           # SURVEY §7 step 5 tolerance
            assert models_close(o, e, rtol=1e-6, atol=1e-8)


def test_refine_with_mask_equals_subset(ctx):
    """Final refinement runs on the inlier subset; masking must equal gathering."""
    from mdrp_b200 import _native as nv
    sc = synth.scene_for("cfg2_calib_shift", 5, n=600)
    x1, x2 = (sc.x1 - synth.PP) / sc.f1, (sc.x2 - synth.PP) / sc.f2
    from scipy.spatial.transform import Rotation as Rot
    qq = Rot.from_matrix(sc.R).as_quat()
    start = struct_models(np.r_[qq[3], qq[0], qq[1], qq[2], sc.t, 1.7, 0.3, -0.2, 1, 1][None])
    bo = nv.bundle_options(max_iterations=100, loss_type="TRUNCATED_CAUCHY", loss_scale=1.0 / 800)
    a, sa = ctx.refine(1, start, x1, x2, sc.d1, sc.d2, 1 / 64., 1.0, bo, mask=sc.inlier_mask.astype(np.uint8))
    m = sc.inlier_mask
    b, sb = ctx.refine(1, start, x1[m], x2[m], sc.d1[m], sc.d2[m], 1 / 64., 1.0, bo)
    assert models_close(a[0], b[0], rtol=1e-9, atol=1e-11)
    assert abs(sa[0]["cost"] - sb[0]["cost"]) <= 1e-10 * sb[0]["cost"]
