"""CPU tier: randomized comparison of the C restatement with the reference binary itself (the
PoseLib wheel unpacked under oracle/_ref, called stage by stage through its exported C++
symbols).  Skipped on boxes without oracle/_ref; the golden fixtures cover those."""
import numpy as np
import pytest

from mdrp_b200 import synth
from util import dedup, models_close, same_set


def _norm(sc, focal):
    if not focal:
        return (sc.x1 - synth.PP) / sc.f1, (sc.x2 - synth.PP) / sc.f2, sc.f1
    x1, x2 = sc.centred()
    ns = (np.linalg.norm(x1, axis=1) + np.linalg.norm(x2, axis=1)).sum() / (np.sqrt(2) * len(x1))
    return x1 / ns, x2 / ns, ns


def test_sampler(ref, port):
    a = b = 987654321
    for n in (3, 4, 50, 2000):
        for _ in range(500):
            s1, a = ref.draw_sample(3, n, a)
            s2, b = port.draw_sample(3, n, b)
            assert (s1 == s2).all() and a == b
    v1, s1 = ref.random_int(0)
    v2, s2 = port.random_int(0)
    assert (v1, s1) == (v2, s2)


def test_prosac_sampler(ref, port):
    """RandomSampler::generate_sample so@0x4f8970 / initialize_prosac so@0x4f8a20 on a hand-built sampler struct."""
    for n, mp, iters, seed in ((3, 100000, 50, 0), (4, 1000, 300, 5), (10, 100000, 500, 3), (50, 100000, 3000, 0),
                               (300, 200, 600, 1), (2000, 100000, 5000, 7), (100000, 100000, 1500, 9)):
        a, growth = ref.generate_samples(n, 3, seed, True, mp, iters)
        b = port.generate_samples(n, 3, seed, True, mp, iters)
        assert np.array_equal(a, b), (n, mp)
        assert np.array_equal(growth, port.prosac_growth(n, 3, mp)), (n, mp)
    a, _ = ref.generate_samples(100, 3, 5, False, 100000, 500)
    assert np.array_equal(a, port.generate_samples(100, 3, 5, False, 100000, 500))


def test_scorer_and_essential_bit_exact(ref, port):
    rng = np.random.default_rng(1)
    sc = synth.scene_for("cfg2_calib_shift", 3, n=700)
    x1, x2, _ = _norm(sc, False)
    thr2 = (2.0 / 800) ** 2
    for k in range(60):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        t = rng.normal(size=3)
        assert (ref.essential_from_motion(q, t) == port.essential_from_motion(q, t)).all()
        assert ref.msac_score_pose(q, t, x1, x2, thr2) == port.msac_score_pose(q, t, x1, x2, thr2)
        F = rng.normal(size=(3, 3))
        assert ref.msac_score_F(F, x1, x2, thr2) == port.msac_score_F(F, x1, x2, thr2)
        assert (ref.get_inliers_F(F, x1, x2, thr2) == port.get_inliers_F(F, x1, x2, thr2)).all()


@pytest.mark.parametrize("variant", ["calib", "calib_shift", "shared", "varying"])
def test_solvers(ref, port, variant):
    pl = ref.poselib()
    rng = np.random.default_rng(5)
    cfg = {"calib": "cfg1_calib_scale", "calib_shift": "cfg2_calib_shift", "shared": "cfg3_shared_focal",
           "varying": "cfg4_varying_focal"}[variant]
    sc = synth.scene_for(cfg, 2, n=600)
    x1, x2, _ = _norm(sc, variant in ("shared", "varying"))
    bad = 0
    n_trials = 600
    for _ in range(n_trials):
        idx = rng.choice(len(x1), 3, replace=False)
        x1h, x2h = np.c_[x1[idx], np.ones(3)], np.c_[x2[idx], np.ones(3)]
        d1, d2 = sc.d1[idx], sc.d2[idx]
        if variant == "calib":
            X = x1h * d1[:, None]
            b = x2h / np.linalg.norm(x2h, axis=1)[:, None]
            r = [np.r_[np.array(p.q).ravel(), np.array(p.t).ravel()] for p in pl.p3p(b, X)]
            g = [np.r_[q, t] for q, t in port.p3p(b, X)]
            ok = len(r) == len(g) and all(np.allclose(a, c, rtol=1e-9, atol=1e-12, equal_nan=True) for a, c in zip(r, g))
        elif variant == "calib_shift":
            r = [np.r_[np.array(m.pose.q).ravel(), np.array(m.pose.t).ravel(), m.scale, m.shift1, m.shift2, 1, 1]
                 for m in pl.monodepth_pose_3pt(x1h, x2h, list(d1), list(d2))]
            ok = same_set(dedup(r), dedup(port.solve_calib_shift(x1h, x2h, d1, d2)))
        else:
            fn = pl.shared_focal_monodepth_pose_3pt if variant == "shared" else pl.varying_focal_monodepth_pose_4pt
            r = [np.r_[np.array(m.geometry.pose.q).ravel(), np.array(m.geometry.pose.t).ravel(), m.geometry.scale, 0, 0,
                       m.camera1.focal(), m.camera2.focal()] for m in fn(x1h, x2h, list(d1), list(d2))]
            mine = port.solve_shared_focal(x1h, x2h, d1, d2) if variant == "shared" else port.solve_varying_focal(x1h, x2h, d1, d2)
            ok = same_set(dedup(r), dedup(mine))
        bad += not ok
    # measured quirk rates of the binary's S2 / S3: 0.45 % / 0.14 % of the samples; budgets 3x
    limit = {"calib": 0, "varying": 0, "calib_shift": n_trials // 75, "shared": n_trials // 250}[variant]
    assert bad <= limit, f"{bad}/{n_trials} solution sets differ"


@pytest.mark.parametrize("variant,cfg", [(0, "cfg1_calib_scale"), (1, "cfg2_calib_shift"), (2, "cfg3_shared_focal"),
                                         (3, "cfg4_varying_focal"), (0, "hard_calib")])
@pytest.mark.parametrize("prosac", [False, True])
def test_end_to_end(ref, port, variant, cfg, prosac):
    pl = ref.poselib()
    for idx in range(2):
        sc = synth.scene_for(cfg, 40 + idx, n=500)
        iters = 400
        mp = 100000 if idx == 0 else 250  # second scene: PROSAC switches to uniform sampling mid-run
        ro = {"max_iterations": iters, "min_iterations": iters, "max_epipolar_error": 2.0, "max_reproj_error": 16.0,
              "seed": 3, "monodepth_estimate_shift": variant == 1, "progressive_sampling": prosac,
              "max_prosac_iterations": mp}
        bo = {"loss_type": "TRUNCATED_CAUCHY"}
        rop = port.ransac_opt(max_iterations=iters, min_iterations=iters, max_epipolar_error=2.0, max_reproj_error=16.0,
                              seed=3, estimate_shift=variant == 1, progressive_sampling=prosac, max_prosac_iterations=mp)
        bop = port.bundle_opt(loss_type="TRUNCATED_CAUCHY", loss_scale=1.0)
        if variant < 2:
            c1, c2 = sc.camera_dicts()
            g, info = pl.estimate_monodepth_relative_pose(sc.x1, sc.x2, sc.d1, sc.d2, c1, c2, ro, bo)
            refm = np.r_[np.array(g.pose.q).ravel(), np.array(g.pose.t).ravel(), g.scale, g.shift1, g.shift2, 1, 1]
            m, st, mask = port.estimate(variant, sc.x1, sc.x2, sc.d1, sc.d2, [sc.f1, sc.f1, 640, 480],
                                        [sc.f2, sc.f2, 640, 480], rop, bop)
        else:
            x1, x2 = sc.centred()
            fn = pl.estimate_monodepth_shared_focal_relative_pose if variant == 2 else pl.estimate_monodepth_varying_focal_relative_pose
            g, info = fn(x1, x2, sc.d1, sc.d2, ro, bo)
            refm = np.r_[np.array(g.geometry.pose.q).ravel(), np.array(g.geometry.pose.t).ravel(), g.geometry.scale, 0, 0,
                         g.camera1.focal(), g.camera2.focal()]
            m, st, mask = port.estimate(variant, x1, x2, sc.d1, sc.d2, None, None, rop, bop)
        assert (st.refinements, st.iterations, st.num_inliers) == (info["refinements"], info["iterations"], info["num_inliers"])
        assert (mask == np.array(info["inliers"])).all()
        assert models_close(m, refm, rtol=1e-8, atol=1e-10)
