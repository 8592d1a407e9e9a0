"""The H5 benchmark reader on an in-memory mapping with the reference's key/column layout
(eval.py:305-349, utils/data.py:14-46); h5py itself is not part of this image."""
import numpy as np
import pytest

from mdrp_b200 import benchmark_reader as br, synth


def fake_benchmark(n_pairs=4, n=200, cfg="cfg1_calib_scale"):
    """dict with corr_/pose_/K_ keys like the reference's H5 files; depth source 1 = clean, 2 = with invalid rows."""
    h5, scenes = {}, []
    for i in range(n_pairs):
        sc = synth.scene_for(cfg, 700 + i, n=n if i != 2 else 3)   # pair 2 has < 5 matches: skipped
        a, b = f"img{i:03d}a_o", f"img{i:03d}b_o"
        data = np.zeros((len(sc.d1), 32))
        data[:, :2], data[:, 2:4] = sc.x1, sc.x2
        data[:, 8], data[:, 9] = sc.d1, sc.d2
        data[:, 10], data[:, 11] = sc.d1, sc.d2
        if len(sc.d1) > 10:
            data[0, 10], data[1, 11], data[2, 10], data[3, 11] = np.inf, np.nan, -1.0, -np.inf
        h5[f"corr_{a}_{b}"] = data
        h5[f"pose_{a}_{b}"] = np.c_[sc.R, sc.t]
        h5[f"K_{a}"] = np.array([[sc.f1, 0, 640.0], [0, sc.f1, 480.0], [0, 0, 1]])
        h5[f"K_{b}"] = np.array([[sc.f2, 0, 640.0], [0, sc.f2, 480.0], [0, 0, 1]])
        scenes.append(sc)
    return h5, scenes


def test_depth_column_map_and_invalid_mask():
    assert br.depth_indices(1) == (8, 9) and br.depth_indices(2) == (10, 11) and br.depth_indices(6) == (18, 19)
    assert br.depth_indices(10) == (26, 27) and br.depth_indices(12) == (30, 31)
    with pytest.raises(ValueError):
        br.depth_indices(13)
    d = np.array([[1.0, 2.0], [np.inf, 1.0], [1.0, np.nan], [-0.5, 1.0], [0.0, 0.0], [1.0, -np.inf]])
    assert br.invalid_depth_mask(d).tolist() == [False, True, True, True, False, True]


def test_read_and_pack():
    h5, scenes = fake_benchmark()
    assert br.pair_names(h5) == [(f"img{i:03d}a_o", f"img{i:03d}b_o") for i in range(4)]
    assert len(br.pair_names(h5, first=2)) == 2
    pairs = list(br.read_pairs(h5, depth=2))
    assert [p.name1 for p in pairs] == ["img000a_o", "img001a_o", "img003a_o"]   # the 3-match pair is skipped
    p = pairs[0]
    assert np.array_equal(p.kp1, scenes[0].x1) and np.array_equal(p.kp2, scenes[0].x2)
    assert (p.d[:4] == 1.0).all() and np.array_equal(p.d[4:, 0], scenes[0].d1[4:])   # invalid rows -> depth 1 in both
    assert np.array_equal(p.R_gt, scenes[0].R) and np.array_equal(p.t_gt, scenes[0].t)
    ones = list(br.read_pairs(h5, depth=None))
    assert (ones[0].d == 1.0).all() and ones[0].d.shape == (200, 2)
    b = br.pack(pairs)
    assert b.offsets.tolist() == [0, 200, 400, 600] and b.x1.shape == (600, 2) and b.cams.shape == (3, 8)
    assert b.cams[1].tolist() == [scenes[1].f1, scenes[1].f1, 640.0, 480.0, scenes[1].f2, scenes[1].f2, 640.0, 480.0]
    assert np.array_equal(b.d2[200:400], pairs[1].d[:, 1])
    # focal drivers: centred keypoints, depths as they are, their own minimum match counts
    sh = list(br.read_pairs(h5, depth=2, driver="shared"))
    c = br.pack(sh, centre=True)
    assert c.cams is None and np.allclose(c.x1[:200], scenes[0].x1 - [640.0, 480.0])
    assert np.isinf(sh[0].d[0, 0]) and np.isnan(sh[0].d[1, 1])          # eval_shared_f.py:353-356: no sanitising
    assert np.allclose(next(br.read_pairs(h5, depth=1, driver="varying", ppbug=True)).kp1, scenes[0].x1 - [320.0, 240.0])
    halved = next(br.read_pairs(h5, depth=1, ppbug=True))
    assert halved.K1[0, 2] == 320.0 and h5["K_img000a_o"][0, 2] == 640.0   # the file's matrix is not modified


def test_focal_driver_rules():
    """eval_shared_f.py:340-351 / eval_varying_f.py:340: minimum match counts and the focal rescale of image 2."""
    h5, scenes = fake_benchmark(n_pairs=2, n=50, cfg="cfg4_varying_focal")   # f1 = 700, f2 = 900
    h5["corr_short_o_short2_o"] = np.zeros((6, 32))
    h5["pose_short_o_short2_o"] = np.eye(3, 4)
    h5["K_short_o"] = h5["K_short2_o"] = np.array([[800.0, 0, 640], [0, 800.0, 480], [0, 0, 1]])
    assert len(list(br.read_pairs(h5, driver="calib"))) == 3 and len(list(br.read_pairs(h5, driver="shared"))) == 3
    assert len(list(br.read_pairs(h5, driver="varying"))) == 2            # 6 matches < 7
    sh = next(br.read_pairs(h5, depth=1, driver="shared"))
    assert np.allclose(sh.kp2, (scenes[0].x2 - [640.0, 480.0]) * 700.0 / 900.0) and abs(sh.K2[0, 0] - 700.0) < 1e-9
    va = next(br.read_pairs(h5, depth=1, driver="varying"))
    assert np.allclose(va.kp2, scenes[0].x2 - [640.0, 480.0]) and va.K2[0, 0] == 900.0
    with pytest.raises(ValueError):
        list(br.read_pairs(h5, driver="nope"))


def test_metrics_match_reference_definitions():
    R = synth.rodrigues(np.array([0.0, 0.0, 1.0]), np.deg2rad(10.0)) if hasattr(synth, "rodrigues") else None
    if R is not None:
        assert abs(br.rotation_error_deg(R, np.eye(3)) - 10.0) < 1e-9
    assert abs(br.translation_error_deg([1, 0, 0], [-2, 0, 0])) < 1e-5          # sign-agnostic
    assert abs(br.translation_error_deg([1, 0, 0], [1, 1, 0]) - 45.0) < 1e-9
    assert abs(br.pose_maa([0.5, 1.5, 20.0, np.nan]) - np.mean([(1 + (t > 1.5)) / 4 for t in range(1, 11)])) < 1e-12


@pytest.mark.gpu
def test_evaluate_equals_per_pair_calls(ctx):
    """The batched benchmark evaluation gives, pair by pair, what the reference-style per-pair call gives."""
    from mdrp_b200 import api
    h5, scenes = fake_benchmark(n_pairs=5, n=300, cfg="cfg2_calib_shift")
    exp = "3p_ours_shift_scale_hybrid-s_ctruncated+1"
    res = br.evaluate(h5, exp, iterations=300)
    assert len(res["errs"]) == 4 and res["mAA"] > 0.9 and res["median"] < 0.5
    pairs = list(br.read_pairs(h5, depth=1))
    for p, e in zip(pairs, res["errs"]):
        cam = lambda K: {"model": "PINHOLE", "width": -1, "height": -1, "params": [K[0, 0], K[1, 1], K[0, 2], K[1, 2]]}
        ransac = api._fork_ransac({"max_iterations": 300, "min_iterations": 300, "max_epipolar_error": 2.0,
                                   "max_reproj_error": 16.0, "use_ours": True, "solver_shift": True, "use_p3p": False})
        pose, info = api.estimate_relative_pose_w_mono_depth(p.kp1, p.kp2, p.d, cam(p.K1), cam(p.K2), ransac,
                                                             {"loss_type": "TRUNCATED_CAUCHY", "max_iterations": 100})
        e1 = max(br.rotation_error_deg(pose.R, p.R_gt), br.translation_error_deg(pose.t, p.t_gt))
        assert abs(e1 - e) < 1e-9


def test_unsupported_experiment_strings_are_refused():
    """eval.py:93-129 knows many more experiments than this build runs; none may run under a wrong name."""
    from mdrp_b200 import benchmark_reader as br
    for ok in ("p3p_hybrid_ctruncated+3", "3p_ours_scale_hybrid_ctruncated", "3p_ours_shift_scale_hybrid-s_ctruncated+11"):
        br.check_experiment(ok)
    for bad in ("p3p_nLO_hybrid_ctruncated", "3p_ours_scale_GLO_hybrid", "3p_ours_scale_reproj_ctruncated", "5p_ctruncated",
                "madpose+3", "mad_poselib_shift_scale_hybrid", "3p_reldepth_hybrid", "3p_ours_scale_sym_reproj"):
        with pytest.raises(ValueError):
            br.check_experiment(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("driver,cfg", [("shared", "cfg3_shared_focal"), ("varying", "cfg4_varying_focal")])
def test_evaluate_focal_drivers_equal_per_pair_calls(ctx, driver, cfg):
    """eval_shared_f.py:177 / eval_varying_f.py:168 through the fork names, pair by pair, against the batched evaluation."""
    from mdrp_b200 import api
    h5, scenes = fake_benchmark(n_pairs=5, n=400, cfg=cfg)
    exp = "3p_ours_scale_hybrid_ctruncated+1"
    res = br.evaluate(h5, exp, iterations=500, driver=driver)
    assert len(res["errs"]) == 4 and res["mAA"] > 0.7 and res["f_err"] < 0.05 and 0.0 <= res["mAA_f"] <= 1.0
    fn = api.estimate_shared_focal_monodepth_relative_pose if driver == "shared" else api.estimate_varying_focal_monodepth_relative_pose
    ransac = {"max_iterations": 500, "min_iterations": 500, "max_epipolar_error": 2.0, "max_reproj_error": 16.0,
              "use_ours": True, "solver_scale": True, "solver_shift": False, "use_p3p": False, "optimize_hybrid": True}
    for p, e in zip(br.read_pairs(h5, depth=1, driver=driver), res["errs"]):
        pair, info = fn(p.kp1, p.kp2, p.d, ransac, {"loss_type": "TRUNCATED_CAUCHY", "max_iterations": 100})
        e1 = max(br.rotation_error_deg(pair.pose.R, p.R_gt), br.translation_error_deg(pair.pose.t, p.t_gt))
        assert abs(e1 - e) < 1e-9
    with pytest.raises(ValueError):
        br.evaluate(h5, "3p_ours_shift_scale_hybrid-s_ctruncated+1", driver=driver)
