"""GPU tier: the batched estimators through the C ABI against the reference's golden outputs, the
oracle on the same inputs, the reference's edge-case behaviour and size-independent properties at
BASELINE.json's full sizes."""
import sys

import numpy as np
import pytest

from mdrp_b200 import _native as nv, api, synth
from util import extra_cases, maa, models_close, rot_err_deg, trans_err_deg

pytestmark = pytest.mark.gpu
DBL_MAX = sys.float_info.max


def _options(iters, shift=False, seed=0, min_iters=None, loss_scale=1.0):
    o = nv.default_options()
    o.max_iterations = iters
    o.min_iterations = iters if min_iters is None else min_iters
    o.max_epipolar_error, o.max_reproj_error, o.seed = 2.0, 16.0, seed
    o.estimate_shift = int(shift)
    o.loss_type = nv.LOSS["TRUNCATED_CAUCHY"]
    o.loss_scale = loss_scale
    return o


def _batch(cfg, indices, n=None):
    c = synth.CONFIGS[cfg]
    scs = [synth.scene_for(cfg, i, n=n) for i in indices]
    variant = {"calib": 1 if c["shift"] else 0, "shared": 2, "varying": 3}[c["variant"]]
    offs = np.r_[0, np.cumsum([len(s.d1) for s in scs])]
    if variant < 2:
        x1, x2 = np.concatenate([s.x1 for s in scs]), np.concatenate([s.x2 for s in scs])
        cams = np.array([[s.f1, s.f1, 640, 480, s.f2, s.f2, 640, 480] for s in scs], dtype=np.float64)
    else:
        x1, x2 = np.concatenate([s.centred()[0] for s in scs]), np.concatenate([s.centred()[1] for s in scs])
        cams = None
    return scs, variant, offs, x1, x2, np.concatenate([s.d1 for s in scs]), np.concatenate([s.d2 for s in scs]), cams


def test_golden_end_to_end(ctx, e2e_golden):
    """Outputs of the reference binary on committed inputs: stats, mask and model must agree."""
    g = e2e_golden
    keys = sorted(k[:-6] for k in g.files if k.endswith("_model"))
    for key in keys:
        name = key.split("_cfg")[0].split("_hard")[0]
        variant = {"calib": 0, "calib_shift": 1, "shared": 2, "varying": 3, "calib_default_iters": 0}[name]
        iters = int(g[key + "_iters"][0])
        idx = int(key[-1])
        o = _options(iters if iters > 0 else 5000, shift=variant == 1, seed=idx, min_iters=None if iters > 0 else 100)
        n = len(g[key + "_d1"])
        f1, f2 = g[key + "_f"]
        cams = np.array([[f1, f1, 640, 480, f2, f2, 640, 480]]) if variant < 2 else None
        models, stats, masks = ctx.estimate_batch_host(variant, [0, n], g[key + "_x1"], g[key + "_x2"], g[key + "_d1"],
                                                       g[key + "_d2"], cams, o)
        ref = g[key + "_stats"]
        assert (stats[0]["refinements"], stats[0]["iterations"], stats[0]["num_inliers"]) == tuple(ref), key
        assert abs(stats[0]["model_score"] - g[key + "_fstats"][1]) <= 1e-11 * g[key + "_fstats"][1], key
        assert abs(stats[0]["inlier_ratio"] - g[key + "_fstats"][0]) <= 1e-12, key
        assert np.array_equal(masks, g[key + "_mask"]), key
        assert models_close(models[0], g[key + "_model"], rtol=1e-6, atol=1e-8), key  # north_star: 1e-6 relative


def test_prosac_golden_end_to_end(ctx, prosac_golden):
    """progressive_sampling=True: outputs of the reference binary on committed, quality-sorted inputs."""
    g = prosac_golden
    keys = sorted(k[:-6] for k in g.files if k.endswith("_model"))
    assert len(keys) >= 12
    for key in keys:
        name = key.split("_cfg")[0].split("_hard")[0]
        variant = {"calib": 0, "calib_shift": 1, "shared": 2, "varying": 3}[name]
        iters, mp = (int(v) for v in g[key + "_opts"])
        o = _options(iters, shift=variant == 1, seed=int(key[-1]))
        o.progressive_sampling, o.max_prosac_iterations = 1, mp
        n = len(g[key + "_d1"])
        f1, f2 = g[key + "_f"]
        cams = np.array([[f1, f1, 640, 480, f2, f2, 640, 480]]) if variant < 2 else None
        models, stats, masks = ctx.estimate_batch_host(variant, [0, n], g[key + "_x1"], g[key + "_x2"], g[key + "_d1"],
                                                       g[key + "_d2"], cams, o)
        ref = g[key + "_stats"]
        assert (stats[0]["refinements"], stats[0]["iterations"], stats[0]["num_inliers"]) == tuple(ref), key
        assert abs(stats[0]["model_score"] - g[key + "_fstats"][1]) <= 1e-11 * g[key + "_fstats"][1], key
        assert np.array_equal(masks, g[key + "_mask"]), key
        assert models_close(models[0], g[key + "_model"], rtol=1e-6, atol=1e-8), key


def test_prosac_batch_vs_oracle_with_early_termination(ctx, port):
    """PROSAC in a ragged batch with the default early-termination mode (min_iterations < max_iterations)."""
    sizes = [300, 64, 1000, 3, 517]
    scs = [synth.scene_for("cfg1_calib_scale", 90 + i, n=n) for i, n in enumerate(sizes)]
    offs = np.r_[0, np.cumsum(sizes)]
    x1, x2 = np.concatenate([s.x1 for s in scs]), np.concatenate([s.x2 for s in scs])
    d1, d2 = np.concatenate([s.d1 for s in scs]), np.concatenate([s.d2 for s in scs])
    cams = np.array([[s.f1, s.f1, 640, 480, s.f2, s.f2, 640, 480] for s in scs], dtype=np.float64)
    o = _options(3000, seed=5, min_iters=100)
    o.progressive_sampling, o.max_prosac_iterations = 1, 400
    models, stats, masks = ctx.estimate_batch_host(0, offs, x1, x2, d1, d2, cams, o)
    ro = port.ransac_opt(max_iterations=3000, min_iterations=100, max_epipolar_error=2.0, max_reproj_error=16.0, seed=5,
                         progressive_sampling=True, max_prosac_iterations=400)
    bo = port.bundle_opt(loss_type="TRUNCATED_CAUCHY", loss_scale=1.0)
    for i, s in enumerate(scs):
        m, st, mask = port.estimate(0, s.x1, s.x2, s.d1, s.d2, [s.f1, s.f1, 640, 480], [s.f2, s.f2, 640, 480], ro, bo)
        assert (stats[i]["refinements"], stats[i]["iterations"], stats[i]["num_inliers"]) == \
            (st.refinements, st.iterations, st.num_inliers), i
        assert np.array_equal(masks[offs[i]:offs[i + 1]].astype(bool), mask), i
        assert models_close(models[i], m, rtol=1e-6, atol=1e-8), i


@pytest.mark.parametrize("cfg", ["cfg1_calib_scale", "cfg2_calib_shift", "cfg3_shared_focal", "cfg4_varying_focal"])
def test_batch_vs_oracle(ctx, port, cfg):
    """A ragged batch (different N per pair) against the oracle run pair by pair."""
    c = synth.CONFIGS[cfg]
    sizes = [300, 64, 1000, 3, 517, 129]
    scs = [synth.scene_for(cfg, 60 + i, n=n) for i, n in enumerate(sizes)]
    variant = {"calib": 1 if c["shift"] else 0, "shared": 2, "varying": 3}[c["variant"]]
    offs = np.r_[0, np.cumsum(sizes)]
    if variant < 2:
        x1, x2 = np.concatenate([s.x1 for s in scs]), np.concatenate([s.x2 for s in scs])
        cams = np.array([[s.f1, s.f1, 640, 480, s.f2, s.f2, 640, 480] for s in scs], dtype=np.float64)
    else:
        x1, x2 = np.concatenate([s.centred()[0] for s in scs]), np.concatenate([s.centred()[1] for s in scs])
        cams = None
    d1, d2 = np.concatenate([s.d1 for s in scs]), np.concatenate([s.d2 for s in scs])
    iters = 500
    models, stats, masks = ctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, _options(iters, c["shift"]))
    rop = port.ransac_opt(max_iterations=iters, min_iterations=iters, max_epipolar_error=2.0, max_reproj_error=16.0,
                          seed=0, estimate_shift=c["shift"])
    bop = port.bundle_opt(loss_type="TRUNCATED_CAUCHY", loss_scale=1.0)
    for i, s in enumerate(scs):
        sl = slice(offs[i], offs[i + 1])
        cam = ([s.f1, s.f1, 640, 480], [s.f2, s.f2, 640, 480]) if variant < 2 else (None, None)
        m, st, mk = port.estimate(variant, x1[sl], x2[sl], d1[sl], d2[sl], cam[0], cam[1], rop, bop)
        if sizes[i] == 3:
            # degenerate: every sample is the same triple, all MSAC scores are ~1e-31 rounding noise, so
            # which of them "improves" is decided by the summation order (tie rule)
            assert (st.iterations, st.num_inliers) == (stats[i]["iterations"], stats[i]["num_inliers"]), (cfg, i)
        else:
            assert (st.refinements, st.iterations, st.num_inliers) == (
                stats[i]["refinements"], stats[i]["iterations"], stats[i]["num_inliers"]), (cfg, i)
        assert np.array_equal(mk, masks[sl].astype(bool)), (cfg, i)
        if sizes[i] > 3:
            assert models_close(models[i], m, rtol=1e-6, atol=1e-8), (cfg, i)


@pytest.mark.parametrize("cfg", ["cfg1_calib_scale", "cfg3_shared_focal"])
def test_large_pair_vs_oracle(ctx, port, cfg):
    """Pairs around the capacity of the LM kernel's shared-memory inlier list (16384): 9000 correspondences use the
    list, 17000 take the mask-on-the-fly path; both must agree with the oracle."""
    c = synth.CONFIGS[cfg]
    scs, variant, offs, x1, x2, d1, d2, cams = _batch(cfg, [77, 78], n=9000)
    big = synth.scene_for(cfg, 79, n=17000)
    scs.append(big)
    bx1, bx2 = (big.x1, big.x2) if variant < 2 else big.centred()
    x1, x2 = np.concatenate([x1, bx1]), np.concatenate([x2, bx2])
    d1, d2 = np.concatenate([d1, big.d1]), np.concatenate([d2, big.d2])
    offs = np.r_[offs, offs[-1] + 17000]
    if cams is not None:
        cams = np.concatenate([cams, [[big.f1, big.f1, 640, 480, big.f2, big.f2, 640, 480]]])
    iters = 200
    models, stats, masks = ctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, _options(iters, c["shift"]))
    ro = port.ransac_opt(max_iterations=iters, min_iterations=iters, max_epipolar_error=2.0, max_reproj_error=16.0,
                         seed=0, estimate_shift=c["shift"])
    bo = port.bundle_opt(loss_type="TRUNCATED_CAUCHY", loss_scale=1.0)
    for i, s in enumerate(scs):
        sl = slice(offs[i], offs[i + 1])
        cam = ([s.f1, s.f1, 640, 480], [s.f2, s.f2, 640, 480]) if variant < 2 else (None, None)
        m, st, mask = port.estimate(variant, x1[sl], x2[sl], d1[sl], d2[sl], cam[0], cam[1], ro, bo)
        assert (stats[i]["refinements"], stats[i]["iterations"], stats[i]["num_inliers"]) == \
            (st.refinements, st.iterations, st.num_inliers), i
        assert np.array_equal(masks[sl].astype(bool), mask), i
        assert models_close(models[i], m, rtol=1e-6, atol=1e-8), i


def test_edge_cases_match_reference_behaviour(ctx):
    """SURVEY §8b 'Errors': N<3 -> identity, iterations 0, model_score DBL_MAX, all-False mask; empty
    batch; a pair with zero correspondences inside a batch."""
    o = _options(200)
    cams = np.tile(np.array([800., 800, 640, 480, 800, 800, 640, 480]), (3, 1))
    sc = synth.scene_for("cfg1_calib_scale", 0, n=200)
    x1 = np.concatenate([sc.x1[:2], sc.x1])
    x2 = np.concatenate([sc.x2[:2], sc.x2])
    d1, d2 = np.concatenate([sc.d1[:2], sc.d1]), np.concatenate([sc.d2[:2], sc.d2])
    models, stats, masks = ctx.estimate_batch_host(0, [0, 2, 2, 202], x1, x2, d1, d2, cams, o)
    for i in (0, 1):  # N=2 and N=0
        assert np.array_equal(models[i]["q"], [1, 0, 0, 0]) and np.array_equal(models[i]["t"], [0, 0, 0])
        assert models[i]["scale"] == 1.0
        assert stats[i]["iterations"] == 0 and stats[i]["refinements"] == 0 and stats[i]["num_inliers"] == 0
        assert stats[i]["model_score"] == DBL_MAX
    assert masks[:2].sum() == 0
    assert stats[2]["num_inliers"] > 100 and masks[2:].sum() == stats[2]["num_inliers"]
    m0, s0, k0 = ctx.estimate_batch_host(0, [0], np.zeros((0, 2)), np.zeros((0, 2)), np.zeros(0), np.zeros(0),
                                         np.zeros((0, 8)), o)
    assert len(m0) == 0 and len(s0) == 0 and len(k0) == 0
    with pytest.raises(ValueError):
        ctx.estimate_batch_host(0, [0, 5], np.zeros((4, 2)), np.zeros((5, 2)), np.zeros(5), np.zeros(5), cams[:1], o)


def test_early_termination_matches_reference_rule(ctx, port):
    """Default options (min 1000 / max 100000): the loop stops at the first it > min_iterations with
    it > dynamic_max_iter — iteration 1001 on easy data (demo/reposed_demo.ipynb:659)."""
    for cfg, n, min_it, max_it in (("cfg1_calib_scale", 150, 1000, 100000), ("hard_calib", 200, 50, 20000)):
        sc = synth.scene_for(cfg, 77, n=n)
        o = _options(max_it, min_iters=min_it)
        cams = np.array([[800., 800, 640, 480, 800, 800, 640, 480]])
        models, stats, masks = ctx.estimate_batch_host(0, [0, n], sc.x1, sc.x2, sc.d1, sc.d2, cams, o)
        rop = port.ransac_opt(max_iterations=max_it, min_iterations=min_it, max_epipolar_error=2.0, max_reproj_error=16.0)
        m, st, mk = port.estimate(0, sc.x1, sc.x2, sc.d1, sc.d2, [800, 800, 640, 480], [800, 800, 640, 480], rop,
                                  port.bundle_opt(loss_type="TRUNCATED_CAUCHY"))
        assert stats[0]["iterations"] == st.iterations, cfg
        assert (stats[0]["refinements"], stats[0]["num_inliers"]) == (st.refinements, st.num_inliers), cfg
        assert np.array_equal(mk, masks.astype(bool))
        assert models_close(models[0], m, rtol=1e-6, atol=1e-8)
    assert True


def test_result_independent_of_batch_composition(ctx):
    """Determinism per pair: seed per pair = opt.seed as in the reference, so a pair's result must
    not depend on what else is in the batch, nor on repetition."""
    scs, variant, offs, x1, x2, d1, d2, cams = _batch("cfg2_calib_shift", range(12), n=400)
    o = _options(1000, shift=True)
    a = ctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    b = ctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    for u, v in zip(a, b):
        assert u.tobytes() == v.tobytes()
    for i in (0, 5, 11):
        sl = slice(offs[i], offs[i + 1])
        m, s, k = ctx.estimate_batch_host(variant, [0, offs[i + 1] - offs[i]], x1[sl], x2[sl], d1[sl], d2[sl],
                                          cams[i:i + 1], o)
        assert m[0].tobytes() == a[0][i].tobytes() and s[0].tobytes() == a[1][i].tobytes()
        assert np.array_equal(k, a[2][sl])


@pytest.mark.parametrize("cfg", ["cfg2_calib_shift", "cfg3_shared_focal", "cfg4_varying_focal", "cfg5_roma_calib"])
def test_full_size_properties(ctx, cfg):
    """BASELINE.json full sizes (2k matches x 10k iterations; 10k matches): size-independent checks —
    recovered pose close to ground truth, mask == get_inliers of the RANSAC model (count identity),
    mask mostly the true inliers, score bound."""
    c = synth.CONFIGS[cfg]
    scs, variant, offs, x1, x2, d1, d2, cams = _batch(cfg, range(200, 216))
    o = _options(c["iters"], shift=c["shift"])
    models, stats, masks = ctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    for i, s in enumerate(scs):
        sl = slice(offs[i], offs[i + 1])
        assert stats[i]["iterations"] == c["iters"]
        assert masks[sl].sum() == stats[i]["num_inliers"]
        assert abs(stats[i]["inlier_ratio"] - stats[i]["num_inliers"] / len(s.d1)) < 1e-12
        assert rot_err_deg(models[i]["q"], s.R) < 0.1
        assert trans_err_deg(models[i]["t"], s.t) < 1.0
        assert abs(models[i]["scale"] - s.scale) < 0.02 * s.scale
        inl = masks[sl].astype(bool)
        assert (inl & s.inlier_mask).sum() > 0.9 * s.inlier_mask.sum()
        if variant >= 2:
            assert abs(models[i]["f1"] - s.f1) < 0.02 * s.f1 and abs(models[i]["f2"] - s.f2) < 0.02 * s.f2
        if variant == 1:
            assert abs(models[i]["shift1"] - s.shift1) < 0.05 and abs(models[i]["shift2"] - s.shift2) < 0.05


def test_pose_auc_parity_on_hard_scenes(ctx, port):
    """north_star: end-to-end pose AUC on synthetic scenes within 0.5 points of the reference path
    (hard variant: sigma 2 px, 10 % depth noise, 60 % outliers, 1000 iterations — SURVEY §8d)."""
    idx = range(300, 340)
    scs, variant, offs, x1, x2, d1, d2, cams = _batch("hard_calib", idx)
    models, stats, masks = ctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, _options(1000))
    rop = port.ransac_opt(max_iterations=1000, min_iterations=1000, max_epipolar_error=2.0, max_reproj_error=16.0)
    bop = port.bundle_opt(loss_type="TRUNCATED_CAUCHY")
    e_gpu, e_ref = [], []
    for i, s in enumerate(scs):
        m, st, mk = port.estimate(0, s.x1, s.x2, s.d1, s.d2, [800, 800, 640, 480], [800, 800, 640, 480], rop, bop)
        e_ref.append(max(rot_err_deg(np.array(m.q), s.R), trans_err_deg(np.array(m.t), s.t)))
        e_gpu.append(max(rot_err_deg(models[i]["q"], s.R), trans_err_deg(models[i]["t"], s.t)))
    for deg in (5, 10, 20):
        assert abs(maa(e_gpu, deg) - maa(e_ref, deg)) <= 0.5


def test_python_surface_matches_reference_api(ctx):
    """Same call and return surface as poselib (whl:_core.pyi:446-501) and the fork names of eval*.py."""
    sc = synth.scene_for("cfg1_calib_scale", 9, n=300)
    c1, c2 = sc.camera_dicts()
    ro = {"max_iterations": 300, "min_iterations": 300, "max_epipolar_error": 2.0, "max_reproj_error": 16.0,
          "lo_iterations": 25, "weight_sampson": 1.0}
    geom, info = api.estimate_monodepth_relative_pose(sc.x1.astype(np.float32), sc.x2, list(sc.d1), sc.d2, c1, c2, ro,
                                                      {"loss_type": "TRUNCATED_CAUCHY"})
    assert set(info) == {"refinements", "iterations", "num_inliers", "inlier_ratio", "model_score", "inliers"}
    assert len(info["inliers"]) == 300 and isinstance(info["inliers"][0], bool)
    assert geom.pose.R.shape == (3, 3) and geom.pose.t.shape == (3,) and geom.pose.q.shape == (4,)
    assert np.allclose(geom.pose.Rt[:, :3], geom.pose.R) and geom.shift1 == 0.0
    assert rot_err_deg(geom.pose.q, sc.R) < 0.5 and abs(geom.scale - 1.7) < 0.05
    x1, x2 = sc.centred()
    pair, info2 = api.estimate_monodepth_shared_focal_relative_pose(x1, x2, sc.d1, sc.d2, ro, {})
    assert abs(pair.camera1.focal() - 800) < 20 and pair.camera1.model_name() == "SIMPLE_PINHOLE"
    assert pair.camera1.params[1:] == [0.0, 0.0] and pair.geometry.pose.R.shape == (3, 3)
    pair3, _ = api.estimate_monodepth_varying_focal_relative_pose(x1, x2, sc.d1, sc.d2, ro, {})
    assert abs(pair3.camera2.focal() - 800) < 40
    # fork names (eval.py:153, eval_shared_f.py:177)
    K = {"model": "PINHOLE", "width": -1, "height": -1, "params": [800.0, 800.0, 640.0, 480.0]}
    pose, info3 = api.estimate_relative_pose_w_mono_depth(sc.x1, sc.x2, np.c_[sc.d1, sc.d2], K, K,
                                                          dict(ro, use_p3p=True, use_ours=False, solver_shift=False), {})
    assert rot_err_deg(pose.q, sc.R) < 0.5 and pose.R.shape == (3, 3)
    ip, _ = api.estimate_shared_focal_monodepth_relative_pose(x1, x2, np.c_[sc.d1, sc.d2], ro, {})
    assert ip.pose.R.shape == (3, 3) and ip.camera1.focal() > 0
    assert len(api.monodepth_pose_3pt(np.c_[sc.x1[:3] / 800, np.ones(3)], np.c_[sc.x2[:3] / 800, np.ones(3)],
                                      sc.d1[:3], sc.d2[:3])) <= 4


@pytest.mark.parametrize("cfg", ["cfg1_calib_scale", "cfg2_calib_shift", "cfg3_shared_focal", "cfg4_varying_focal", "hard_calib"])
def test_pruning_does_not_change_results(cfg, monkeypatch):
    """The FP32 bound kernel only drops minimal models that provably cannot trigger anything in
    score_models(): with RP_NO_PRUNE=1 every model is scored exactly, and all outputs must be bytewise
    identical."""
    c = synth.CONFIGS[cfg]
    scs, variant, offs, x1, x2, d1, d2, cams = _batch(cfg, range(400, 424), n=700)
    o = _options(2500, shift=c["shift"])
    monkeypatch.delenv("RP_NO_PRUNE", raising=False)
    monkeypatch.delenv("RP_NO_WAVES", raising=False)
    monkeypatch.delenv("RP_NO_TC", raising=False)
    monkeypatch.delenv("RP_TC_ONE_PASS", raising=False)
    pruned_ctx = nv.Context(0)
    a = pruned_ctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    _, cnt = pruned_ctx.last_timing()
    assert cnt["exact_models"] < 0.5 * cnt["hypotheses"]  # pruning is actually on
    # survivors scored all at once instead of in bar-raising waves: same results, more exact work
    monkeypatch.setenv("RP_NO_WAVES", "1")
    flat_ctx = nv.Context(0)
    w = flat_ctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    _, cntw = flat_ctx.last_timing()
    monkeypatch.delenv("RP_NO_WAVES")
    assert cnt["exact_models"] <= cntw["exact_models"] < 0.5 * cntw["hypotheses"]
    assert a[0].tobytes() == w[0].tobytes() and a[2].tobytes() == w[2].tobytes()
    for f in ("refinements", "iterations", "num_inliers", "inlier_ratio"):
        assert np.array_equal(a[1][f], w[1][f]), f
    assert np.allclose(a[1]["model_score"], w[1]["model_score"], rtol=1e-12, atol=0)
    flat_ctx.close()
    # without the tensor-core tier (FP32 bound kernel over every model, the round-1 path): same bytes
    monkeypatch.setenv("RP_NO_TC", "1")
    notc_ctx = nv.Context(0)
    t = notc_ctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    _, cntt = notc_ctx.last_timing()
    monkeypatch.delenv("RP_NO_TC")
    assert cntt["tc_evaluated"] == 0 and cnt["tc_evaluated"] > 0 and 0 < cnt["tc_selected"] < 0.5 * cnt["hypotheses"]
    assert a[0].tobytes() == t[0].tobytes() and a[2].tobytes() == t[2].tobytes()
    for f in ("refinements", "iterations", "num_inliers", "inlier_ratio"):
        assert np.array_equal(a[1][f], t[1][f]), f
    assert np.allclose(a[1]["model_score"], t[1]["model_score"], rtol=1e-12, atol=0)
    notc_ctx.close()
    # the tensor-core tier in ONE pass over the correspondences instead of two (no mid-way abandonment): same bytes
    monkeypatch.setenv("RP_TC_ONE_PASS", "1")
    one_ctx = nv.Context(0)
    u = one_ctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    _, cntu = one_ctx.last_timing()
    monkeypatch.delenv("RP_TC_ONE_PASS")
    assert cntu["tc_evaluated"] >= cnt["tc_evaluated"] and cntu["tc_selected"] == cnt["tc_selected"]
    assert a[0].tobytes() == u[0].tobytes() and a[2].tobytes() == u[2].tobytes() and a[1].tobytes() == u[1].tobytes()
    one_ctx.close()
    monkeypatch.setenv("RP_NO_PRUNE", "1")
    full_ctx = nv.Context(0)
    b = full_ctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    _, cnt2 = full_ctx.last_timing()
    assert cnt2["exact_models"] == cnt2["hypotheses"]
    # models, masks, counters: bytewise.  model_score: the order in which the exact kernel sums the inlier
    # residuals of a model depends on which other models share its work item, so the two modes may differ in
    # the last bits (tie rule)
    assert a[0].tobytes() == b[0].tobytes() and a[2].tobytes() == b[2].tobytes()
    for f in ("refinements", "iterations", "num_inliers", "inlier_ratio"):
        assert np.array_equal(a[1][f], b[1][f]), f
    assert np.allclose(a[1]["model_score"], b[1]["model_score"], rtol=1e-12, atol=0)
    pruned_ctx.close()
    full_ctx.close()


@pytest.mark.parametrize("cfg", ["cfg1_calib_scale", "cfg2_calib_shift", "cfg3_shared_focal"])
def test_two_phase_solve_kernels_are_bit_identical(cfg, monkeypatch):
    """solve2_kernel (roots per sample, then one thread per queued root) against the thread-per-iteration
    solve_kernel (RP_SOLVE_GENERIC=1): same arithmetic, so every output byte must agree."""
    c = synth.CONFIGS[cfg]
    scs, variant, offs, x1, x2, d1, d2, cams = _batch(cfg, range(500, 516), n=500)
    o = _options(3000, shift=c["shift"])
    monkeypatch.delenv("RP_SOLVE_GENERIC", raising=False)
    two = nv.Context(0)
    a = two.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    monkeypatch.setenv("RP_SOLVE_GENERIC", "1")
    gen = nv.Context(0)
    b = gen.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    for u, v in zip(a, b):
        assert u.tobytes() == v.tobytes()
    two.close()
    gen.close()


def test_extra_goldens_end_to_end(ctx, extra_golden):
    """The option combinations the randomised runs showed to matter (tests/golden/make_golden_extra.py): outputs of
    the reference binary for weight_sampson != 1, early termination after a late LO, one-iteration runs, all losses,
    PROSAC on/off.  Stats, mask and model_score must agree; the model where the refinement is determined."""
    g = extra_golden
    n = ties = 0
    for key, variant, o in extra_cases(g):
        opt = nv.default_options()
        opt.max_iterations, opt.min_iterations = o["iters"], o["min_iters"]
        opt.max_epipolar_error, opt.max_reproj_error, opt.seed = o["t_epi"], o["t_rep"], o["seed"]
        opt.estimate_shift = int(variant == 1)
        opt.weight_sampson = o["weight_sampson"]
        opt.progressive_sampling, opt.max_prosac_iterations = int(o["prosac"]), o["max_prosac"]
        opt.loss_type, opt.loss_scale = nv.LOSS[o["loss"]], 0.5 * o["t_epi"]
        opt.bundle_max_iterations = o["bundle_iters"]
        npts = len(g[key + "_d1"])
        f1, f2 = g[key + "_f"]
        cams = np.array([[f1, f1, 640, 480, f2, f2, 640, 480]]) if variant < 2 else None
        models, stats, masks = ctx.estimate_batch_host(variant, [0, npts], g[key + "_x1"], g[key + "_x2"],
                                                       g[key + "_d1"], g[key + "_d2"], cams, opt)
        ref = tuple(int(v) for v in g[key + "_stats"])
        assert (stats[0]["iterations"], stats[0]["num_inliers"]) == ref[1:], (key, o)
        if stats[0]["refinements"] != ref[0]:
            # tie rule: the same minimal model met twice; nothing but the counter may change
            ties += 1
            assert models_close(models[0], g[key + "_model"], rtol=1e-6, atol=1e-8), (key, o)
        assert np.array_equal(masks, g[key + "_mask"]), key
        ref_score = g[key + "_fstats"][1]
        assert stats[0]["model_score"] == ref_score or abs(stats[0]["model_score"] - ref_score) <= 1e-9 * abs(ref_score) + 1e-24, key
        if ref[2] >= 10:
            assert models_close(models[0], g[key + "_model"], rtol=1e-6, atol=1e-8), key
        n += 1
    assert n == 160 and ties <= 5


@pytest.mark.parametrize("variant", [2, 3])
def test_no_minimal_model_found(ctx, port, variant):
    """One iteration whose sample yields no solution: the reference refines the identity model (NaN cost under the
    truncated loss, so the LM moves nothing), reports num_inliers of that model and leaves model_score at DBL_MAX."""
    sc = synth.make_scene(11, 40, outlier_ratio=0.5, f1=700.0, f2=700.0 if variant == 2 else 900.0)
    x1, x2 = sc.x1 - [640.0, 480.0], sc.x2 - [640.0, 480.0]
    seen_empty = False
    for seed in range(8):
        o = _options(1, seed=seed)
        o.bundle_max_iterations = 0
        models, stats, masks = ctx.estimate_batch_host(variant, [0, 40], x1, x2, sc.d1, sc.d2, None, o)
        ro = port.ransac_opt(max_iterations=1, min_iterations=1, max_epipolar_error=2.0, max_reproj_error=16.0, seed=seed)
        m, st, mask = port.estimate(variant, x1, x2, sc.d1, sc.d2, None, None, ro,
                                    port.bundle_opt(max_iterations=0, loss_type="TRUNCATED_CAUCHY", loss_scale=1.0))
        assert (stats[0]["refinements"], stats[0]["iterations"], stats[0]["num_inliers"]) == \
            (st.refinements, st.iterations, st.num_inliers), seed
        assert np.array_equal(masks.astype(bool), mask), seed
        if st.refinements == 1:   # nothing but the final refinement ran: no minimal model
            seen_empty = True
            assert stats[0]["model_score"] == DBL_MAX and st.model_score == DBL_MAX
            assert np.array_equal(models[0]["q"], [1, 0, 0, 0]) and models[0]["scale"] == 1.0
            assert abs(models[0]["f1"] - m.f1) <= 1e-12 * m.f1
        else:
            assert abs(stats[0]["model_score"] - st.model_score) <= 1e-11 * st.model_score, seed
    assert seen_empty


@pytest.mark.parametrize("cfg", ["cfg1_calib_scale", "cfg2_calib_shift", "cfg3_shared_focal", "cfg4_varying_focal"])
def test_lo_kernels_agree(cfg, monkeypatch):
    """The LO refinements run one warp per LM problem (lm_warp_kernel), the final LO / final refinement one block per
    problem (lm_kernel); RP_LM_WARP=0 runs everything on the block kernel.  Same LM, different summation order of the
    normal equations: inlier masks and counters equal, models to 1e-9 (a pair whose refinement count differs by the
    tie rule of DESIGN.md §3 is tolerated, at most one in the batch)."""
    c = synth.CONFIGS[cfg]
    scs, variant, offs, x1, x2, d1, d2, cams = _batch(cfg, range(600, 632), n=600)
    o = _options(1500, shift=c["shift"])
    monkeypatch.delenv("RP_LM_WARP", raising=False)
    wctx = nv.Context(0)
    a = wctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    wctx.close()
    monkeypatch.setenv("RP_LM_WARP", "0")
    bctx = nv.Context(0)
    b = bctx.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    bctx.close()
    monkeypatch.delenv("RP_LM_WARP")
    assert np.array_equal(a[1]["num_inliers"], b[1]["num_inliers"]) and np.array_equal(a[1]["iterations"], b[1]["iterations"])
    assert (a[1]["refinements"] != b[1]["refinements"]).sum() <= 1
    assert a[2].tobytes() == b[2].tobytes()
    assert np.allclose(a[1]["model_score"], b[1]["model_score"], rtol=1e-9, atol=0)
    for f in ("q", "t", "scale", "shift1", "shift2", "f1", "f2"):
        if f in a[0].dtype.names:
            assert np.allclose(a[0][f], b[0][f], rtol=1e-9, atol=1e-9), f


@pytest.mark.parametrize("cfg", ["cfg2_calib_shift", "cfg3_shared_focal", "cfg4_varying_focal"])
def test_pruning_does_not_change_results_at_baseline_size(cfg, monkeypatch):
    """The same guard as above at BASELINE.json's size (2 000 matches x 10 000 iterations: ten 1 024-iteration segments
    per pair, mid stage and both tensor-core passes active) on the bench's own generator: the whole cascade against
    RP_NO_PRUNE=1 (every minimal model scored exactly), bytewise."""
    c = synth.CONFIGS[cfg]
    b = synth.make_batch(cfg, 24, seed=4242)
    variant = {"calib": nv.CALIB_SHIFT if c["shift"] else nv.CALIB, "shared": nv.SHARED, "varying": nv.VARYING}[c["variant"]]
    o = _options(c["iters"], shift=c["shift"])
    for k in ("RP_NO_PRUNE", "RP_NO_WAVES", "RP_NO_TC", "RP_TC_ONE_PASS", "RP_MID_END"):
        monkeypatch.delenv(k, raising=False)
    ctx = nv.Context(0)
    a = ctx.estimate_batch_host(variant, b["offsets"], b["x1"], b["x2"], b["d1"], b["d2"], b["cams"], o)
    _, cnt = ctx.last_timing()
    ctx.close()
    assert cnt["exact_models"] < 0.05 * cnt["hypotheses"] and cnt["tc_evaluated"] > 0
    monkeypatch.setenv("RP_NO_PRUNE", "1")
    full = nv.Context(0)
    f = full.estimate_batch_host(variant, b["offsets"], b["x1"], b["x2"], b["d1"], b["d2"], b["cams"], o)
    _, cntf = full.last_timing()
    full.close()
    monkeypatch.delenv("RP_NO_PRUNE")
    assert cntf["exact_models"] == cntf["hypotheses"]
    assert a[0].tobytes() == f[0].tobytes() and a[2].tobytes() == f[2].tobytes()
    for k in ("refinements", "iterations", "num_inliers", "inlier_ratio"):
        assert np.array_equal(a[1][k], f[1][k]), k
    assert np.allclose(a[1]["model_score"], f[1]["model_score"], rtol=1e-12, atol=0)
    # the plain 128-model exact head instead of the short head + mid stage: same bytes
    monkeypatch.setenv("RP_MID_END", "0")
    head = nv.Context(0)
    h = head.estimate_batch_host(variant, b["offsets"], b["x1"], b["x2"], b["d1"], b["d2"], b["cams"], o)
    _, cnth = head.last_timing()
    head.close()
    monkeypatch.delenv("RP_MID_END")
    assert cnth["exact_models"] > cnt["exact_models"]
    assert a[0].tobytes() == h[0].tobytes() and a[2].tobytes() == h[2].tobytes()
    for k in ("refinements", "iterations", "num_inliers", "inlier_ratio"):
        assert np.array_equal(a[1][k], h[1][k]), k
