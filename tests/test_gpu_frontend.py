"""GPU tier: the device-resident front end — depth lookup + inf mask + compaction on
the device and the torch-tensor entry points — against the reference's host-side recipe
(/root/reference/make_pair.py:97-113)."""
import numpy as np
import pytest

from mdrp_b200 import _native as nv, api, synth
from util import models_close

pytestmark = pytest.mark.gpu


def _scene_with_maps(n=600, seed=3):
    """Keypoints + dense depth maps whose values at the keypoints equal the scene's depths."""
    import torch
    rng = np.random.default_rng(seed)
    sc = synth.scene_for("cfg1_calib_scale", 31, n=n)
    H, W = 960, 1280
    dm1 = rng.uniform(1, 9, size=(H, W)).astype(np.float32)
    dm2 = rng.uniform(1, 9, size=(H, W)).astype(np.float32)
    kp1 = np.clip(sc.x1, 0, [W - 1.001, H - 1.001]).astype(np.float32)
    kp2 = np.clip(sc.x2, 0, [W - 1.001, H - 1.001]).astype(np.float32)
    dm1[kp1[:, 1].astype(int), kp1[:, 0].astype(int)] = sc.d1.astype(np.float32)
    dm2[kp2[:, 1].astype(int), kp2[:, 0].astype(int)] = sc.d2.astype(np.float32)
    # MoGe marks invalid pixels with inf: both-inf rows must be dropped, single-inf rows kept (make_pair.py:106)
    both = rng.choice(n, 25, replace=False)
    dm1[kp1[both, 1].astype(int), kp1[both, 0].astype(int)] = np.inf
    dm2[kp2[both, 1].astype(int), kp2[both, 0].astype(int)] = np.inf
    t = lambda a: torch.from_numpy(a).cuda()
    return sc, dm1, dm2, kp1, kp2, t(dm1), t(dm2), t(kp1), t(kp2)


def test_gather_depths_matches_reference_recipe(ctx):
    from mdrp_b200 import torch_frontend as tf
    sc, dm1, dm2, kp1, kp2, g1, g2, k1, k2 = _scene_with_maps()
    d1 = dm1[kp1[:, 1].astype(int), kp1[:, 0].astype(int)]
    d2 = dm2[kp2[:, 1].astype(int), kp2[:, 0].astype(int)]
    keep = ~np.logical_and(np.isinf(d1), np.isinf(d2))
    x1, x2, e1, e2 = tf.gather_depths(g1, g2, k1, k2)
    assert x1.shape[0] == keep.sum() and x1.dtype.is_floating_point and x1.is_cuda
    assert np.array_equal(x1.cpu().numpy(), kp1[keep].astype(np.float64))
    assert np.array_equal(x2.cpu().numpy(), kp2[keep].astype(np.float64))
    assert np.array_equal(e1.cpu().numpy(), d1[keep].astype(np.float64))
    assert np.array_equal(e2.cpu().numpy(), d2[keep].astype(np.float64))


def test_torch_path_equals_host_path(ctx):
    """Same estimator, inputs already on the device: results must be bytewise those of the host-buffer call."""
    import torch
    from mdrp_b200 import torch_frontend as tf
    sc, dm1, dm2, kp1, kp2, g1, g2, k1, k2 = _scene_with_maps()
    x1, x2, d1, d2 = tf.gather_depths(g1, g2, k1, k2)
    finite = torch.isfinite(d1) & torch.isfinite(d2)
    x1, x2, d1, d2 = x1[finite], x2[finite], d1[finite], d2[finite]
    c1, c2 = sc.camera_dicts()
    ro = {"max_iterations": 500, "min_iterations": 500, "max_epipolar_error": 2.0, "max_reproj_error": 16.0}
    bo = {"loss_type": "TRUNCATED_CAUCHY"}
    g_dev, info_dev = tf.estimate_monodepth_relative_pose(x1, x2, d1, d2, c1, c2, ro, bo)
    g_host, info_host = api.estimate_monodepth_relative_pose(x1.cpu().numpy(), x2.cpu().numpy(), d1.cpu().numpy(),
                                                             d2.cpu().numpy(), c1, c2, ro, bo)
    assert info_dev == info_host
    assert np.array_equal(g_dev.pose.q, g_host.pose.q) and np.array_equal(g_dev.pose.t, g_host.pose.t)
    assert g_dev.scale == g_host.scale and info_dev["num_inliers"] > 300


def test_inf_depths_do_not_poison_the_batch(ctx, port):
    """Rows with one infinite depth stay in (reference behaviour); they must not break the estimate of
    that pair nor of its neighbours, and the result must equal the oracle's on the same rows."""
    scs = [synth.scene_for("cfg1_calib_scale", 70 + i, n=400) for i in range(3)]
    d1 = [s.d1.copy() for s in scs]
    d1[1][::17] = np.inf
    offs = np.r_[0, np.cumsum([400] * 3)]
    cams = np.tile(np.array([800., 800, 640, 480, 800, 800, 640, 480]), (3, 1))
    o = nv.default_options()
    o.max_iterations = o.min_iterations = 400
    o.max_epipolar_error, o.max_reproj_error = 2.0, 16.0
    o.loss_type = nv.LOSS["TRUNCATED_CAUCHY"]
    models, stats, masks = ctx.estimate_batch_host(0, offs, np.concatenate([s.x1 for s in scs]),
                                                   np.concatenate([s.x2 for s in scs]), np.concatenate(d1),
                                                   np.concatenate([s.d2 for s in scs]), cams, o)
    rop = port.ransac_opt(max_iterations=400, min_iterations=400, max_epipolar_error=2.0, max_reproj_error=16.0)
    bop = port.bundle_opt(loss_type="TRUNCATED_CAUCHY")
    for i, s in enumerate(scs):
        m, st, mk = port.estimate(0, s.x1, s.x2, d1[i], s.d2, [800, 800, 640, 480], [800, 800, 640, 480], rop, bop)
        assert np.isfinite(models[i]["q"]).all() and stats[i]["num_inliers"] > 200
        assert (st.refinements, st.num_inliers) == (stats[i]["refinements"], stats[i]["num_inliers"]), i
        assert np.array_equal(mk, masks[offs[i]:offs[i + 1]].astype(bool)), i
        assert models_close(models[i], m, rtol=1e-6, atol=1e-8), i


def test_batched_gather_equals_per_pair(ctx):
    """rp_gather_depths_batch_dev over a ragged batch of frame pairs == rp_gather_depths_dev pair by pair, and the packed
    result feeds the batched estimator."""
    import torch
    from mdrp_b200 import torch_frontend as tf
    rng = np.random.default_rng(5)
    F, H, W = 4, 120, 160
    maps = rng.uniform(1, 9, size=(F, H, W)).astype(np.float32)
    maps[rng.uniform(size=maps.shape) < 0.3] = np.inf
    sizes = [700, 0, 33, 1025, 256]
    f1, f2 = [0, 1, 2, 3, 0], [1, 2, 3, 0, 2]
    offs = np.r_[0, np.cumsum(sizes)]
    kp1 = rng.uniform(-3, [W + 3, H + 3], size=(offs[-1], 2)).astype(np.float32)   # a few outside: clamped to the border
    kp2 = rng.uniform(-3, [W + 3, H + 3], size=(offs[-1], 2)).astype(np.float32)
    g = torch.from_numpy(maps).cuda()
    k1, k2 = torch.from_numpy(kp1).cuda(), torch.from_numpy(kp2).cuda()
    out, x1, x2, d1, d2 = tf.gather_depths_batch(g, f1, f2, offs, k1, k2)
    assert out[0] == 0 and len(out) == len(sizes) + 1 and x1.shape[0] == out[-1]
    for p in range(len(sizes)):
        sl = slice(offs[p], offs[p + 1])
        if sizes[p] == 0:
            assert out[p + 1] == out[p]
            continue
        a1, a2, e1, e2 = tf.gather_depths(g[f1[p]], g[f2[p]], k1[sl], k2[sl])
        o = slice(out[p], out[p + 1])
        assert a1.shape[0] == out[p + 1] - out[p] and 0 < a1.shape[0] < sizes[p]
        assert torch.equal(x1[o], a1) and torch.equal(x2[o], a2) and torch.equal(d1[o], e1) and torch.equal(d2[o], e2)
