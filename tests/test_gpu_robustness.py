"""One hard pair must not poison its batch (VERDICT r1 "What's weak" #4).

* A pair with more LO trigger events than the first pass keeps is finished on its own (RP_EV_CAP shrinks the list so that
  ordinary pairs exercise that path): same bytes as with the default capacity.
* Early termination with the reference's default options (min 1000 / max 100000 iterations): a near-zero-inlier pair in a
  large batch keeps iterating on its own; every other pair is bytewise what it is without the hard pair in the batch,
  the hard pair equals the oracle, memory stays bounded, and rp_pair_status says which pairs were re-run."""
import numpy as np
import pytest

from mdrp_b200 import _native as nv, synth

pytestmark = pytest.mark.gpu


def _pack(scs):
    offs = np.r_[0, np.cumsum([len(s.d1) for s in scs])]
    x1, x2 = np.concatenate([s.x1 for s in scs]), np.concatenate([s.x2 for s in scs])
    d1, d2 = np.concatenate([s.d1 for s in scs]), np.concatenate([s.d2 for s in scs])
    cams = np.array([[s.f1, s.f1, 640, 480, s.f2, s.f2, 640, 480] for s in scs], dtype=np.float64)
    return offs, x1, x2, d1, d2, cams


@pytest.mark.parametrize("variant,cfg", [(0, "cfg1_calib_scale"), (1, "cfg2_calib_shift")])
def test_event_list_overflow_is_rerun_per_pair(variant, cfg, monkeypatch):
    scs = [synth.scene_for(cfg, 700 + i, n=[400, 150, 3, 900][i % 4]) for i in range(48)]
    offs, x1, x2, d1, d2, cams = _pack(scs)
    o = nv.default_options()
    o.max_iterations = o.min_iterations = 2000
    o.max_epipolar_error, o.max_reproj_error, o.estimate_shift = 2.0, 16.0, int(variant == 1)
    o.loss_type = nv.LOSS["TRUNCATED_CAUCHY"]
    monkeypatch.delenv("RP_EV_CAP", raising=False)
    base = nv.Context(0)
    a = base.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    st_a = base.pair_status(len(scs))
    assert set(st_a.tolist()) <= {nv.PAIR_OK, nv.PAIR_DEGENERATE}
    monkeypatch.setenv("RP_EV_CAP", "4")
    small = nv.Context(0)
    b = small.estimate_batch_host(variant, offs, x1, x2, d1, d2, cams, o)
    st_b = small.pair_status(len(scs))
    # more than 4 triggers is the normal case (refinements counts them, plus the final LO)
    rerun = (st_b & nv.PAIR_EVENTS_RERUN) != 0
    assert rerun.sum() >= len(scs) // 2 and not (st_b < 0).any()
    assert np.array_equal(st_b == nv.PAIR_DEGENERATE, st_a == nv.PAIR_DEGENERATE)
    assert a[0].tobytes() == b[0].tobytes() and a[2].tobytes() == b[2].tobytes()
    for f in ("refinements", "iterations", "num_inliers", "inlier_ratio"):
        assert np.array_equal(a[1][f], b[1][f]), f
    assert np.allclose(a[1]["model_score"], b[1]["model_score"], rtol=1e-12, atol=0)
    base.close()
    small.close()


def test_default_options_with_one_hard_pair(ctx, port):
    """4096 pairs under PoseLib's default RansacOptions (early termination, max 100000 iterations); pair 1234 has 2 %
    inliers and iterates 100x longer than the rest."""
    P, n, hard = 4096, 300, 1234
    scs = [synth.scene_for("cfg1_calib_scale", 9000 + i, n=n) for i in range(P)]
    scs[hard] = synth.make_scene(424242, n, outlier_ratio=0.98)
    offs, x1, x2, d1, d2, cams = _pack(scs)
    o = nv.default_options()          # 1000 / 100000, dyn_num_trials_mult 3, success_prob 0.9999
    o.max_epipolar_error, o.max_reproj_error = 2.0, 16.0
    o.loss_type = nv.LOSS["TRUNCATED_CAUCHY"]
    models, stats, masks = ctx.estimate_batch_host(nv.CALIB, offs, x1, x2, d1, d2, cams, o)
    status = ctx.pair_status(P)
    assert not (status < 0).any()
    assert status[hard] & nv.PAIR_CONTINUED
    assert (status & nv.PAIR_CONTINUED).sum() <= P // 100          # only the hard pair(s) were run again
    assert stats[hard]["iterations"] > 10 * np.median(stats["iterations"])
    # the same batch without the hard pair: every other pair bytewise identical
    keep = np.r_[0:hard, hard + 1:P]
    scs2 = [scs[i] for i in keep]
    offs2, a1, a2, b1, b2, cams2 = _pack(scs2)
    m2, s2, k2 = ctx.estimate_batch_host(nv.CALIB, offs2, a1, a2, b1, b2, cams2, o)
    assert models[keep].tobytes() == m2.tobytes() and stats[keep].tobytes() == s2.tobytes()
    mk = np.concatenate([masks[offs[i]:offs[i + 1]] for i in keep])
    assert np.array_equal(mk, k2)
    # the hard pair (and a few ordinary ones) against the oracle
    ro = port.ransac_opt(max_epipolar_error=2.0, max_reproj_error=16.0)
    bo = port.bundle_opt(loss_type="TRUNCATED_CAUCHY")
    for i in (hard, 0, 77, P - 1):
        s = scs[i]
        m, st, mask = port.estimate(0, s.x1, s.x2, s.d1, s.d2, [800, 800, 640, 480], [800, 800, 640, 480], ro, bo)
        assert (stats[i]["refinements"], stats[i]["iterations"], stats[i]["num_inliers"]) == \
            (st.refinements, st.iterations, st.num_inliers), i
        assert np.array_equal(masks[offs[i]:offs[i + 1]].astype(bool), mask), i
