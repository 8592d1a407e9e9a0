"""Tensor-core count tier (mdrp_b200/csrc/rp_tc.cuh) through its stage entry point rp_tc_count_batch.

The tier is a filter: for every model it must report AT MOST as many certain outliers as the reference's Sampson test
(compute_sampson_msac_score so@0x4f61d0 / so@0x4f65d0, restated in numpy here and pinned against the wheel by
tests/test_oracle_vs_ref.py) has outliers — never more — and it should be sharp enough to be useful."""
import numpy as np
import pytest

from mdrp_b200 import _native as nv, synth
from util import struct_models

pytestmark = pytest.mark.gpu


def _E(q, t):
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    return tx @ R


def _outliers(F, x1, x2, sq_thr):
    """#{k : not (r2_k < thr^2)} in FP64 (NaN counts as outlier, as in the reference)."""
    h1 = np.c_[x1, np.ones(len(x1))]
    h2 = np.c_[x2, np.ones(len(x2))]
    Fx1 = h1 @ F.T
    Ftx2 = h2 @ F
    C = np.sum(h2 * Fx1, axis=1)
    with np.errstate(all="ignore"):
        r2 = C * C / (Fx1[:, 0] ** 2 + Fx1[:, 1] ** 2 + Ftx2[:, 0] ** 2 + Ftx2[:, 1] ** 2)
    return int(np.sum(~(r2 < sq_thr)))


def _quat(axis, ang):
    axis = axis / np.linalg.norm(axis)
    return np.r_[np.cos(ang / 2), np.sin(ang / 2) * axis]


def _qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def _models_around(sc, rng, n):
    """The true pose perturbed by a log-uniform amount: from near-perfect to useless models."""
    from scipy.spatial.transform import Rotation
    q0 = Rotation.from_matrix(sc.R).as_quat()  # x y z w
    q0 = np.r_[q0[3], q0[:3]]
    rows = np.zeros((n, 12))
    for i in range(n):
        mag = 10.0 ** rng.uniform(-4.5, 0.0)
        q = _qmul(_quat(rng.normal(size=3), mag * rng.uniform()), q0)
        rows[i] = np.r_[q, sc.t + mag * rng.normal(size=3), 1.7, 0, 0, 1, 1]
    return rows


@pytest.mark.parametrize("n", [2000, 1999, 63, 64, 65, 10000, 3])
def test_counts_are_rigorous_and_sharp_pose(ctx, n):
    sc = synth.make_scene(7, n, outlier_ratio=0.3)
    rng = np.random.default_rng(n)
    rows = _models_around(sc, rng, 700)
    rows[5, 4] = np.nan            # NaN model: every point is an outlier in the reference
    rows[6, :4] = [0, 0, 0, 0]     # zero rotation quaternion -> E = [t]x * 0-ish
    x1, x2 = (sc.x1 - synth.PP) / 800.0, (sc.x2 - synth.PP) / 800.0
    thr2 = (2.0 / 800.0) ** 2
    out = ctx.tc_count(nv.CALIB, struct_models(rows), x1, x2, thr2)
    tot_out = tot_true = 0
    for i, r in enumerate(rows):
        E = _E(r[:4], r[4:7])
        true = _outliers(E, x1, x2, thr2)
        assert 0 <= out[i] <= true, (i, out[i], true)
        if not np.isfinite(E).all():
            assert out[i] == n
        elif i != 6:
            tot_out += out[i]
            tot_true += true
    if n >= 63:
        assert tot_out >= 0.97 * tot_true   # sharp: within 3 % of the exact outlier count


def test_counts_are_rigorous_focal(ctx):
    """F = diag(1,1,f2) E diag(1,1,f1) on normalize_points-scaled coordinates (focal variants), and a threshold
    sweep down to where the filter has to give up."""
    sc = synth.make_scene(9, 1500, outlier_ratio=0.5, f1=700.0, f2=900.0)
    rng = np.random.default_rng(3)
    x1, x2 = sc.x1 - synth.PP, sc.x2 - synth.PP
    s = (np.linalg.norm(x1, axis=1).sum() + np.linalg.norm(x2, axis=1).sum()) / (np.sqrt(2) * len(x1))
    x1, x2 = x1 / s, x2 / s
    rows = _models_around(sc, rng, 500)
    rows[:, 10] = 700.0 / s * (1 + 0.2 * rng.normal(size=500))
    rows[:, 11] = 900.0 / s * (1 + 0.2 * rng.normal(size=500))
    for thr_px in (2.0, 0.5, 1e-3):
        thr2 = (thr_px / s) ** 2
        out = ctx.tc_count(nv.VARYING, struct_models(rows), x1, x2, thr2)
        for i, r in enumerate(rows):
            F = np.diag([1, 1, r[11]]) @ _E(r[:4], r[4:7]) @ np.diag([1, 1, r[10]])
            assert 0 <= out[i] <= _outliers(F, x1, x2, thr2), (thr_px, i)


def test_adversarial_scales(ctx):
    """Models and points at the edges of the ranges the error model covers: the count may drop to 0 (filter disabled)
    but must never exceed the exact outlier count."""
    rng = np.random.default_rng(11)
    sc = synth.make_scene(3, 800)
    x1, x2 = (sc.x1 - synth.PP) / 800.0, (sc.x2 - synth.PP) / 800.0
    base = _models_around(sc, rng, 64)
    for tscale in (1e-6, 1e-3, 1.0, 1e2, 1e5):
        for pscale in (1e-3, 1.0, 30.0, 5e3):
            rows = base.copy()
            rows[:, 4:7] *= tscale
            a, b = x1 * pscale, x2 * pscale
            thr2 = (2.0 / 800.0 * pscale) ** 2
            out = ctx.tc_count(nv.CALIB, struct_models(rows), a, b, thr2)
            for i, r in enumerate(rows):
                assert 0 <= out[i] <= _outliers(_E(r[:4], r[4:7]), a, b, thr2), (tscale, pscale, i)
