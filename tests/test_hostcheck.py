"""CPU tier: the PRODUCT's device math (mdrp_b200/csrc/*.cuh, RP_HD functions) compiled for the host
by tests/hostcheck and compared with the oracle.  This is how kernels are debugged on the GPU-less
build box; the product library itself has no host path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mdrp_b200 import synth
from util import VARIANT_ID, model_vec

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
SO = os.path.join(HERE, "hostcheck", "libhostcheck.so")
DP = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def hc():
    csrc = os.path.join(os.path.dirname(HERE), "mdrp_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc))
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(newest, os.path.getmtime(SRC)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-fvisibility=hidden", "-o", SO, SRC])
    L = C.CDLL(SO)
    L.hc_solve.restype = C.c_int
    L.hc_score.restype = C.c_double
    return L


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(DP)


def _scene(variant, idx=0, n=500):
    cfg = {"calib": "cfg1_calib_scale", "calib_shift": "cfg2_calib_shift", "shared": "cfg3_shared_focal",
           "varying": "cfg4_varying_focal"}[variant]
    sc = synth.scene_for(cfg, idx, n=n)
    if variant in ("calib", "calib_shift"):
        return sc, (sc.x1 - synth.PP) / sc.f1, (sc.x2 - synth.PP) / sc.f2, sc.f1
    x1, x2 = sc.centred()
    ns = (np.linalg.norm(x1, axis=1) + np.linalg.norm(x2, axis=1)).sum() / (np.sqrt(2) * len(x1))
    return sc, x1 / ns, x2 / ns, ns


@pytest.mark.parametrize("variant", ["calib", "calib_shift", "shared", "varying"])
def test_device_solvers_equal_oracle(hc, port, variant):
    fn = {"calib": port.solve_calib_scale, "calib_shift": port.solve_calib_shift,
          "shared": port.solve_shared_focal, "varying": port.solve_varying_focal}[variant]
    sc, x1, x2, _ = _scene(variant)
    rng = np.random.default_rng(3)
    for _ in range(800):
        idx = rng.choice(len(x1), 3, replace=False)
        x1h, x2h = _d(np.c_[x1[idx], np.ones(3)]), _d(np.c_[x2[idx], np.ones(3)])
        d1, d2 = _d(sc.d1[idx]), _d(sc.d2[idx])
        out = (port.Model * 4)()
        n = hc.hc_solve(VARIANT_ID[variant], _p(x1h), _p(x2h), _p(d1), _p(d2), out)
        ref = fn(x1h, x2h, d1, d2)
        assert n == len(ref)
        for k in range(n):
            # same operation order, no FMA contraction on either side: bit identical on the host
            assert np.array_equal(model_vec(out[k].as_tuple()), model_vec(ref[k]), equal_nan=True)


@pytest.mark.parametrize("variant", ["calib", "calib_shift", "shared", "varying"])
def test_two_tier_scorer_exact_counts(hc, port, variant):
    """FP32 filter + exact FP64 tier reproduce the oracle's count / mask; the filter rejects most."""
    fn = {"calib": port.solve_calib_scale, "calib_shift": port.solve_calib_shift,
          "shared": port.solve_shared_focal, "varying": port.solve_varying_focal}[variant]
    sc, x1, x2, ns = _scene(variant, idx=1, n=800)
    thr2 = (2.0 / ns) ** 2
    rng = np.random.default_rng(4)
    X1, X2 = _d(x1), _d(x2)
    tested = t0_total = 0
    for _ in range(250):
        idx = rng.choice(len(x1), 3, replace=False)
        for m in fn(np.c_[x1[idx], np.ones(3)], np.c_[x2[idx], np.ones(3)], sc.d1[idx], sc.d2[idx]):
            mm = port.make_model(*m)
            if variant in ("calib", "calib_shift"):
                s, c = port.msac_score_pose(m[0], m[1], x1, x2, thr2)
                mk = port.get_inliers_pose(m[0], m[1], x1, x2, thr2)
            else:
                F = port.fundamental_from_model(mm)
                s, c = port.msac_score_F(F, x1, x2, thr2)
                mk = port.get_inliers_F(F, x1, x2, thr2)
            cnt, t0 = C.c_long(), C.c_long()
            mask = np.zeros(len(x1), dtype=np.uint8)
            s2 = hc.hc_score(VARIANT_ID[variant], C.byref(mm), _p(X1), _p(X2), C.c_long(len(x1)), C.c_double(thr2),
                             C.byref(cnt), C.byref(t0), mask.ctypes.data_as(C.c_void_p))
            assert cnt.value == c
            assert np.array_equal(mask.astype(bool), mk)
            if s == s:
                assert abs(s2 - s) <= 1e-12 * abs(s)
            tested += 1
            t0_total += t0.value
    assert tested > 50
    assert t0_total / (tested * len(x1)) > 0.5  # the FP32 tier does most of the work
    st = (C.c_long * 3)()
    hc.hc_screen_stats(st)
    assert st[0] == 0  # the FP32 cheirality screen never contradicts the exact FP64 test
    if variant in ("calib", "calib_shift"):
        assert st[1] > 20 * max(1, st[2])  # and it decides almost every point


@pytest.mark.parametrize("variant", ["calib", "calib_shift", "shared", "varying"])
@pytest.mark.parametrize("loss", ["TRUNCATED", "TRUNCATED_CAUCHY", "CAUCHY", "HUBER", "TRIVIAL"])
def test_device_lm_normal_equations(hc, port, variant, loss):
    from scipy.spatial.transform import Rotation as Rot
    sc, x1, x2, ns = _scene(variant, idx=2, n=400)
    rng = np.random.default_rng(6)
    qq = Rot.from_matrix(sc.R).as_quat()
    q = np.array([qq[3], qq[0], qq[1], qq[2]]) + 0.01 * rng.normal(size=4)
    q /= np.linalg.norm(q)
    t = sc.t + 0.02 * rng.normal(size=3)
    focal = variant in ("shared", "varying")
    f1 = sc.f1 / ns * 1.02 if focal else 1.0
    f2 = (f1 if variant == "shared" else sc.f2 / ns * 0.97) if focal else 1.0
    m = port.make_model(q, t, 1.72, 0.05 if variant == "calib_shift" else 0, -0.04 if variant == "calib_shift" else 0, f1, f2)
    v = VARIANT_ID[variant]
    thr, sr = 2.0 / ns, (2.0 / 16.0) ** 2
    JtJ, Jtr = port.accumulate(v, x1, x2, sc.d1, sc.d2, m, sr, 1.0, loss, thr)
    cost = port.cost(v, x1, x2, sc.d1, sc.d2, m, sr, 1.0, loss, thr)
    J2, g2, c2 = np.zeros(81), np.zeros(9), C.c_double()
    X1, X2, D1, D2 = _d(x1), _d(x2), _d(sc.d1), _d(sc.d2)
    hc.hc_accumulate(v, C.byref(m), _p(X1), _p(X2), _p(D1), _p(D2), C.c_long(len(x1)), C.c_double(sr), C.c_double(1.0),
                     port.LOSS[loss], C.c_double(thr), _p(J2), _p(g2), C.byref(c2))
    J2 = J2.reshape(9, 9)
    assert np.abs(JtJ - J2).max() <= 1e-12 * np.abs(JtJ).max()
    assert np.abs(Jtr - g2).max() <= 1e-11 * np.abs(Jtr).max()
    assert abs(cost - c2.value) <= 1e-12 * cost
    # Cholesky solve and parameter step
    npar = {0: 7, 1: 9, 2: 8, 3: 9}[v]
    x = np.zeros(9)
    hc.hc_llt_solve(npar, _p(np.ascontiguousarray(JtJ.reshape(-1))), C.c_double(1e-3), _p(Jtr), _p(x))
    ref_x = np.linalg.solve(JtJ[:npar, :npar] + 1e-3 * np.eye(npar), Jtr[:npar])
    assert np.allclose(x[:npar], ref_x, rtol=1e-7, atol=1e-12)
