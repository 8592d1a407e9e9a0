"""ctypes binding of oracle/librepose_oracle.so (the C restatement).

TEST INFRASTRUCTURE ONLY — see repose_oracle.c header.  Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "librepose_oracle.so")

CALIB, CALIB_SHIFT, SHARED, VARYING = 0, 1, 2, 3
LOSS = {"TRIVIAL": 0, "TRUNCATED": 1, "HUBER": 2, "CAUCHY": 3, "TRUNCATED_CAUCHY": 4,
        "TRUNCATED_LE_ZACH": 5}


class Model(C.Structure):
    _fields_ = [("q", C.c_double * 4), ("t", C.c_double * 3), ("scale", C.c_double),
                ("shift1", C.c_double), ("shift2", C.c_double), ("f1", C.c_double),
                ("f2", C.c_double)]

    def as_tuple(self):
        return (np.array(self.q), np.array(self.t), self.scale, self.shift1, self.shift2,
                self.f1, self.f2)


def make_model(q=(1, 0, 0, 0), t=(0, 0, 0), scale=1.0, shift1=0.0, shift2=0.0, f1=1.0, f2=1.0):
    m = Model()
    m.q[:] = list(map(float, q))
    m.t[:] = list(map(float, t))
    m.scale, m.shift1, m.shift2, m.f1, m.f2 = scale, shift1, shift2, f1, f2
    return m


class BundleOpt(C.Structure):
    _fields_ = [("max_iterations", C.c_int64), ("loss_type", C.c_int), ("loss_scale", C.c_double),
                ("gradient_tol", C.c_double), ("step_tol", C.c_double), ("initial_lambda", C.c_double),
                ("min_lambda", C.c_double), ("max_lambda", C.c_double)]


class BundleStats(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("initial_cost", C.c_double), ("cost", C.c_double),
                ("lam", C.c_double), ("invalid_steps", C.c_int64), ("step_norm", C.c_double),
                ("grad_norm", C.c_double)]


class RansacOpt(C.Structure):
    _fields_ = [("max_iterations", C.c_int64), ("min_iterations", C.c_int64),
                ("dyn_num_trials_mult", C.c_double), ("success_prob", C.c_double),
                ("max_reproj_error", C.c_double), ("max_epipolar_error", C.c_double),
                ("seed", C.c_uint64), ("estimate_shift", C.c_int), ("weight_sampson", C.c_double),
                ("progressive_sampling", C.c_int), ("max_prosac_iterations", C.c_int64)]


class RansacStats(C.Structure):
    _fields_ = [("refinements", C.c_int64), ("iterations", C.c_int64), ("num_inliers", C.c_int64),
                ("inlier_ratio", C.c_double), ("model_score", C.c_double)]


def bundle_opt(max_iterations=100, loss_type="CAUCHY", loss_scale=1.0, gradient_tol=1e-10,
               step_tol=1e-8, initial_lambda=1e-3, min_lambda=1e-10, max_lambda=1e10):
    lt = LOSS[loss_type] if isinstance(loss_type, str) else int(loss_type)
    return BundleOpt(max_iterations, lt, loss_scale, gradient_tol, step_tol, initial_lambda,
                     min_lambda, max_lambda)


def ransac_opt(max_iterations=100000, min_iterations=1000, dyn_num_trials_mult=3.0,
               success_prob=0.9999, max_reproj_error=12.0, max_epipolar_error=1.0, seed=0,
               estimate_shift=False, weight_sampson=1.0, progressive_sampling=False, max_prosac_iterations=100000):
    return RansacOpt(max_iterations, min_iterations, dyn_num_trials_mult, success_prob,
                     max_reproj_error, max_epipolar_error, seed, int(estimate_shift), weight_sampson,
                     int(progressive_sampling), max_prosac_iterations)


def build(force=False):
    src = os.path.join(HERE, "repose_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
    return SO


_lib = None
DP = C.POINTER(C.c_double)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(DP)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        L.ro_random_int.restype = C.c_int
        L.ro_random_int.argtypes = [C.POINTER(C.c_uint64)]
        L.ro_draw_sample.restype = None
        L.ro_draw_sample.argtypes = [C.c_size_t, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)]
        L.ro_prosac_growth.restype = None
        L.ro_prosac_growth.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ro_generate_samples.restype = None
        L.ro_generate_samples.argtypes = [C.c_size_t, C.c_size_t, C.c_uint64, C.c_int, C.c_size_t, C.c_size_t,
                                          C.POINTER(C.c_size_t)]
        L.ro_essential_from_motion.argtypes = [DP, DP, DP]
        L.ro_msac_score_pose.restype = C.c_double
        L.ro_msac_score_pose.argtypes = [DP, DP, DP, DP, C.c_size_t, C.c_double, C.POINTER(C.c_size_t)]
        L.ro_msac_score_F.restype = C.c_double
        L.ro_msac_score_F.argtypes = [DP, DP, DP, C.c_size_t, C.c_double, C.POINTER(C.c_size_t)]
        L.ro_get_inliers_pose.restype = C.c_size_t
        L.ro_get_inliers_pose.argtypes = [DP, DP, DP, DP, C.c_size_t, C.c_double, C.c_char_p]
        L.ro_get_inliers_F.restype = C.c_size_t
        L.ro_get_inliers_F.argtypes = [DP, DP, DP, C.c_size_t, C.c_double, C.c_char_p]
        L.ro_fundamental_from_model.argtypes = [C.POINTER(Model), DP]
        L.ro_solve_cubic_single_real.restype = C.c_int
        L.ro_solve_cubic_single_real.argtypes = [C.c_double] * 3 + [DP]
        L.ro_solve_quartic_real.restype = C.c_int
        L.ro_solve_quartic_real.argtypes = [C.c_double] * 4 + [DP]
        L.ro_p3p.restype = C.c_int
        L.ro_p3p.argtypes = [DP, DP, DP, DP]
        for name in ("ro_solve_calib_scale", "ro_solve_calib_shift", "ro_solve_varying_focal",
                     "ro_solve_shared_focal"):
            if hasattr(L, name):
                f = getattr(L, name)
                f.restype = C.c_int
                f.argtypes = [DP, DP, DP, DP, C.POINTER(Model)]
        if hasattr(L, "ro_refine"):
            L.ro_refine.restype = None
            L.ro_refine.argtypes = [C.c_int, DP, DP, DP, DP, C.c_size_t, C.POINTER(Model), C.c_double,
                                    C.c_double, C.POINTER(BundleOpt), C.POINTER(BundleStats)]
        if hasattr(L, "ro_cost"):
            L.ro_cost.restype = C.c_double
            L.ro_cost.argtypes = [C.c_int, DP, DP, DP, DP, C.c_size_t, C.POINTER(Model), C.c_double,
                                  C.c_double, C.c_int, C.c_double]
        if hasattr(L, "ro_accumulate"):
            L.ro_accumulate.restype = None
            L.ro_accumulate.argtypes = [C.c_int, DP, DP, DP, DP, C.c_size_t, C.POINTER(Model), C.c_double,
                                        C.c_double, C.c_int, C.c_double, DP, DP]
        if hasattr(L, "ro_ransac"):
            L.ro_ransac.restype = None
            L.ro_ransac.argtypes = [C.c_int, DP, DP, DP, DP, C.c_size_t, C.POINTER(RansacOpt),
                                    C.POINTER(Model), C.POINTER(RansacStats), C.c_char_p]
        if hasattr(L, "ro_estimate"):
            L.ro_estimate.restype = None
            L.ro_estimate.argtypes = [C.c_int, DP, DP, DP, DP, C.c_size_t, DP, DP,
                                      C.POINTER(RansacOpt), C.POINTER(BundleOpt), C.POINTER(Model),
                                      C.POINTER(RansacStats), C.c_char_p]
        _lib = L
    return _lib


# ---- stage wrappers ---------------------------------------------------------
def random_int(state):
    s = C.c_uint64(state)
    v = lib().ro_random_int(C.byref(s))
    return v, s.value


def draw_sample(sample_sz, n, state):
    s = C.c_uint64(state)
    out = (C.c_size_t * sample_sz)()
    lib().ro_draw_sample(sample_sz, n, C.byref(s), out)
    return np.array(list(out), dtype=np.int64), s.value


def prosac_growth(num_data, sample_sz, max_prosac_iterations):
    out = (C.c_size_t * max(num_data, sample_sz))()
    lib().ro_prosac_growth(num_data, sample_sz, max_prosac_iterations, out)
    return np.array(list(out), dtype=np.int64)


def generate_samples(num_data, sample_sz, seed, use_prosac, max_prosac_iterations, iters):
    out = (C.c_size_t * (iters * sample_sz))()
    lib().ro_generate_samples(num_data, sample_sz, seed, int(bool(use_prosac)), max_prosac_iterations, iters, out)
    return np.array(list(out), dtype=np.int64).reshape(iters, sample_sz)


def essential_from_motion(q, t):
    E = np.zeros(9)
    q, t = _d(q), _d(t)
    lib().ro_essential_from_motion(_p(q), _p(t), _p(E))
    return E.reshape(3, 3)


def msac_score_pose(q, t, x1, x2, sq_thr):
    q, t, x1, x2 = _d(q), _d(t), _d(x1), _d(x2)
    c = C.c_size_t(0)
    s = lib().ro_msac_score_pose(_p(q), _p(t), _p(x1), _p(x2), len(x1), sq_thr, C.byref(c))
    return s, c.value


def msac_score_F(F, x1, x2, sq_thr):
    F, x1, x2 = _d(F).reshape(-1), _d(x1), _d(x2)
    c = C.c_size_t(0)
    s = lib().ro_msac_score_F(_p(F), _p(x1), _p(x2), len(x1), sq_thr, C.byref(c))
    return s, c.value


def get_inliers_pose(q, t, x1, x2, sq_thr):
    q, t, x1, x2 = _d(q), _d(t), _d(x1), _d(x2)
    buf = C.create_string_buffer(len(x1))
    lib().ro_get_inliers_pose(_p(q), _p(t), _p(x1), _p(x2), len(x1), sq_thr, buf)
    return np.frombuffer(buf.raw, dtype=np.uint8).astype(bool)


def get_inliers_F(F, x1, x2, sq_thr):
    F, x1, x2 = _d(F).reshape(-1), _d(x1), _d(x2)
    buf = C.create_string_buffer(len(x1))
    lib().ro_get_inliers_F(_p(F), _p(x1), _p(x2), len(x1), sq_thr, buf)
    return np.frombuffer(buf.raw, dtype=np.uint8).astype(bool)


def fundamental_from_model(m: Model):
    F = np.zeros(9)
    lib().ro_fundamental_from_model(C.byref(m), _p(F))
    return F.reshape(3, 3)


def solve_cubic_single_real(c2, c1, c0):
    r = C.c_double(0)
    g = lib().ro_solve_cubic_single_real(c2, c1, c0, C.byref(r))
    return bool(g), r.value


def solve_quartic_real(b, c, d, e):
    r = (C.c_double * 4)()
    n = lib().ro_solve_quartic_real(b, c, d, e, r)
    return [r[i] for i in range(n)]


def p3p(x, X):
    x, X = _d(x), _d(X)
    q = np.zeros((4, 4))
    t = np.zeros((4, 3))
    n = lib().ro_p3p(_p(x), _p(X), _p(q), _p(t))
    return [(q[i].copy(), t[i].copy()) for i in range(n)]


def _solve(name, x1h, x2h, d1, d2):
    x1h, x2h, d1, d2 = _d(x1h), _d(x2h), _d(d1), _d(d2)
    out = (Model * 8)()
    n = getattr(lib(), name)(_p(x1h), _p(x2h), _p(d1), _p(d2), out)
    return [out[i].as_tuple() for i in range(n)]


def solve_calib_scale(x1h, x2h, d1, d2):
    return _solve("ro_solve_calib_scale", x1h, x2h, d1, d2)


def solve_calib_shift(x1h, x2h, d1, d2):
    return _solve("ro_solve_calib_shift", x1h, x2h, d1, d2)


def solve_varying_focal(x1h, x2h, d1, d2):
    return _solve("ro_solve_varying_focal", x1h, x2h, d1, d2)


def solve_shared_focal(x1h, x2h, d1, d2):
    return _solve("ro_solve_shared_focal", x1h, x2h, d1, d2)


def cost(variant, x1, x2, d1, d2, m: Model, scale_reproj, weight_sampson, loss_type, loss_scale):
    x1, x2, d1, d2 = _d(x1), _d(x2), _d(d1), _d(d2)
    lt = LOSS[loss_type] if isinstance(loss_type, str) else int(loss_type)
    return lib().ro_cost(variant, _p(x1), _p(x2), _p(d1), _p(d2), len(x1), C.byref(m), scale_reproj,
                         weight_sampson, lt, loss_scale)


def accumulate(variant, x1, x2, d1, d2, m: Model, scale_reproj, weight_sampson, loss_type, loss_scale):
    x1, x2, d1, d2 = _d(x1), _d(x2), _d(d1), _d(d2)
    lt = LOSS[loss_type] if isinstance(loss_type, str) else int(loss_type)
    JtJ = np.zeros(81)
    Jtr = np.zeros(9)
    lib().ro_accumulate(variant, _p(x1), _p(x2), _p(d1), _p(d2), len(x1), C.byref(m), scale_reproj,
                        weight_sampson, lt, loss_scale, _p(JtJ), _p(Jtr))
    return JtJ.reshape(9, 9), Jtr


def refine(variant, x1, x2, d1, d2, m: Model, scale_reproj, weight_sampson, bopt: BundleOpt):
    x1, x2, d1, d2 = _d(x1), _d(x2), _d(d1), _d(d2)
    out = Model.from_buffer_copy(m)
    st = BundleStats()
    lib().ro_refine(variant, _p(x1), _p(x2), _p(d1), _p(d2), len(x1), C.byref(out), scale_reproj,
                    weight_sampson, C.byref(bopt), C.byref(st))
    return out, st


def ransac(variant, x1, x2, d1, d2, ropt: RansacOpt):
    x1, x2, d1, d2 = _d(x1), _d(x2), _d(d1), _d(d2)
    m = make_model()
    st = RansacStats()
    buf = C.create_string_buffer(max(len(x1), 1))
    lib().ro_ransac(variant, _p(x1), _p(x2), _p(d1), _p(d2), len(x1), C.byref(ropt), C.byref(m),
                    C.byref(st), buf)
    return m, st, np.frombuffer(buf.raw, dtype=np.uint8)[:len(x1)].astype(bool)


def estimate(variant, x1, x2, d1, d2, cam1, cam2, ropt: RansacOpt, bopt: BundleOpt):
    """cam1/cam2: (fx, fy, cx, cy) for the calibrated variant, ignored (None) for focal variants."""
    x1, x2, d1, d2 = _d(x1), _d(x2), _d(d1), _d(d2)
    c1 = _d(cam1 if cam1 is not None else [1, 1, 0, 0])
    c2 = _d(cam2 if cam2 is not None else [1, 1, 0, 0])
    m = make_model()
    st = RansacStats()
    buf = C.create_string_buffer(max(len(x1), 1))
    lib().ro_estimate(variant, _p(x1), _p(x2), _p(d1), _p(d2), len(x1), _p(c1), _p(c2),
                      C.byref(ropt), C.byref(bopt), C.byref(m), C.byref(st), buf)
    return m, st, np.frombuffer(buf.raw, dtype=np.uint8)[:len(x1)].astype(bool)
