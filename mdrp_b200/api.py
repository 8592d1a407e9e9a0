"""Host-side mirror of the reference's Python surface for the RePoseD hot path.

Same names, argument meaning, return objects and `info` keys as the PoseLib binding the
reference calls (whl:_core.pyi:446-501; call sites /root/reference/make_pair.py:111,
make_video.py:284, README.md:86-96), plus the PoseLib-mdrp fork names the eval drivers use
(/root/reference/eval.py:153, eval_shared_f.py:177, eval_varying_f.py:168) and batched entry
points, which are the unit the GPU actually wants.  All numerics run in librepose_b200.so
(CUDA, sm_100a); there is no CPU fallback.

Error behaviour: unknown dict keys are ignored; N < 3 returns the identity
model with iterations = 0, model_score = DBL_MAX and an all-False mask; x1/x2/d1/d2 length
mismatches raise ValueError (the reference silently reads out of bounds there).
"""
import sys
import threading

import numpy as np

from . import _native as nv

DBL_MAX = sys.float_info.max

_CAMERA_MODELS = {"SIMPLE_PINHOLE": 0, "PINHOLE": 1}


class Camera:
    """poselib.Camera for the two models the reference's callers use (whl:_core.pyi:76-123)."""

    def __init__(self, model="SIMPLE_PINHOLE", params=None, width=-1, height=-1):
        if model not in _CAMERA_MODELS:
            raise ValueError(f"camera model {model!r} is outside this build (SIMPLE_PINHOLE / PINHOLE; "
                             "undistort on the host first)")
        self._model = model
        self.model_id = _CAMERA_MODELS[model]
        self.params = [float(p) for p in (params if params is not None else [1.0, 0.0, 0.0])]
        self.width, self.height = int(width), int(height)

    @classmethod
    def from_any(cls, cam):
        if isinstance(cam, Camera):
            return cam
        if isinstance(cam, dict):
            return cls(cam.get("model", "SIMPLE_PINHOLE"), cam.get("params"), cam.get("width", -1),
                       cam.get("height", -1))
        # duck-typed poselib.Camera
        return cls(cam.model_name(), list(cam.params), cam.width, cam.height)

    def model_name(self):
        return self._model

    def focal_x(self):
        return self.params[0]

    def focal_y(self):
        return self.params[0] if self._model == "SIMPLE_PINHOLE" else self.params[1]

    def focal(self):
        return 0.5 * (self.focal_x() + self.focal_y())

    def principal_point(self):
        return np.array(self.params[1:3] if self._model == "SIMPLE_PINHOLE" else self.params[2:4])

    def fxfycxcy(self):
        pp = self.principal_point()
        return [self.focal_x(), self.focal_y(), float(pp[0]), float(pp[1])]

    def unproject(self, x):
        x = np.asarray(x, dtype=np.float64)
        f = np.array([self.focal_x(), self.focal_y()])
        return (x - self.principal_point()) / f

    def project(self, x):
        x = np.asarray(x, dtype=np.float64)
        f = np.array([self.focal_x(), self.focal_y()])
        return x * f + self.principal_point()

    def __repr__(self):
        return f"[{self._model} {self.width} {self.height} {self.params}]"


class CameraPose:
    """poselib.CameraPose: q (w first), t; R/Rt derived (whl:_core.pyi:125-152)."""

    def __init__(self, q=(1.0, 0.0, 0.0, 0.0), t=(0.0, 0.0, 0.0)):
        self.q = np.array(q, dtype=np.float64)
        self.t = np.array(t, dtype=np.float64)

    @property
    def R(self):
        w, x, y, z = self.q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])

    @property
    def Rt(self):
        return np.concatenate([self.R, self.t[:, None]], axis=1)

    def center(self):
        return -self.R.T @ self.t

    def __repr__(self):
        return f"[q: {self.q}, t: {self.t}]"


class MonoDepthTwoViewGeometry:
    """X2 = scale (d2+shift2) K2^-1 x2 = R (d1+shift1) K1^-1 x1 + t (whl:_core.pyi:178-204)."""

    def __init__(self, pose=None, scale=1.0, shift1=0.0, shift2=0.0):
        self.pose = pose if pose is not None else CameraPose()
        self.scale, self.shift1, self.shift2 = float(scale), float(shift1), float(shift2)

    def __repr__(self):
        return f"[pose: {self.pose}, scale: {self.scale}, shift1: {self.shift1}, shift2: {self.shift2}]"


class MonoDepthImagePair:
    """geometry + two SIMPLE_PINHOLE cameras [f, 0, 0] (whl:_core.pyi:171-176).  `.pose` aliases
    geometry.pose: the eval drivers of the fork read image_pair.pose (eval_shared_f.py:84)."""

    def __init__(self, geometry=None, camera1=None, camera2=None):
        self.geometry = geometry if geometry is not None else MonoDepthTwoViewGeometry()
        self.camera1 = camera1 if camera1 is not None else Camera()
        self.camera2 = camera2 if camera2 is not None else Camera()

    @property
    def pose(self):
        return self.geometry.pose


# ---- option dicts --------------------------------------------------------------------------------
_RANSAC_KEYS = ("max_iterations", "min_iterations", "dyn_num_trials_mult", "success_prob", "max_reproj_error",
                "max_epipolar_error", "seed")


def make_options(ransac_opt=None, bundle_opt=None, focal_variant=False) -> nv.Options:
    """dict -> rp_options with the binding's defaults (whl:METADATA:71-107); unknown keys ignored."""
    ransac_opt = ransac_opt or {}
    bundle_opt = bundle_opt or {}
    o = nv.default_options()
    for k in _RANSAC_KEYS:
        if k in ransac_opt:
            setattr(o, k, type(getattr(o, k))(ransac_opt[k]))
    # PROSAC (RandomSampler::initialize_prosac so@0x4f8a20): the caller passes correspondences sorted by
    # decreasing quality, exactly as with the reference
    o.progressive_sampling = int(bool(ransac_opt.get("progressive_sampling", False)))
    if "max_prosac_iterations" in ransac_opt:
        o.max_prosac_iterations = int(ransac_opt["max_prosac_iterations"])
    o.estimate_shift = int(bool(ransac_opt.get("monodepth_estimate_shift", False)))
    o.weight_sampson = float(np.float32(ransac_opt.get("monodepth_weight_sampson", 1.0)))
    if "max_iterations" in bundle_opt:
        o.bundle_max_iterations = int(bundle_opt["max_iterations"])
    # the binding matches the loss name case-insensitively and keeps its default (CAUCHY) for a name it
    # does not know (verified on the wheel: 'cauchy' == 'CAUCHY', 'bogus' -> CAUCHY)
    lt = bundle_opt.get("loss_type", "CAUCHY")
    if isinstance(lt, str):
        if lt.upper() == "TRUNCATED_LE_ZACH":
            raise NotImplementedError("TRUNCATED_LE_ZACH loss is outside this build")
        o.loss_type = nv.LOSS.get(lt.upper(), nv.LOSS["CAUCHY"])
    else:
        o.loss_type = int(lt)
    # focal variants honour the user's loss_scale (default 0.5*max_epipolar_error); the calibrated
    # variant overwrites it with half the normalised threshold (SURVEY.md §8b "Errors")
    o.loss_scale = float(bundle_opt.get("loss_scale", 0.5 * o.max_epipolar_error if focal_variant else 1.0))
    for k in ("gradient_tol", "step_tol", "initial_lambda", "min_lambda", "max_lambda"):
        if k in bundle_opt:
            setattr(o, k, float(bundle_opt[k]))
    return o


# ---- context per device ----------------------------------------------------------------------------
_ctx = {}
_ctx_lock = threading.Lock()


def context(device: int = 0) -> nv.Context:
    with _ctx_lock:
        if device not in _ctx:
            _ctx[device] = nv.Context(device)
        return _ctx[device]


def _pack(list_x1, list_x2, list_d1, list_d2):
    n = [len(d) for d in list_d1]
    for a, b, c, d in zip(list_x1, list_x2, list_d1, list_d2):
        a = np.asarray(a)
        if a.ndim != 2 or a.shape[1] != 2 or np.asarray(b).shape != a.shape:
            raise TypeError("points must be [N, 2] arrays")
        if not (len(a) == len(c) == len(d)):
            raise ValueError("points and depths must have the same length")
    offsets = np.zeros(len(n) + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(n)
    cat = lambda xs, shape: (np.concatenate([np.asarray(x, dtype=np.float64).reshape(shape) for x in xs])
                             if len(xs) else np.zeros(shape if shape[0] != -1 else (0,) + shape[1:]))
    x1 = cat(list_x1, (-1, 2))
    x2 = cat(list_x2, (-1, 2))
    d1 = cat(list_d1, (-1,))
    d2 = cat(list_d2, (-1,))
    return offsets, x1, x2, d1, d2


def _info(stats_row, mask):
    return {"refinements": int(stats_row["refinements"]), "iterations": int(stats_row["iterations"]),
            "num_inliers": int(stats_row["num_inliers"]), "inlier_ratio": float(stats_row["inlier_ratio"]),
            "model_score": float(stats_row["model_score"]), "inliers": [bool(b) for b in mask]}


def _geometry(m):
    return MonoDepthTwoViewGeometry(CameraPose(m["q"], m["t"]), m["scale"], m["shift1"], m["shift2"])


# ---- batched entry points (the unit of GPU work) -----------------------------------------------------
def estimate_monodepth_relative_pose_batch(points2D_1, points2D_2, depth_1, depth_2, cameras1, cameras2,
                                           ransac_opt=None, bundle_opt=None, device=0):
    """Lists (one entry per image pair) in, list of (MonoDepthTwoViewGeometry, info) out."""
    offsets, x1, x2, d1, d2 = _pack(points2D_1, points2D_2, depth_1, depth_2)
    cams = np.array([Camera.from_any(a).fxfycxcy() + Camera.from_any(b).fxfycxcy()
                     for a, b in zip(cameras1, cameras2)], dtype=np.float64).reshape(-1, 8)
    opt = make_options(ransac_opt, bundle_opt, focal_variant=False)
    variant = nv.CALIB_SHIFT if opt.estimate_shift else nv.CALIB
    models, stats, masks = context(device).estimate_batch_host(variant, offsets, x1, x2, d1, d2, cams, opt)
    return [(_geometry(models[i]), _info(stats[i], masks[offsets[i]:offsets[i + 1]])) for i in range(len(models))]


def _focal_batch(variant, points2D_1, points2D_2, depth_1, depth_2, ransac_opt, bundle_opt, device):
    offsets, x1, x2, d1, d2 = _pack(points2D_1, points2D_2, depth_1, depth_2)
    opt = make_options(ransac_opt, bundle_opt, focal_variant=True)
    models, stats, masks = context(device).estimate_batch_host(variant, offsets, x1, x2, d1, d2, None, opt)
    out = []
    for i, m in enumerate(models):
        pair = MonoDepthImagePair(_geometry(m), Camera("SIMPLE_PINHOLE", [m["f1"], 0.0, 0.0]),
                                  Camera("SIMPLE_PINHOLE", [m["f2"], 0.0, 0.0]))
        out.append((pair, _info(stats[i], masks[offsets[i]:offsets[i + 1]])))
    return out


def estimate_monodepth_shared_focal_relative_pose_batch(points2D_1, points2D_2, depth_1, depth_2,
                                                        ransac_opt=None, bundle_opt=None, device=0):
    return _focal_batch(nv.SHARED, points2D_1, points2D_2, depth_1, depth_2, ransac_opt, bundle_opt, device)


def estimate_monodepth_varying_focal_relative_pose_batch(points2D_1, points2D_2, depth_1, depth_2,
                                                         ransac_opt=None, bundle_opt=None, device=0):
    return _focal_batch(nv.VARYING, points2D_1, points2D_2, depth_1, depth_2, ransac_opt, bundle_opt, device)


# ---- the reference's single-pair surface ---------------------------------------------------------------
def estimate_monodepth_relative_pose(points2D_1, points2D_2, depth_1, depth_2, camera1, camera2,
                                     ransac_opt={}, bundle_opt={}, initial_pose=None):
    """poselib.estimate_monodepth_relative_pose (whl:_core.pyi:446-475).  `initial_pose` is accepted
    and ignored, exactly like the reference (ransac_monodepth_relpose so@0x228c20 resets it)."""
    return estimate_monodepth_relative_pose_batch([points2D_1], [points2D_2], [depth_1], [depth_2], [camera1],
                                                  [camera2], ransac_opt, bundle_opt)[0]


def estimate_monodepth_shared_focal_relative_pose(points2D_1, points2D_2, depth_1, depth_2, ransac_opt={},
                                                  bundle_opt={}, initial_image_pair=None):
    """poselib.estimate_monodepth_shared_focal_relative_pose (whl:_core.pyi:477-488)."""
    return estimate_monodepth_shared_focal_relative_pose_batch([points2D_1], [points2D_2], [depth_1], [depth_2],
                                                               ransac_opt, bundle_opt)[0]


def estimate_monodepth_varying_focal_relative_pose(points2D_1, points2D_2, depth_1, depth_2, ransac_opt={},
                                                   bundle_opt={}, initial_image_pair=None):
    """poselib.estimate_monodepth_varying_focal_relative_pose (whl:_core.pyi:490-501)."""
    return estimate_monodepth_varying_focal_relative_pose_batch([points2D_1], [points2D_2], [depth_1], [depth_2],
                                                                ransac_opt, bundle_opt)[0]


# ---- PoseLib-mdrp fork names used by eval.py / eval_shared_f.py / eval_varying_f.py --------------------
def _fork_ransac(ransac_dict):
    """Experiment flags of eval.py:105-123 -> upstream keys (SURVEY.md §8b fork-API row)."""
    r = dict(ransac_dict)
    r["monodepth_estimate_shift"] = bool(r.get("solver_shift", False)) and bool(r.get("use_ours", False)) \
        and not bool(r.get("use_p3p", False))
    if "weight_sampson" in r:
        r["monodepth_weight_sampson"] = r["weight_sampson"]
    return r


def estimate_relative_pose_w_mono_depth(kp1, kp2, d, camera1, camera2, ransac_dict={}, bundle_dict={}):
    """eval.py:153 — depths arrive as one [N,2] array; returns (CameraPose-like, info)."""
    d = np.asarray(d, dtype=np.float64)
    geom, info = estimate_monodepth_relative_pose(kp1, kp2, d[:, 0], d[:, 1], camera1, camera2,
                                                  _fork_ransac(ransac_dict), bundle_dict)
    return geom.pose, info


def estimate_shared_focal_monodepth_relative_pose(kp1, kp2, d, ransac_dict={}, bundle_dict={}):
    """eval_shared_f.py:177."""
    d = np.asarray(d, dtype=np.float64)
    return estimate_monodepth_shared_focal_relative_pose(kp1, kp2, d[:, 0], d[:, 1], _fork_ransac(ransac_dict),
                                                         bundle_dict)


def estimate_varying_focal_monodepth_relative_pose(kp1, kp2, d, ransac_dict={}, bundle_dict={}):
    """eval_varying_f.py:168."""
    d = np.asarray(d, dtype=np.float64)
    return estimate_monodepth_varying_focal_relative_pose(kp1, kp2, d[:, 0], d[:, 1], _fork_ransac(ransac_dict),
                                                          bundle_dict)


# ---- minimal solvers exposed by the reference (whl:_core.pyi:614, :871, :914) ---------------------------
def _solver(variant, x1, x2, d1, d2):
    x1 = np.asarray(x1, dtype=np.float64).reshape(1, 3, 3)
    x2 = np.asarray(x2, dtype=np.float64).reshape(1, 3, 3)
    models, counts = context(0).solve(variant, x1, x2, np.asarray(d1, dtype=np.float64).reshape(1, 3),
                                      np.asarray(d2, dtype=np.float64).reshape(1, 3))
    return [models[0, k] for k in range(int(counts[0]))]


def monodepth_pose_3pt(x1, x2, d1, d2):
    return [_geometry(m) for m in _solver(nv.CALIB_SHIFT, x1, x2, d1, d2)]


def shared_focal_monodepth_pose_3pt(x1, x2, d1, d2):
    return [MonoDepthImagePair(_geometry(m), Camera("SIMPLE_PINHOLE", [m["f1"], 0, 0]),
                               Camera("SIMPLE_PINHOLE", [m["f2"], 0, 0])) for m in _solver(nv.SHARED, x1, x2, d1, d2)]


def varying_focal_monodepth_pose_4pt(x1, x2, d1, d2):
    return [MonoDepthImagePair(_geometry(m), Camera("SIMPLE_PINHOLE", [m["f1"], 0, 0]),
                               Camera("SIMPLE_PINHOLE", [m["f2"], 0, 0])) for m in _solver(nv.VARYING, x1, x2, d1, d2)]
