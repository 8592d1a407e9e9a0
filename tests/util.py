"""Shared helpers of the parity tests."""
import numpy as np

VARIANT_ID = {"calib": 0, "calib_shift": 1, "shared": 2, "varying": 3}


def model_vec(m):
    """12-vector (q4, t3, scale, shift1, shift2, f1, f2) from any model representation."""
    if isinstance(m, np.void) or (isinstance(m, np.ndarray) and m.dtype.names):
        return np.r_[m["q"], m["t"], m["scale"], m["shift1"], m["shift2"], m["f1"], m["f2"]]
    if isinstance(m, tuple):
        return np.r_[m[0], m[1], m[2:]]
    if hasattr(m, "q"):
        return np.r_[np.array(m.q), np.array(m.t), m.scale, m.shift1, m.shift2, m.f1, m.f2]
    return np.asarray(m, dtype=np.float64)


def canon(v):
    """Quaternion sign made canonical (q and -q are the same rotation)."""
    v = np.array(v, dtype=np.float64)
    if v[0] < 0 or (v[0] == 0 and v[1] < 0):
        v[:4] = -v[:4]
    return v


def models_close(a, b, rtol=1e-6, atol=1e-9):
    a, b = canon(model_vec(a)), canon(model_vec(b))
    return np.allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)


def same_set(ref, got, rtol=1e-6, atol=1e-9):
    """Unordered comparison of two solution lists, NaN / invalid solutions dropped on both sides."""
    ref = [canon(model_vec(m)) for m in ref if np.isfinite(model_vec(m)).all()]
    got = [canon(model_vec(m)) for m in got if np.isfinite(model_vec(m)).all()]
    used = set()
    for a in ref:
        hit = None
        for j, b in enumerate(got):
            if j not in used and np.allclose(a, b, rtol=rtol, atol=atol):
                hit = j
                break
        if hit is None:
            return False
        used.add(hit)
    return len(used) == len(got)


def dedup(models, rtol=1e-6):
    out = []
    for m in models:
        v = canon(model_vec(m))
        if not np.isfinite(v).all():
            continue
        if not any(np.allclose(v, o, rtol=rtol, atol=1e-9) for o in out):
            out.append(v)
    return out


def struct_models(arr):
    """[n,12] float array -> structured MODEL_DTYPE array."""
    from mdrp_b200 import _native as nv
    arr = np.asarray(arr, dtype=np.float64).reshape(-1, 12)
    out = np.zeros(len(arr), dtype=nv.MODEL_DTYPE)
    out["q"], out["t"] = arr[:, :4], arr[:, 4:7]
    out["scale"], out["shift1"], out["shift2"], out["f1"], out["f2"] = arr[:, 7], arr[:, 8], arr[:, 9], arr[:, 10], arr[:, 11]
    return out


def rot_err_deg(q, R_gt):
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    c = (np.trace(R_gt.T @ R) - 1) / 2
    return float(np.degrees(np.arccos(np.clip(c, -1, 1))))


def trans_err_deg(t, t_gt):
    n = np.linalg.norm(t) * np.linalg.norm(t_gt)
    if not np.isfinite(n) or n < 1e-12:
        return 180.0
    return float(np.degrees(np.arccos(np.clip(np.dot(t, t_gt) / n, -1, 1))))


def maa(errs, max_deg=10):
    """mAA(10 deg) of /root/reference/utils/eval_utils.py:49-52; NaN counts as 180."""
    e = np.array([180.0 if not np.isfinite(x) else x for x in errs])
    return float(np.mean([np.mean(e < t) for t in range(1, max_deg + 1)]) * 100.0)


LOSS_NAMES = ["TRIVIAL", "TRUNCATED", "HUBER", "CAUCHY", "TRUNCATED_CAUCHY"]


def extra_cases(g):
    """(key, variant, dict of options) for every case of tests/golden/extra.npz (make_golden_extra.py)."""
    for key in sorted(k[:-6] for k in g.files if k.endswith("_model")):
        io, fo = g[key + "_iopts"], g[key + "_fopts"]
        yield key, int(key[1]), dict(iters=int(io[0]), min_iters=int(io[1]), seed=int(io[2]), prosac=bool(io[3]),
                                     max_prosac=int(io[4]), bundle_iters=int(io[5]), loss=LOSS_NAMES[int(io[6])],
                                     t_epi=float(fo[0]), t_rep=float(fo[1]), weight_sampson=float(fo[2]))
