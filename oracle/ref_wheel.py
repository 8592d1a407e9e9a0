"""ctypes view of the reference's PoseLib binary (oracle/_ref) — stage oracles.

TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product never does.

The wheel keeps its C++ dynamic symbols, so
every stage of the hot path has a bit-exact oracle, not only the end-to-end
Python call:

  random_int / draw_sample                so@0x4f87a0 / so@0x4f87f0
  compute_sampson_msac_score (pose / F)   so@0x4f61d0 / so@0x4f65d0
  get_inliers (pose / F)                  so@0x4f7a10 / so@0x4f77f0
  refine_monodepth_relpose (+focal)       so@0x261030 / 0x2592e0 / 0x260fa0
  essential_from_motion                   so@0x1dcb60

Struct layouts follow SURVEY.md Appendix A.2 (validated there against the
binary).
"""
import ctypes as C
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return bool(glob.glob(os.path.join(REF, "poselib", "_core*.so")))


_poselib = None
_lib = None


def poselib():
    """The reference's python module (imported from oracle/_ref)."""
    global _poselib
    if _poselib is None:
        if not available():
            raise RuntimeError("oracle/_ref missing: run python oracle/build_ref.py")
        # Make sure a product module also called `poselib` never shadows it.
        saved = sys.modules.pop("poselib", None)
        sys.path.insert(0, REF)
        try:
            import poselib as _p
        finally:
            sys.path.remove(REF)
        _poselib = _p
        # keep the reference reachable only through this accessor
        sys.modules.pop("poselib", None)
        for k in [k for k in sys.modules if k.startswith("poselib.")]:
            sys.modules["_ref_" + k] = sys.modules.pop(k)
        if saved is not None:
            sys.modules["poselib"] = saved
    return _poselib


def lib():
    global _lib
    if _lib is None:
        poselib()
        path = glob.glob(os.path.join(REF, "poselib", "_core*.so"))[0]
        _lib = C.CDLL(path)
    return _lib


class StdVec(C.Structure):
    """libstdc++ std::vector<T>: three pointers."""
    _fields_ = [("begin", C.c_void_p), ("end", C.c_void_p), ("cap", C.c_void_p)]


def vec_of(arr: np.ndarray) -> StdVec:
    """Fake a std::vector over a C-contiguous numpy buffer (caller keeps arr alive)."""
    assert arr.flags["C_CONTIGUOUS"]
    p = arr.ctypes.data
    return StdVec(p, p + arr.nbytes, p + arr.nbytes)


def aligned_zeros(nbytes: int, align: int = 32) -> np.ndarray:
    raw = np.zeros(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]


class BundleOptions(C.Structure):
    _fields_ = [("max_iterations", C.c_size_t), ("loss_type", C.c_int),
                ("loss_scale", C.c_double), ("gradient_tol", C.c_double),
                ("step_tol", C.c_double), ("initial_lambda", C.c_double),
                ("min_lambda", C.c_double), ("max_lambda", C.c_double),
                ("verbose", C.c_bool)]


class BundleStats(C.Structure):
    _fields_ = [("iterations", C.c_size_t), ("initial_cost", C.c_double),
                ("cost", C.c_double), ("lam", C.c_double),
                ("invalid_steps", C.c_size_t), ("step_norm", C.c_double),
                ("grad_norm", C.c_double)]


LOSS = {"TRIVIAL": 0, "TRUNCATED": 1, "HUBER": 2, "CAUCHY": 3,
        "TRUNCATED_CAUCHY": 4, "TRUNCATED_LE_ZACH": 5}


def bundle_options(max_iterations=100, loss_type="TRIVIAL", loss_scale=1.0,
                   gradient_tol=1e-10, step_tol=1e-8, initial_lambda=1e-3,
                   min_lambda=1e-10, max_lambda=1e10) -> BundleOptions:
    return BundleOptions(max_iterations, LOSS[loss_type], loss_scale, gradient_tol,
                         step_tol, initial_lambda, min_lambda, max_lambda, False)


# ---- geometry buffers -------------------------------------------------------
def geom_buf(q, t, scale=1.0, shift1=0.0, shift2=0.0) -> np.ndarray:
    """MonoDepthTwoViewGeometry: 96 B, 32-aligned (q@0, t@0x20, scale@0x40, shifts)."""
    raw = aligned_zeros(96)
    d = raw.view(np.float64)
    d[0:4] = q
    d[4:7] = t
    d[8] = scale
    d[9] = shift1
    d[10] = shift2
    return raw


def geom_read(raw: np.ndarray):
    d = raw.view(np.float64)
    return d[0:4].copy(), d[4:7].copy(), float(d[8]), float(d[9]), float(d[10])


def pose_buf(q, t) -> np.ndarray:
    raw = aligned_zeros(64)
    d = raw.view(np.float64)
    d[0:4] = q
    d[4:7] = t
    return raw


class ImagePairBuf:
    """MonoDepthImagePair: geometry@0, camera1@0x60, camera2@0x88 (Camera =
    {int model_id,width,height; std::vector<double> params}); SIMPLE_PINHOLE id 0."""

    def __init__(self, q, t, scale, f1, f2, shift1=0.0, shift2=0.0):
        self.raw = aligned_zeros(192)
        d = self.raw.view(np.float64)
        d[0:4] = q
        d[4:7] = t
        d[8] = scale
        d[9] = shift1
        d[10] = shift2
        self.p1 = np.array([f1, 0.0, 0.0])
        self.p2 = np.array([f2, 0.0, 0.0])
        base = self.raw.ctypes.data
        for off, p in ((0x60, self.p1), (0x88, self.p2)):
            C.c_int.from_address(base + off).value = 0
            C.c_int.from_address(base + off + 4).value = 0
            C.c_int.from_address(base + off + 8).value = 0
            v = StdVec.from_address(base + off + 16)
            v.begin = p.ctypes.data
            v.end = p.ctypes.data + 24
            v.cap = p.ctypes.data + 24

    def read(self):
        d = self.raw.view(np.float64)
        return (d[0:4].copy(), d[4:7].copy(), float(d[8]), float(self.p1[0]),
                float(self.p2[0]))


# ---- stage oracles ----------------------------------------------------------
def random_int(state: int):
    """poselib::random_int(size_t&) → (int value, new state)."""
    L = lib()
    f = L._ZN7poselib10random_intERm
    f.restype = C.c_int
    f.argtypes = [C.POINTER(C.c_uint64)]
    s = C.c_uint64(state)
    v = f(C.byref(s))
    return v, s.value


def draw_sample(sample_sz: int, n: int, state: int):
    """poselib::draw_sample(sample_sz, N, vector<size_t>*, size_t& state)."""
    L = lib()
    f = L._ZN7poselib11draw_sampleEmmPSt6vectorImSaImEERm
    f.restype = None
    f.argtypes = [C.c_size_t, C.c_size_t, C.POINTER(StdVec), C.POINTER(C.c_uint64)]
    out = np.zeros(sample_sz, dtype=np.uint64)
    v = vec_of(out)
    s = C.c_uint64(state)
    f(sample_sz, n, C.byref(v), C.byref(s))
    return out.astype(np.int64), s.value


class RandomSamplerBuf(C.Structure):
    """poselib::RandomSampler: num_data@0, sample_sz@8, state@0x10, use_prosac@0x18,
    max_prosac_iterations@0x20, sample_k@0x28, subset_sz@0x30, growth vector@0x38."""
    _fields_ = [("num_data", C.c_size_t), ("sample_sz", C.c_size_t), ("state", C.c_uint64),
                ("use_prosac", C.c_uint8), ("_pad", C.c_uint8 * 7), ("max_prosac_iterations", C.c_size_t),
                ("sample_k", C.c_size_t), ("subset_sz", C.c_size_t), ("growth", StdVec)]


def generate_samples(num_data: int, sample_sz: int, seed: int, use_prosac: bool, max_prosac_iterations: int,
                     iters: int):
    """`iters` calls of RandomSampler::generate_sample so@0x4f8970 on a sampler set up like the constructor does
    (initialize_prosac so@0x4f8a20 when use_prosac).  Returns (samples [iters, sample_sz], growth table)."""
    L = lib()
    init = L._ZN7poselib13RandomSampler17initialize_prosacEv
    init.restype = None
    init.argtypes = [C.POINTER(RandomSamplerBuf)]
    gen = L._ZN7poselib13RandomSampler15generate_sampleEPSt6vectorImSaImEE
    gen.restype = None
    gen.argtypes = [C.POINTER(RandomSamplerBuf), C.POINTER(StdVec)]
    growth = np.zeros(max(num_data, sample_sz), dtype=np.uint64)  # pre-sized: the binary must not reallocate it
    sb = RandomSamplerBuf()
    sb.num_data, sb.sample_sz, sb.state = num_data, sample_sz, seed
    sb.use_prosac, sb.max_prosac_iterations = int(bool(use_prosac)), max_prosac_iterations
    sb.growth = vec_of(growth)
    if use_prosac:
        init(C.byref(sb))
    out = np.zeros((iters, sample_sz), dtype=np.int64)
    cur = np.zeros(sample_sz, dtype=np.uint64)
    v = vec_of(cur)
    for i in range(iters):
        gen(C.byref(sb), C.byref(v))
        out[i] = cur
    return out, growth.astype(np.int64)


def _pts(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def msac_score_pose(q, t, x1, x2, sq_thr):
    """compute_sampson_msac_score(CameraPose,…) → (score, inlier_count)."""
    L = lib()
    f = getattr(L, "_ZN7poselib26compute_sampson_msac_scoreERKNS_10CameraPoseERKSt6vector"
                   "IN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEESaIS6_EESA_dPm")
    f.restype = C.c_double
    f.argtypes = [C.c_void_p, C.POINTER(StdVec), C.POINTER(StdVec), C.c_double,
                  C.POINTER(C.c_size_t)]
    pb = pose_buf(q, t)
    x1 = _pts(x1)
    x2 = _pts(x2)
    v1, v2 = vec_of(x1), vec_of(x2)
    cnt = C.c_size_t(0)
    s = f(pb.ctypes.data, C.byref(v1), C.byref(v2), sq_thr, C.byref(cnt))
    return s, cnt.value


def msac_score_F(F, x1, x2, sq_thr):
    """compute_sampson_msac_score(Matrix3d F,…) (F stored column-major)."""
    L = lib()
    f = getattr(L, "_ZN7poselib26compute_sampson_msac_scoreERKN5Eigen6MatrixIdLi3ELi3ELi0ELi3ELi3EEE"
                   "RKSt6vectorINS1_IdLi2ELi1ELi0ELi2ELi1EEESaIS6_EESA_dPm")
    f.restype = C.c_double
    f.argtypes = [C.c_void_p, C.POINTER(StdVec), C.POINTER(StdVec), C.c_double,
                  C.POINTER(C.c_size_t)]
    fb = aligned_zeros(72)
    fb.view(np.float64)[:] = np.asarray(F, dtype=np.float64).T.reshape(-1)
    x1 = _pts(x1)
    x2 = _pts(x2)
    v1, v2 = vec_of(x1), vec_of(x2)
    cnt = C.c_size_t(0)
    s = f(fb.ctypes.data, C.byref(v1), C.byref(v2), sq_thr, C.byref(cnt))
    return s, cnt.value


def _read_charvec(v: StdVec, n: int) -> np.ndarray:
    if not v.begin:
        return np.zeros(0, dtype=bool)
    m = v.end - v.begin
    return np.ctypeslib.as_array(C.cast(v.begin, C.POINTER(C.c_uint8)), shape=(m,)).astype(bool).copy()


def get_inliers_pose(q, t, x1, x2, sq_thr):
    L = lib()
    f = getattr(L, "_ZN7poselib11get_inliersERKNS_10CameraPoseERKSt6vectorIN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEE"
                   "SaIS6_EESA_dPS3_IcSaIcEE")
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.POINTER(StdVec), C.POINTER(StdVec), C.c_double, C.POINTER(StdVec)]
    pb = pose_buf(q, t)
    x1 = _pts(x1)
    x2 = _pts(x2)
    v1, v2 = vec_of(x1), vec_of(x2)
    out = StdVec(None, None, None)
    f(pb.ctypes.data, C.byref(v1), C.byref(v2), sq_thr, C.byref(out))
    return _read_charvec(out, len(x1))


def get_inliers_F(F, x1, x2, sq_thr):
    L = lib()
    f = getattr(L, "_ZN7poselib11get_inliersERKN5Eigen6MatrixIdLi3ELi3ELi0ELi3ELi3EEERKSt6vector"
                   "INS1_IdLi2ELi1ELi0ELi2ELi1EEESaIS6_EESA_dPS5_IcSaIcEE")
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.POINTER(StdVec), C.POINTER(StdVec), C.c_double, C.POINTER(StdVec)]
    fb = aligned_zeros(72)
    fb.view(np.float64)[:] = np.asarray(F, dtype=np.float64).T.reshape(-1)
    x1 = _pts(x1)
    x2 = _pts(x2)
    v1, v2 = vec_of(x1), vec_of(x2)
    out = StdVec(None, None, None)
    f(fb.ctypes.data, C.byref(v1), C.byref(v2), sq_thr, C.byref(out))
    return _read_charvec(out, len(x1))


def essential_from_motion(q, t):
    L = lib()
    f = L._ZN7poselib21essential_from_motionERKNS_10CameraPoseEPN5Eigen6MatrixIdLi3ELi3ELi0ELi3ELi3EEE
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p]
    pb = pose_buf(q, t)
    eb = aligned_zeros(72)
    f(pb.ctypes.data, eb.ctypes.data)
    return eb.view(np.float64).reshape(3, 3).T.copy()


def refine_calib(x1, x2, d1, d2, q, t, scale, shift1, shift2, scale_reproj,
                 weight_sampson, bopt: BundleOptions, estimate_shift: bool):
    """refine_monodepth_relpose(x1,x2,d1,d2,&geom,scale_reproj,weight_sampson,opt,shift,weights={})
    → ((q,t,scale,shift1,shift2), BundleStats)."""
    L = lib()
    f = getattr(L, "_ZN7poselib24refine_monodepth_relposeERKSt6vectorIN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEE"
                   "SaIS3_EES7_RKS0_IdSaIdEESB_PNS_24MonoDepthTwoViewGeometryEddRKNS_13BundleOptionsEbSB_")
    f.restype = BundleStats
    f.argtypes = [C.POINTER(StdVec), C.POINTER(StdVec), C.POINTER(StdVec), C.POINTER(StdVec),
                  C.c_void_p, C.c_double, C.c_double, C.POINTER(BundleOptions), C.c_bool,
                  C.POINTER(StdVec)]
    x1, x2 = _pts(x1), _pts(x2)
    d1, d2 = _pts(d1), _pts(d2)
    v1, v2, w1, w2 = vec_of(x1), vec_of(x2), vec_of(d1), vec_of(d2)
    g = geom_buf(q, t, scale, shift1, shift2)
    wv = StdVec(None, None, None)
    st = f(C.byref(v1), C.byref(v2), C.byref(w1), C.byref(w2), g.ctypes.data,
           scale_reproj, weight_sampson, C.byref(bopt), estimate_shift, C.byref(wv))
    return geom_read(g), st


def _refine_focal(sym, x1, x2, d1, d2, q, t, scale, f1, f2, scale_reproj, weight_sampson, bopt):
    L = lib()
    f = getattr(L, sym)
    f.restype = BundleStats
    f.argtypes = [C.POINTER(StdVec), C.POINTER(StdVec), C.POINTER(StdVec), C.POINTER(StdVec),
                  C.c_void_p, C.c_double, C.c_double, C.POINTER(BundleOptions), C.POINTER(StdVec)]
    x1, x2 = _pts(x1), _pts(x2)
    d1, d2 = _pts(d1), _pts(d2)
    v1, v2, w1, w2 = vec_of(x1), vec_of(x2), vec_of(d1), vec_of(d2)
    ip = ImagePairBuf(q, t, scale, f1, f2)
    wv = StdVec(None, None, None)
    st = f(C.byref(v1), C.byref(v2), C.byref(w1), C.byref(w2), ip.raw.ctypes.data,
           scale_reproj, weight_sampson, C.byref(bopt), C.byref(wv))
    return ip.read(), st


def refine_shared(x1, x2, d1, d2, q, t, scale, f, scale_reproj, weight_sampson, bopt):
    sym = ("_ZN7poselib37refine_monodepth_shared_focal_relposeERKSt6vectorIN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEE"
           "SaIS3_EES7_RKS0_IdSaIdEESB_PNS_18MonoDepthImagePairEddRKNS_13BundleOptionsESB_")
    return _refine_focal(sym, x1, x2, d1, d2, q, t, scale, f, f, scale_reproj, weight_sampson, bopt)


def refine_varying(x1, x2, d1, d2, q, t, scale, f1, f2, scale_reproj, weight_sampson, bopt):
    sym = ("_ZN7poselib38refine_monodepth_varying_focal_relposeERKSt6vectorIN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEE"
           "SaIS3_EES7_RKS0_IdSaIdEESB_PNS_18MonoDepthImagePairEddRKNS_13BundleOptionsESB_")
    return _refine_focal(sym, x1, x2, d1, d2, q, t, scale, f1, f2, scale_reproj, weight_sampson, bopt)
