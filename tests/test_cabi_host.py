"""CPU tier: the C-ABI library loads and exports every symbol include/repose_b200.h declares; the
product fails loudly without a CUDA device; host-side option / argument handling of the Python
mirror (no compute calls)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from mdrp_b200 import build
    return build.build_native()


def test_header_symbols_exported(lib_path):
    header = open(os.path.join(ROOT, "include", "repose_b200.h")).read()
    declared = re.findall(r"RP_API\s+[\w\s\*]+?\b(rp_\w+)\s*\(", header)
    assert len(declared) >= 13
    L = C.CDLL(lib_path)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    from mdrp_b200 import _native as nv
    assert sorted(declared) == sorted(nv.EXPORTS)


def test_library_is_sm100a_and_has_no_oracle_dependency(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", lib_path], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "poselib" not in ldd


def test_struct_layouts_match_header():
    from mdrp_b200 import _native as nv
    assert C.sizeof(nv.Model) == 96 and C.sizeof(nv.Stats) == 40
    assert C.sizeof(nv.Options) == 8 * 2 + 8 * 4 + 8 + 8 + 8 + 8 + 8 + 8 * 6 + 8
    assert nv.Options.progressive_sampling.offset == 60 and nv.Options.max_prosac_iterations.offset == 136
    assert C.sizeof(nv.BundleOptions) == 8 + 8 + 8 * 6
    assert C.sizeof(nv.BundleStats) == 56


def test_ctypes_structs_match_the_c_header(tmp_path):
    """sizeof / offsetof of every struct of include/repose_b200.h as gcc sees them vs the ctypes mirror."""
    from mdrp_b200 import _native as nv
    structs = {"rp_model": nv.Model, "rp_stats": nv.Stats, "rp_options": nv.Options,
               "rp_bundle_options": nv.BundleOptions, "rp_bundle_stats": nv.BundleStats}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "repose_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            cfield = {"lam": "lambda"}.get(fname, fname)
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {cfield}));')
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_default_options_are_poselib_defaults(lib_path):
    from mdrp_b200 import _native as nv
    o = nv.default_options()
    assert (o.max_iterations, o.min_iterations) == (100000, 1000)
    assert (o.dyn_num_trials_mult, o.success_prob) == (3.0, 0.9999)
    assert (o.max_reproj_error, o.max_epipolar_error) == (12.0, 1.0)
    assert o.bundle_max_iterations == 100 and o.loss_type == nv.LOSS["CAUCHY"]
    assert (o.progressive_sampling, o.max_prosac_iterations) == (0, 100000)
    assert (o.gradient_tol, o.step_tol, o.initial_lambda, o.min_lambda, o.max_lambda) == (1e-10, 1e-8, 1e-3, 1e-10, 1e10)


def test_fails_loudly_without_gpu(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mdrp_b200 import _native as nv
    with pytest.raises(nv.NativeError):
        nv.Context(0)
    from mdrp_b200 import api
    with pytest.raises(nv.NativeError):
        api.estimate_monodepth_relative_pose(np.zeros((5, 2)), np.zeros((5, 2)), np.ones(5), np.ones(5),
                                             {"model": "SIMPLE_PINHOLE", "params": [1, 0, 0]},
                                             {"model": "SIMPLE_PINHOLE", "params": [1, 0, 0]})


def test_option_dict_mapping(lib_path):
    from mdrp_b200 import _native as nv, api
    o = api.make_options({"max_iterations": 10000, "min_iterations": 10000, "max_epipolar_error": 2.0,
                          "max_reproj_error": 16.0, "seed": 7, "monodepth_estimate_shift": True,
                          "lo_iterations": 25, "weight_sampson": 3.0, "unknown_key": 1},
                         {"loss_type": "TRUNCATED_CAUCHY", "verbose": False})
    assert (o.max_iterations, o.min_iterations, o.seed, o.estimate_shift) == (10000, 10000, 7, 1)
    assert o.weight_sampson == 1.0  # `weight_sampson` is not a key of the upstream binding (make_pair.py:31-33)
    assert o.loss_type == nv.LOSS["TRUNCATED_CAUCHY"] and o.bundle_max_iterations == 100
    # focal variants default loss_scale to half the epipolar threshold (whl:METADATA:143)
    assert api.make_options({"max_epipolar_error": 2.0}, {}, focal_variant=True).loss_scale == 1.0
    assert api.make_options({"max_epipolar_error": 2.0}, {"loss_scale": 0.3}, focal_variant=True).loss_scale == 0.3
    p = api.make_options({"progressive_sampling": True, "max_prosac_iterations": 5000}, {})
    assert (p.progressive_sampling, p.max_prosac_iterations) == (1, 5000)
    assert api.make_options({"progressive_sampling": False}, {}).progressive_sampling == 0
    # loss names: case-insensitive, an unknown name keeps the binding's default CAUCHY (checked against the wheel in
    # test_loss_name_parsing_matches_reference below)
    assert api.make_options({}, {"loss_type": "truncated_cauchy"}).loss_type == nv.LOSS["TRUNCATED_CAUCHY"]
    assert api.make_options({}, {"loss_type": "bogus"}).loss_type == nv.LOSS["CAUCHY"]
    assert api.make_options({}, {}).loss_type == nv.LOSS["CAUCHY"]
    with pytest.raises(NotImplementedError):
        api.make_options({}, {"loss_type": "truncated_le_zach"})
    # fork flags of eval.py:105-123
    r = api._fork_ransac({"use_ours": True, "solver_shift": True, "use_p3p": False, "weight_sampson": 1.0})
    assert r["monodepth_estimate_shift"] is True
    assert api._fork_ransac({"use_p3p": True, "use_ours": False, "solver_shift": False})["monodepth_estimate_shift"] is False


def test_camera_models():
    from mdrp_b200 import api
    c = api.Camera.from_any({"model": "PINHOLE", "width": -1, "height": -1, "params": [700.0, 900.0, 320.0, 240.0]})
    assert c.focal() == 800.0 and c.fxfycxcy() == [700.0, 900.0, 320.0, 240.0]
    s = api.Camera.from_any({"model": "SIMPLE_PINHOLE", "width": 1, "height": 1, "params": [800.0, 640.0, 480.0]})
    assert s.fxfycxcy() == [800.0, 800.0, 640.0, 480.0] and s.model_name() == "SIMPLE_PINHOLE"
    x = np.array([[650.0, 470.0]])
    assert np.allclose(s.project(s.unproject(x)), x)
    with pytest.raises(ValueError):
        api.Camera.from_any({"model": "OPENCV", "params": [1] * 8})


def test_argument_errors():
    from mdrp_b200 import api
    with pytest.raises(TypeError):
        api._pack([np.zeros((4, 3))], [np.zeros((4, 3))], [np.ones(4)], [np.ones(4)])
    with pytest.raises(ValueError):
        api._pack([np.zeros((4, 2))], [np.zeros((4, 2))], [np.ones(3)], [np.ones(4)])
    off, x1, x2, d1, d2 = api._pack([np.zeros((4, 2), dtype=np.float32), np.zeros((0, 2))],
                                    [np.zeros((4, 2)), np.zeros((0, 2))], [[1, 2, 3, 4], []], [np.ones(4), np.ones(0)])
    assert off.tolist() == [0, 4, 4] and x1.dtype == np.float64 and d1.tolist() == [1, 2, 3, 4]


def test_loss_name_parsing_matches_reference(ref, port):
    """The wheel parses bundle_opt['loss_type'] case-insensitively and falls back to CAUCHY for unknown names:
    its result for 'cauchy' / 'bogus' equals its result for 'CAUCHY', and differs from 'TRIVIAL'."""
    from mdrp_b200 import synth
    pl = ref.poselib()
    sc = synth.scene_for("cfg1_calib_scale", 3, n=200)
    c1, c2 = sc.camera_dicts()
    ro = {"max_iterations": 100, "min_iterations": 100, "max_epipolar_error": 2.0, "max_reproj_error": 16.0}

    def run(name):
        g, _ = pl.estimate_monodepth_relative_pose(sc.x1, sc.x2, sc.d1, sc.d2, c1, c2, ro, {"loss_type": name})
        return np.r_[np.array(g.pose.q), np.array(g.pose.t), g.scale]
    base = run("CAUCHY")
    assert np.array_equal(run("cauchy"), base) and np.array_equal(run("bogus"), base)
    assert not np.array_equal(run("TRIVIAL"), base)
    assert np.array_equal(run("truncated_cauchy"), run("TRUNCATED_CAUCHY"))


def test_context_serialises_concurrent_calls():
    """nv.Context owns one device workspace: its native calls go through one lock."""
    import threading
    from mdrp_b200 import _native as nv
    c = nv.Context.__new__(nv.Context)
    c._lock = threading.RLock()
    c._lib = None
    c._h = None
    inside, seen = [0], []

    def fake(*a):
        inside[0] += 1
        seen.append(inside[0])
        import time
        time.sleep(0.01)
        inside[0] -= 1
        return 0
    ts = [threading.Thread(target=lambda: c._call(fake, 1)) for _ in range(6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert seen == [1] * 6
