"""The driver shim: tools/eval_synth.py runs the experiment strings of eval.py /
eval_shared_f.py / eval_varying_f.py through this package's `poselib` surface and through the reference wheel; the
reference's table metrics must agree (north_star: pose AUC within 0.5 points)."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tool():
    spec = importlib.util.spec_from_file_location("eval_synth", os.path.join(ROOT, "tools", "eval_synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("driver,config", [("calib", "hard_calib"), ("shared", "cfg3_shared_focal"), ("varying", "cfg4_varying_focal")])
def test_eval_synth_matches_reference_backend(ctx, ref, driver, config):
    t = _tool()
    quiet = lambda *_: None  # noqa: E731
    ours = t.run(driver, config, 24, 400, 500, "b200", out=quiet)
    theirs = t.run(driver, config, 24, 400, 500, "reference", out=quiet)
    assert set(ours) == set(theirs) == set(t.EXPERIMENTS[driver])
    for exp in ours:
        assert abs(ours[exp][1] - theirs[exp][1]) <= 0.5, (exp, ours[exp], theirs[exp])            # mAA(10 deg), points
        assert abs(ours[exp][0] - theirs[exp][0]) <= 1e-6 + 1e-3 * theirs[exp][0], (exp, ours[exp], theirs[exp])
        if ours[exp][2] is not None:
            assert abs(ours[exp][2] - theirs[exp][2]) <= 1e-6 + 1e-3 * theirs[exp][2]
