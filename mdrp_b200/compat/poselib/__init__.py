"""Drop-in `poselib` module for the RePoseD hot path of kocurvik/mdrp.

    import sys; sys.path.insert(0, "<repo>/mdrp_b200/compat"); import poselib

exports the three upstream estimators the reference's demos call (make_pair.py:111, make_video.py:284,
README.md:86-96), the fork names its eval drivers call (eval.py:153, eval_shared_f.py:177,
eval_varying_f.py:168), the minimal solvers and the return types — all running on the B200 through
librepose_b200.so.  Everything else of PoseLib is outside this build.
"""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.insert(0, _root)

from mdrp_b200.api import (  # noqa: E402,F401
    Camera, CameraPose, MonoDepthImagePair, MonoDepthTwoViewGeometry,
    estimate_monodepth_relative_pose, estimate_monodepth_shared_focal_relative_pose,
    estimate_monodepth_varying_focal_relative_pose,
    estimate_monodepth_relative_pose_batch, estimate_monodepth_shared_focal_relative_pose_batch,
    estimate_monodepth_varying_focal_relative_pose_batch,
    estimate_relative_pose_w_mono_depth, estimate_shared_focal_monodepth_relative_pose,
    estimate_varying_focal_monodepth_relative_pose,
    monodepth_pose_3pt, shared_focal_monodepth_pose_3pt, varying_focal_monodepth_pose_4pt,
)

__version__ = "2.0.5+b200"
