import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdrp_b200 import _native as nv, synth
for cfg, var, shift, iters in (("cfg5_roma_calib", 0, 0, 1000), ("cfg2_calib_shift", 1, 1, 10000), ("cfg1_calib_scale", 0, 0, 1000)):
    b = synth.make_batch(cfg, 300, seed=5)
    o = nv.default_options(); o.max_iterations = o.min_iterations = iters; o.max_epipolar_error, o.max_reproj_error = 2.0, 16.0
    o.estimate_shift = shift; o.loss_type = nv.LOSS["TRUNCATED_CAUCHY"]
    for env in ({}, {"RP_NO_WAVES": "1"}):
        for k in ("RP_NO_WAVES",): os.environ.pop(k, None)
        os.environ.update(env)
        ctx = nv.Context(0)
        ctx.estimate_batch_host(var, b["offsets"], b["x1"], b["x2"], b["d1"], b["d2"], b["cams"], o)
        ms, cn = ctx.last_timing()
        print(cfg, env, "models/pair %.0f exact/pair %.1f (head %d) lm problems/pair %.1f" % (cn["hypotheses"]/300, cn["exact_models"]/300, cn["head_models"], cn["lm_problems"]/300))
        ctx.close()
