import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdrp_b200 import _native as nv, synth
ctx = nv.Context(0)
b = synth.make_batch("cfg2_calib_shift", 2000, seed=1)
o = nv.default_options(); o.max_iterations = o.min_iterations = 10000; o.max_epipolar_error, o.max_reproj_error = 2.0, 16.0
o.estimate_shift = 1; o.loss_type = nv.LOSS["TRUNCATED_CAUCHY"]
for _ in range(2):
    models, stats, masks = ctx.estimate_batch_host(1, b["offsets"], b["x1"], b["x2"], b["d1"], b["d2"], b["cams"], o)
ms, cn = ctx.last_timing()
print(ms); print(cn)
print("LM problems/pair", cn["lm_problems"]/2000, "iters/problem", cn["lm_iterations"]/cn["lm_problems"], "refinements mean", stats["refinements"].mean())
