#!/usr/bin/env python
"""How many minimal models SHOULD survive the prune (exact count > B0 or exact score < S0, computed with the stage
entry points) vs how many the pipeline sends to the exact kernel, for single pairs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdrp_b200 import _native as nv, synth

ctx = nv.Context(0)
for cfg, var, iters in (("cfg1_calib_scale", 0, 1000), ("cfg5_roma_calib", 0, 1000), ("cfg2_calib_shift", 1, 3000)):
    for idx in range(3):
        s = synth.scene_for(cfg, 900 + idx)
        n = len(s.d1)
        f = s.f1
        x1 = (s.x1 - [640, 480]) / f
        x2 = (s.x2 - [640, 480]) / f
        thr = 2.0 * 0.5 * (1 / s.f1 + 1 / s.f2)
        smp = ctx.sample(n, 0, iters)
        x1h = np.concatenate([x1[smp], np.ones((iters, 3, 1))], axis=2)
        x2h = np.concatenate([x2[smp], np.ones((iters, 3, 1))], axis=2)
        models, counts = ctx.solve(var, x1h, x2h, s.d1[smp], s.d2[smp])
        flat = np.concatenate([models[i, :counts[i]] for i in range(iters)])
        sc, cn = ctx.score(0, flat, x1, x2, thr * thr)
        B0, S0 = cn[:128].max(), sc[:128].min()
        true_surv = int(((cn[128:] > B0) | (sc[128:] < S0)).sum())
        o = nv.default_options()
        o.max_iterations = o.min_iterations = iters
        o.max_epipolar_error, o.max_reproj_error = 2.0, 16.0
        o.estimate_shift = int(var == 1)
        o.loss_type = nv.LOSS["TRUNCATED_CAUCHY"]
        cams = np.array([[s.f1, s.f1, 640, 480, s.f2, s.f2, 640, 480]])
        ctx.estimate_batch_host(var, [0, n], s.x1, s.x2, s.d1, s.d2, cams, o)
        _, c = ctx.last_timing()
        nan_models = int(np.isnan(flat["q"]).any(axis=1).sum())
        print(f"{cfg} #{idx}: models {len(flat)} (pipeline {c['hypotheses']}), B0 {B0}, S0/thr2 {S0 / thr ** 2:.1f}, "
              f"true survivors {true_surv}, pipeline exact beyond head {c['exact_models'] - min(128, len(flat))}, NaN models {nan_models}")
