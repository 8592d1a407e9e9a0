#!/usr/bin/env python
"""GPU path vs the reference wheel at BASELINE.json's full sizes, on the bench's own generator.

    python tools/parity_at_size.py [--pairs 512] [--configs cfg2_calib_shift,...] [--out gpurun_out/parity.json]

For every config: `pairs` pairs of synth.make_batch (the generator bench.py times) through
rp_estimate_batch_host and through the wheel's estimate_monodepth_* (oracle/parity.py), then the
per-pair classification identical / tie (refinements +-1 only) / different, and a dump of the
differing pairs for follow-up.  Test infrastructure: touches oracle/.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_config(ctx, cfg, pairs, seed):
    from mdrp_b200 import _native as nv, synth
    from oracle import parity
    c = synth.CONFIGS[cfg]
    batch = synth.make_batch(cfg, pairs, seed=seed)
    variant = {"calib": nv.CALIB_SHIFT if c["shift"] else nv.CALIB, "shared": nv.SHARED, "varying": nv.VARYING}[c["variant"]]
    o = nv.default_options()
    o.max_iterations = o.min_iterations = c["iters"]
    o.max_epipolar_error, o.max_reproj_error, o.seed = 2.0, 16.0, 0
    o.estimate_shift = int(c["shift"])
    o.loss_type = nv.LOSS["TRUNCATED_CAUCHY"]
    o.loss_scale = 1.0
    models, stats, masks = ctx.estimate_batch_host(variant, batch["offsets"], batch["x1"], batch["x2"], batch["d1"],
                                                   batch["d2"], batch["cams"], o)
    ro = {"max_iterations": c["iters"], "min_iterations": c["iters"], "max_epipolar_error": 2.0,
          "max_reproj_error": 16.0, "seed": 0}
    ref = parity.run_reference(c["variant"], c["shift"], batch, ro, {"loss_type": "TRUNCATED_CAUCHY"})
    rec, cls = parity.compare(ref, batch["offsets"], models, stats, masks)
    rec["config"] = cfg
    rec["cpu_pairs_per_s"] = pairs / ref["seconds"]
    rec["cpu_cores"] = ref["cores"]
    bad = []
    rows = parity.struct_to_rows(models)
    for i in np.nonzero(cls != 0)[0][:40]:
        bad.append({"pair": int(i), "class": int(cls[i]),
                    "ours": [int(stats[i]["refinements"]), int(stats[i]["iterations"]), int(stats[i]["num_inliers"]),
                             float(stats[i]["model_score"])],
                    "ref": [int(v) for v in ref["stats"][i]] + [float(ref["score"][i])],
                    "model_absdiff": float(np.nanmax(np.abs(parity._canon(rows[i]) - parity._canon(ref["models"][i]))))})
    rec["not_identical"] = bad
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=512)
    ap.add_argument("--seed", type=int, default=777)
    ap.add_argument("--configs", default="cfg1_calib_scale,cfg2_calib_shift,cfg3_shared_focal,cfg4_varying_focal,cfg5_roma_calib")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_at_size.json"))
    args = ap.parse_args()
    from mdrp_b200 import _native as nv
    ctx = nv.Context(0)
    out = []
    for cfg in args.configs.split(","):
        pairs = args.pairs if cfg != "cfg5_roma_calib" else max(args.pairs // 2, 16)
        rec = run_config(ctx, cfg, pairs, args.seed)
        print(json.dumps({k: v for k, v in rec.items() if k != "not_identical"}), flush=True)
        out.append(rec)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
