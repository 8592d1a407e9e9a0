#!/bin/bash
# Benchmarks alternative builds of librepose_b200.so (build_variants/*.so, made with RP_LM_* / RP_*_MIN_BLOCKS
# build knobs) back to back on one GPU box:  gpurun -- 'bash tools/variant_bench.sh A B C'
cd "$(dirname "$0")/.."
cp mdrp_b200/librepose_b200.so /tmp/orig.so
for v in "$@"; do
  if [ "$v" = cur ]; then cp /tmp/orig.so mdrp_b200/librepose_b200.so; else cp build_variants/$v.so mdrp_b200/librepose_b200.so; fi   # `cur`: the library as built in-tree
  python bench.py --no-cpu-baseline --steps 2 --warmup 3 ${BENCH_ARGS} > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/variant_{v}.json"))
    s = d["stage_ms_per_step"]
    print(v, "value %.0f e2e %.0f | solve %.1f score %.1f (tc %.1f bound %.1f) lo %.1f final %.1f total %.1f" % (
        d["value"], d["e2e"]["value"], s["solve"], s["score_minimal"], s["tc_kernel"], s["bound_kernel"], s["lo_refine"], s["final_refine"], s["device_total"]))
except Exception as e:
    print(v, "failed", e)
PY
done
cp /tmp/orig.so mdrp_b200/librepose_b200.so
