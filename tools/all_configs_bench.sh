#!/bin/bash
# One short bench line per BASELINE config (the headline is cfg2; the others are reported in README):
#   gpurun -- 'bash tools/all_configs_bench.sh'
cd "$(dirname "$0")/.."
for spec in "cfg1_calib_scale 20000" "cfg2_calib_shift 10000" "cfg3_shared_focal 10000" "cfg4_varying_focal 10000" "cfg5_roma_calib 4000" "hard_calib 20000"; do
  set -- $spec
  python bench.py --config $1 --pairs $2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/cfg_$1.json 2> gpurun_out/cfg_$1.err
  python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/cfg_{sys.argv[1]}.json"))
    s = d["stage_ms_per_step"]
    print("%-20s value %8.0f e2e %8.0f pairs/s | solve %.1f score %.1f (tc %.1f bound %.1f) lo %.1f final %.1f total %.1f ms" % (
        sys.argv[1], d["value"], d["e2e"]["value"], s["solve"], s["score_minimal"], s["tc_kernel"], s["bound_kernel"], s["lo_refine"], s["final_refine"], s["device_total"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
