cd /root/repo
python -m pytest tests -m gpu -q 2>&1 | tail -2
for h in 32 64 128; do
  for spec in "cfg2_calib_shift 10000" "cfg1_calib_scale 20000" "hard_calib 20000" "cfg5_roma_calib 4000" "cfg4_varying_focal 10000"; do
    set -- $spec
    RP_HEAD=$h python bench.py --config $1 --pairs $2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/h.json 2> gpurun_out/h.err
    python - "$1" $h <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/h.json")); s = d["stage_ms_per_step"]
    print("head %3s %-20s value %8.0f e2e %8.0f | score %.1f (bound %.1f) total %.1f ms exact %.4f" % (sys.argv[2], sys.argv[1], d["value"], d["e2e"]["value"], s["score_minimal"], s["bound_kernel"], s["device_total"], d["roofline"]["exact_models_fraction"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
  done
done
