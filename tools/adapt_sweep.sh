#!/bin/bash
# first-pass share of the tensor-core tier: adaptive per pair (RP_TC_ADAPT_PCT = percent of the pair's abandonment
# threshold) against the fixed half (RP_TC_SPLIT=8):  PCTS="105 115 125" bash tools/adapt_sweep.sh
cd "$(dirname "$0")/.."
for cfg in "cfg1_calib_scale 20000" "cfg2_calib_shift 10000" "cfg3_shared_focal 10000" "cfg4_varying_focal 10000" "cfg5_roma_calib 4000" "hard_calib 20000"; do
  set -- $cfg
  for v in fixed8 ${PCTS:-105 115 125}; do
    if [ $v = fixed8 ]; then export RP_TC_SPLIT=8; unset RP_TC_ADAPT_PCT; else unset RP_TC_SPLIT; export RP_TC_ADAPT_PCT=$v; fi
    export V=$v C=$1; python bench.py --steps 2 --warmup 2 --no-cpu-baseline --config $1 --pairs $2 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print(os.environ['C'], os.environ['V'], round(d['value']), 'tc', round(s['tc_kernel'],2), 'bound', round(s['bound_kernel'],2), 'score', round(s['score_minimal'],2))"
  done
done
