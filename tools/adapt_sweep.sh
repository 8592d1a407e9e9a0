#!/bin/bash
# adaptive first-pass share of the tensor-core tier: percent of the pair's abandonment threshold (RP_TC_ADAPT_PCT)
for cfg in "cfg1_calib_scale 20000" "cfg2_calib_shift 10000" "cfg3_shared_focal 10000" "cfg4_varying_focal 10000" "cfg5_roma_calib 4000" "hard_calib 20000"; do
  set -- $cfg
  for v in ${PCTS:-125 135 150}; do
    export V=$v C=$1; RP_TC_ADAPT_PCT=$v python bench.py --steps 2 --warmup 2 --no-cpu-baseline --config $1 --pairs $2 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print(os.environ['C'], 'pct', os.environ['V'], round(d['value']), 'tc', round(s['tc_kernel'],2), 'bound', round(s['bound_kernel'],2))"
  done
done
