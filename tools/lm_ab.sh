#!/bin/bash
# A/B of the LO refinement kernels: RP_LM_WARP=1 (one warp per problem) against 0 (one block per problem)
cd "$(dirname "$0")/.."
for cfg in ${CFGS:-"cfg2_calib_shift 10000" "cfg4_varying_focal 10000" "cfg1_calib_scale 20000" "cfg3_shared_focal 10000" "cfg5_roma_calib 4000"}; do
  set -- $cfg
  for v in ${VARS:-0 1}; do
    export V=$v C=$1; RP_LM_WARP=$v python bench.py --steps 2 --warmup 2 --no-cpu-baseline --config $1 --pairs $2 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print(os.environ['C'], 'lm_warp', os.environ['V'], round(d['value']), 'lo', round(s['lo_refine'],2), 'lo_score_merge', round(s['lo_score_merge'],2), 'final', round(s['final_refine'],2), 'total', round(s['device_total'],1), 'parity', d.get('parity'))"
  done
done
