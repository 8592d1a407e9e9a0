#!/bin/bash
# mid stage of the minimal-model scoring (RP_HEAD_SMALL exact models, then the cascade over positions up to RP_MID_END,
# then the bulk); RP_MID_END=0 is the plain exact head (RP_HEAD, 128):  CFG="cfg2_calib_shift 10000" bash tools/mid_sweep.sh
cd "$(dirname "$0")/.."
set -- ${CFG:-cfg2_calib_shift 10000}
[ ${#COMBOS[@]} -gt 0 ] || COMBOS=("32 0" "32 256" "32 96,512" "32 128,1024" "32 64,256,1024" "0 32,256" "0 64,512")
for v in "${COMBOS[@]}"; do
  set -- $1 $2 $v
  export RP_HEAD_SMALL=$3 RP_MID_END=$4 C=$1
  python bench.py --steps 2 --warmup 2 --no-cpu-baseline --config $1 --pairs $2 2>/tmp/mid_err.txt | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print(os.environ['C'], 'head_small', os.environ['RP_HEAD_SMALL'], 'mid_end', os.environ['RP_MID_END'], round(d['value']), 'score', round(s['score_minimal'],2), 'tc', round(s['tc_kernel'],2), 'bound', round(s['bound_kernel'],2), 'total', round(s['device_total'],1), 'exact_frac', round(d['roofline'].get('exact_models_fraction',0),4))"
  [ -s /tmp/mid_err.txt ] && tail -3 /tmp/mid_err.txt; set -- $1 $2
done
