#!/bin/bash
# first-pass share of the tensor-core tier (RP_TC_SPLIT sixteenths of a pair's correspondences)
for v in 6 8 9 10 11 12; do
  export V=$v; RP_TC_SPLIT=$v python bench.py --steps 2 --warmup 2 --no-cpu-baseline ${1:+--config $1} ${2:+--pairs $2} 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print('split', os.environ.get('V'), '/16', round(d['value']), 'score', round(s['score_minimal'],2), 'tc', round(s['tc_kernel'],2), 'bound', round(s['bound_kernel'],2))"
done
