#!/bin/bash
# exact-head size sweep (RP_HEAD) on the bench workload: pairs/s and the scoring-stage split
for h in 32 64 96 128; do
  export H=$h; RP_HEAD=$h python bench.py --steps 2 --warmup 2 --no-cpu-baseline ${1:+--config $1} 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print('head', os.environ.get('H'), round(d['value']), 'score', round(s['score_minimal'],2), 'tc', round(s['tc_kernel'],2), 'bound', round(s['bound_kernel'],2), 'exact_frac', round(d['roofline']['exact_models_fraction'],4))"
done
