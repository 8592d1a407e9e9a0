#!/bin/bash
# LM kernel A/B over alternative builds (build_variants/*.so) x RP_LM_WARP masks:  LIBS="base fold" MASKS="0 3 15" CFGS="..." bash tools/lm_variants.sh
cd "$(dirname "$0")/.."
cp mdrp_b200/librepose_b200.so /tmp/orig.so
IFS=';' read -ra CF <<< "${CFGS:-cfg2_calib_shift 10000;cfg4_varying_focal 10000;cfg1_calib_scale 20000}"
for lib in ${LIBS:-cur}; do
  if [ "$lib" = cur ]; then cp /tmp/orig.so mdrp_b200/librepose_b200.so; else cp build_variants/$lib.so mdrp_b200/librepose_b200.so; fi
  for cfg in "${CF[@]}"; do
    set -- $cfg
    for m in ${MASKS:-0 15}; do
      export L=$lib M=$m C=$1; RP_LM_WARP=$m python bench.py --steps 2 --warmup 2 --no-cpu-baseline --config $1 --pairs $2 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print(os.environ['L'], os.environ['C'], 'mask', os.environ['M'], round(d['value']), 'lo', round(s['lo_refine'],2), 'lo_score_merge', round(s['lo_score_merge'],2), 'final', round(s['final_refine'],2), 'total', round(s['device_total'],1))"
    done
  done
done
cp /tmp/orig.so mdrp_b200/librepose_b200.so
