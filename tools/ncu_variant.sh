#!/bin/bash
# ncu --set full capture of one kernel launch for alternative builds (build_variants/*.so):
#   gpurun -- 'KERNEL=lm_kernel SKIP=0 bash tools/ncu_variant.sh F D2'
cd "$(dirname "$0")/.."
cp mdrp_b200/librepose_b200.so /tmp/orig.so
K=${KERNEL:-lm_kernel}; S=${SKIP:-0}
for v in "$@"; do
  if [ "$v" = cur ]; then cp /tmp/orig.so mdrp_b200/librepose_b200.so; else cp build_variants/$v.so mdrp_b200/librepose_b200.so; fi   # `cur`: the library as built in-tree
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/ncu_$v \
      python bench.py --no-cpu-baseline --steps 1 --warmup 0 --pairs ${PAIRS:-2000} > gpurun_out/ncu_$v.log 2>&1
done
cp /tmp/orig.so mdrp_b200/librepose_b200.so
