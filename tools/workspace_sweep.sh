cd /root/repo
for gb in 32 48 80; do for cfg in "cfg2_calib_shift 10000" "cfg5_roma_calib 4000"; do set -- $cfg; export G=$gb C=$1; RP_WORKSPACE_GB=$gb python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config $1 --pairs $2 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); print(os.environ['C'],'ws_gb',os.environ['G'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1))"; done; done
