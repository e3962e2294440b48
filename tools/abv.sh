#!/bin/bash
# GPU box: A/B of library variants (kiraray_b200/build.py build_variant) on one bench workload.
# usage: bash tools/abv.sh <workload> [variant ...]    ("base" = the default library); env PARAMS='{"..."}' extra pass params
WL=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset KRR_WFPT_LIB; else export KRR_WFPT_LIB=$PWD/kiraray_b200/lib/libkrr_wfpt_$v.so; fi
  timeout 300 python bench.py --workload $WL --no-cpu-baseline --min-seconds 0 --steps 3 --warmup 2 ${PARAMS:+--params "$PARAMS"} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$WL [$v]', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2),'ms', 'launches', d['gpu_launches'], {k:round(v*d['ms_per_step'],2) for k,v in d['roofline']['stage_share'].items() if v})
    else: print(l.rstrip()[-300:])
"
done
