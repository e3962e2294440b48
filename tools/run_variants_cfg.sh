# usage: bash tools/run_variants_cfg.sh [variant ...]: cbox bench + configs 3/5 for each library variant ("" = default)
for v in "" "$@"; do
  if [ -n "$v" ]; then export KRR_WFPT_LIB=$PWD/kiraray_b200/lib/libkrr_wfpt_$v.so; fi
  echo "== variant [$v]"
  python bench.py --no-cpu-baseline --steps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cbox', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2),'ms')
"
  python tools/bench_configs.py 3 5 --scale 0.5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['workload'][:12], d['ms_frames'], d['stage_ms'])
"
done
