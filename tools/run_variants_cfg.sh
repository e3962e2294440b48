# usage: bash tools/run_variants_cfg.sh "<bench_configs args>" [variant ...]   ("" = the default library)
A="$1"; shift
for v in "" "$@"; do
  if [ -n "$v" ]; then export KRR_WFPT_LIB=$PWD/kiraray_b200/lib/libkrr_wfpt_$v.so; fi
  echo "== variant [$v]"; python tools/bench_configs.py $A 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['workload'][:16], round(d['Mrays_per_s'],1), 'Mrays/s', round(d['ms_per_frame'],2),'ms', d['stage_ms'])
    else: print(l.rstrip()[-300:])
"
done
