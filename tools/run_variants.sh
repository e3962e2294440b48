# usage: bash tools/run_variants.sh [variant ...]   ("" = the default library)
python -m pytest tests -m gpu -q 2>&1 | grep -E "names|passed|failed|AssertionError: \(" | cut -c1-1500
for v in "" "$@"; do
  if [ -n "$v" ]; then export KRR_WFPT_LIB=$PWD/kiraray_b200/lib/libkrr_wfpt_$v.so; fi
  echo "== variant [$v]"; python bench.py --no-cpu-baseline --steps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2),'ms', {k:round(v*d['ms_per_step'],2) for k,v in d['roofline']['stage_share'].items()})
    else: print(l.rstrip()[-300:])
"
done
