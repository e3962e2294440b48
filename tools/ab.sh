# usage: bash tools/ab.sh '<params json A>' '<params json B>' ...   (GPU box; prints Mrays/s and stage ms of each)
for P in "$@"; do
  echo "== params $P"
  for r in 1 2; do python bench.py --no-cpu-baseline --steps 5 --params "$P" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2),'ms', {k:round(v*d['ms_per_step'],2) for k,v in d['roofline']['stage_share'].items() if v})
    else: print(l.rstrip()[-300:])
"; done
done
