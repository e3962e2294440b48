#!/bin/bash
# GPU box: the round-2 measurement set for the current tree.  usage: tools/round2.sh <tag> [tests|notests] [ncu|noncu]
#   GPU tests + smoke, one bench line per BASELINE workload (with clocks, roofline, cpu_baseline, sustained leg),
#   the reference arm of the headline, and per workload: the ncu launch list of a short run (time + DRAM bytes of
#   every launch) and one `--set full` capture of the first launches of the traced / shaded stages.
#   Every leg has its own timeout so a hang cannot eat the box.
TAG=${1:-r2h}
mkdir -p gpurun_out
if [ "${2:-tests}" = tests ]; then
  timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/gputests_$TAG.txt
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
fi
for WL in cbox tess20m inst10k smoke; do
  timeout 600 python bench.py --workload $WL > gpurun_out/bench_${WL}_$TAG.json 2> gpurun_out/bench_${WL}_$TAG.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${WL}_$TAG.json"))
    print("$WL", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "sust", round(d.get("sustained", {}).get("value", 0)), d["clocks"], "cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("$WL failed", e)
PY
done
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 300 gpurun_out/bench_ref_$TAG.json
if [ "${3:-ncu}" = ncu ]; then
  KERNELS='regex:k_trace|k_scatter|k_generate|k_resolve|k_film|k_handle|k_begin_frame|k_fold|k_medium|k_tail'
  for WL in cbox tess20m inst10k smoke; do
    timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KERNELS" -c 6000 --csv \
        --log-file gpurun_out/launches_${WL}_$TAG.csv python bench.py --workload $WL --steps 1 --warmup 1 --min-seconds 0 --no-cpu-baseline > gpurun_out/launches_${WL}_$TAG.log 2>&1
    # first launches of the traced / shaded stages of the SECOND render (the first one pays the lazy module load)
    case $WL in cbox) SKIP=168;; tess20m) SKIP=42;; inst10k) SKIP=11;; smoke) SKIP=77;; esac
    timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_trace_fused|k_trace_closest|k_scatter|k_trace_shadow|k_medium' -s $SKIP -c 6 -f \
        -o gpurun_out/prof_${WL}_$TAG python bench.py --workload $WL --steps 1 --warmup 1 --min-seconds 0 --no-cpu-baseline --params '{"bands": 1}' > gpurun_out/prof_${WL}_$TAG.log 2>&1
    # the report itself is tens of MB per captured launch and gpurun_out/ is capped at 64 MiB: export what
    # tools/ncu_summary.py and tools/ncu_lines.py read (raw page of every launch; source page of the tree trace kernel), drop the report
    ncu -i gpurun_out/prof_${WL}_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${WL}_$TAG.raw.csv 2>/dev/null
    case $WL in tess20m|inst10k) ncu -i gpurun_out/prof_${WL}_$TAG.ncu-rep --page source --csv --print-source sass -k regex:k_trace_fused > gpurun_out/prof_${WL}_$TAG.src.csv 2>/dev/null;; esac
    ls -la gpurun_out/prof_${WL}_$TAG.ncu-rep; rm -f gpurun_out/prof_${WL}_$TAG.ncu-rep
  done
fi
