#!/bin/bash
# GPU box: the round-2 measurement set for the current tree.  usage: tools/round2.sh <tag> [tests|notests]
#   GPU tests + smoke, one bench line per BASELINE workload (with clocks, roofline, cpu_baseline, sustained leg),
#   the reference arm of the headline, the ncu launch list and one `--set full` capture of the top kernel of the
#   tree workloads.  Every leg has its own timeout so a hang cannot eat the box.
TAG=${1:-r2h}
mkdir -p gpurun_out
if [ "${2:-tests}" = tests ]; then
  timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/gputests_$TAG.txt
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
fi
for WL in cbox tess20m inst10k smoke; do
  timeout 600 python bench.py --workload $WL > gpurun_out/bench_${WL}_$TAG.json 2> gpurun_out/bench_${WL}_$TAG.err
  python - <<EOF
import json
try:
    d = json.load(open("gpurun_out/bench_${WL}_$TAG.json"))
    print("$WL", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "sust", round(d.get("sustained", {}).get("value", 0)), d["clocks"], "cpu", d.get("cpu_baseline", {}).get("value"), "share", d["roofline"].get("stage_share"))
except Exception as e:
    print("$WL failed", e)
EOF
done
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 300 gpurun_out/bench_ref_$TAG.json
for WL in tess20m inst10k; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${WL}_$TAG.csv \
      python bench.py --workload $WL --steps 1 --warmup 1 --min-seconds 0 --no-cpu-baseline > gpurun_out/launches_${WL}_$TAG.log 2>&1
  timeout 900 bash tools/profile_wl.sh $WL k_trace_fused 1 2 ${WL}_k_trace_fused_$TAG
done
