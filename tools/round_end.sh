#!/bin/bash
# GPU box: what the driver runs at round end, for the current tree, plus the profile set.  usage: tools/round_end.sh <tag>
TAG=${1:-r1n}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.json
python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 400 gpurun_out/bench_ref_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
bash tools/profile_one.sh k_trace_fused 1 k_trace_fused_$TAG
bash tools/profile_one.sh k_scatter 0 k_scatter_$TAG
bash tools/profile_one.sh k_trace_closest 1 k_trace_closest_$TAG
