#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one `ncu --set full` capture of the hot kernels.
# usage: tools/profile_r1.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --spp 2 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/launches_$TAG.log 2>&1
for K in k_trace_closest k_scatter k_trace_fused; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 2 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/prof_${K}_$TAG.log 2>&1
done
ls -la gpurun_out
