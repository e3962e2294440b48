#!/bin/bash
# GPU box: `ncu --set full` with source correlation of the depth-1 closest launch, the depth-0 scatter
# launch and the depth-1 shadow launch of the bench workload.  usage: tools/profile_r1e.sh <tag>
TAG=${1:-r1e}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --spp 1 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:k_trace_closest -s 1 -c 1 -f -o gpurun_out/prof_closest_d1_$TAG $CMD > gpurun_out/prof_closest_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scatter -s 0 -c 1 -f -o gpurun_out/prof_scatter_d0_$TAG $CMD > gpurun_out/prof_scatter_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_shadow -s 1 -c 1 -f -o gpurun_out/prof_shadow_d1_$TAG $CMD > gpurun_out/prof_shadow_$TAG.log 2>&1
ls -la gpurun_out | head -30
