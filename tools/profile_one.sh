#!/bin/bash
# GPU box: one `ncu --set full` capture with source correlation.
# usage: tools/profile_one.sh <kernel regex> <skip> <tag> [pass params JSON]
# The captures that feed profiles/traffic.json run the frame as ONE band ('{"bands": 1}'): bench.py's
# roofline.achieved is per launch of the one-band profiling pass, and traffic must be per the same launch.
mkdir -p gpurun_out
P=${4:-"{\"bands\": 1}"}
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/prof_$3 python bench.py --steps 1 --warmup 1 --spp 1 --no-cpu-baseline --params "$P" > gpurun_out/prof_$3.log 2>&1
ls -la gpurun_out/prof_$3.ncu-rep
