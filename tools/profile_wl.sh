#!/bin/bash
# GPU box: one `ncu --set full` capture with source correlation of a kernel of a bench workload.
# usage: tools/profile_wl.sh <workload> <kernel regex> <skip> <count> <tag> [pass params JSON]
mkdir -p gpurun_out
P=${6:-"{\"bands\": 1}"}
ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/prof_$5 \
    python bench.py --workload $1 --steps 1 --warmup 1 --min-seconds 0 --no-cpu-baseline --params "$P" > gpurun_out/prof_$5.log 2>&1
ls -la gpurun_out/prof_$5.ncu-rep
