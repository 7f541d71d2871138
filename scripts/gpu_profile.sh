#!/bin/bash
# Runs on the GPU box (through gpurun), after scripts/gpu_check.sh has passed without ncu:
#   1. launch list of the default bench command (per-launch gpu__time_duration, cold-cache and serialised)
#   2. one `ncu --set full` capture of one wave set (I wave, P wave, B+B wave: every kernel of the path), sources imported
# Usage: scripts/gpu_profile.sh <tag> [full-capture launch count]
tag=${1:-x}; cnt=${2:-16}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_l$tag.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -c $cnt -f -o gpurun_out/prof_$tag \
    python bench.py --no-cpu-baseline --frames 4 --steps 1 > gpurun_out/ncu_$tag.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/prof_$tag.ncu-rep
