#!/bin/bash
# Runs on the GPU box (through gpurun), after scripts/gpu_check.sh has passed without ncu:
#   1. launch list of the default bench command (per-launch gpu__time_duration, cold-cache and serialised)
#   2. one `ncu --set full` capture of the first flush of a 4-picture run (waves: I x64, P x64, B+B x128 = 8 launches: every
#      kernel of the path on every picture type), sources imported
# Usage: scripts/gpu_profile.sh <tag> [full-capture launch count] [H264R_LIB=...]
tag=${1:-x}; cnt=${2:-8}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --no-cpu-baseline --no-parity-check --no-ceiling --steps 2 --warmup 3 > gpurun_out/ncu_l$tag.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -c $cnt -f -o gpurun_out/prof_$tag \
    python bench.py --no-cpu-baseline --no-parity-check --no-ceiling --frames 4 --steps 1 > gpurun_out/ncu_$tag.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/prof_$tag.ncu-rep
