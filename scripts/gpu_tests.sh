#!/bin/bash
# Runs on the GPU box (through gpurun): smoke, then the GPU parity tests (compute-sanitizer is closed on this pool).  Usage: scripts/gpu_tests.sh <tag> [pytest args]
tag=${1:-x}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > gpurun_out/gpu_$tag.txt
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$tag.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 "$@" > gpurun_out/pytest_$tag.log 2>&1
rc=$?
tail -25 gpurun_out/pytest_$tag.log
exit $rc
