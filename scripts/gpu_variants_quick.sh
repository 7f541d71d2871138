#!/bin/bash
# Runs on the GPU box: bench (4 steps) of several engine builds.  Variants are built here with
#   make -C arrow-h264_b200 OUT=variants/libx.so EXTRA="-DH264R_..." variants/libx.so
# Usage: scripts/gpu_variants_quick.sh <tag> <lib>...
tag=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  name=$(basename $lib .so)
  H264R_LIB=$PWD/$lib python bench.py --no-cpu-baseline --no-ceiling --steps 4 > gpurun_out/bench_${tag}_$name.json 2> gpurun_out/bench_${tag}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_$name.json").read())
    print("$name: ms/step %.2f  kernels %s parity %s" % (d["ms_per_step"], {k: round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["parity_checked"]))
except Exception as e:
    print("$name: FAILED", e, open("gpurun_out/bench_${tag}_$name.err").read()[-300:])
PY
done
