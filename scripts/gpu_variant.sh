#!/bin/bash
# Runs on the GPU box: parity tests and the bench against ANOTHER build of the engine (arrow-h264_b200/variants/*.so, built
# with `make OUT=variants/libx.so EXTRA=-D... variants/libx.so`).  Usage: scripts/gpu_variant.sh <tag> <lib> [bench args]
tag=$1; lib=$2; shift 2
mkdir -p gpurun_out
export H264R_LIB=$PWD/$lib
python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 > gpurun_out/pytest_$tag.log 2>&1
rc=$?
tail -3 gpurun_out/pytest_$tag.log
if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/pytest_$tag.log | head -20; fi
python bench.py --no-cpu-baseline --no-ceiling "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read())
print("$tag: value %.1f M MB/s  ms/step %.2f  kernels %s parity %s" % (d["value"]/1e6, d["ms_per_step"], {k: round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["parity_checked"]))
PY
