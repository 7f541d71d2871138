#!/bin/bash
# Runs on the GPU box (through gpurun): parity tests first, then the default bench.  Usage: scripts/gpu_check.sh <tag> [bench args]
tag=${1:-x}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1
rc=$?
tail -5 gpurun_out/pytest_$tag.log
if [ $rc -ne 0 ]; then echo "PYTEST FAILED rc=$rc"; exit $rc; fi
python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
rc=$?
tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read())
print("value %.1f M MB/s  ms/step %.2f  e2e %.1f M  kernels %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, {k: round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}))
PY
exit $rc
