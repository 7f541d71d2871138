#!/bin/bash
# Runs on the GPU box (through gpurun): parity tests first, then the default bench.  Usage: scripts/gpu_check.sh <tag> [bench args]
tag=${1:-x}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_$tag.log 2>&1
rc=$?
tail -5 gpurun_out/pytest_$tag.log
if [ $rc -ne 0 ]; then echo "PYTEST FAILED rc=$rc"; grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/pytest_$tag.log | head -30; exit $rc; fi
python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
rc=$?
tail -12 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read())
e=d["e2e"]
print("value %.1f M MB/s  ms/step %.2f  kernels %s" % (d["value"]/1e6, d["ms_per_step"], {k: round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}))
print("e2e %.1f M MB/s  %.1f ms/step  fill %.1f ms (per thread %.1f)  flush %.1f ms  feeders %d  parity %s" % (e["value"]/1e6, e["ms_per_step"], e["host_fill_ms_per_step"], e["host_fill_ms_per_step_per_thread"], e["host_flush_ms_per_step"], e["feeder_threads"], d["parity"]))
print("ceiling", d.get("e2e_roofline"))
PY
exit $rc
