#!/bin/bash
# Runs on the GPU box: (optionally) the parity tests, then the bench under several stream topologies.
# The environment variables below are read by the stream-groups build only (profiles/r2_stream_groups_variant.diff applied
# to arrow-h264_b200/csrc); the product build ignores them.
# Usage: scripts/gpu_groups.sh <tag> <tests: 0|1> <G:M:K>...     G = stream groups, M = inter stream mode, K = wavefront CTAs per SM
tag=$1; shift
tests=$1; shift
mkdir -p gpurun_out
rc=0
if [ "$tests" = 1 ]; then
  python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_$tag.log 2>&1
  rc=$?
  tail -3 gpurun_out/pytest_$tag.log
  if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/pytest_$tag.log | head -30; fi
fi
for cfg in "$@"; do
  IFS=: read g m k <<< "$cfg"
  H264R_STREAM_GROUPS=$g H264R_INTER_STREAM_MODE=$m H264R_WAVEFRONT_CTAS_PER_SM=$k timeout 600 python bench.py --no-cpu-baseline --no-ceiling --steps 4 > gpurun_out/bench_${tag}_$g$m$k.json 2> gpurun_out/bench_${tag}_$g$m$k.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_$g$m$k.json").read())
    print("G=$g M=$m K=$k: value %.1f M  ms/step %.2f  e2e %.1f M  serial kernels %s parity %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, {k: round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["parity_checked"]))
except Exception as e:
    print("G=$g M=$m K=$k: FAILED", e, open("gpurun_out/bench_${tag}_$g$m$k.err").read()[-400:])
PY
done
exit $rc
