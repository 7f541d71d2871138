#!/bin/bash
# Kernel-tuning helper.  Here (CPU box):   scripts/tune.sh build "<name> <nvcc -D flags>" ...   builds library variants
# into arrow-h264_b200/variants/.  On the GPU box: scripts/tune.sh run [bench args]   benches every variant.
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
    shift; mkdir -p arrow-h264_b200/variants; rm -f arrow-h264_b200/variants/*.so
    for spec in "$@"; do
        set -- $spec; name=$1; shift
        (cd arrow-h264_b200 && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
            -I../include -Icsrc "$@" -shared -o variants/$name.so csrc/engine.cu csrc/kernels.cu csrc/host_helpers.cc csrc/decoder_facade.cc -lcudart) || exit 1
    done
    ls -la arrow-h264_b200/variants
else
    shift
    for so in arrow-h264_b200/variants/*.so; do
        H264R_LIB=$PWD/$so python bench.py --no-cpu-baseline --steps 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$(basename $so)', 'value %.1f M  ms %.2f  e2e %.1f M ' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6), {k: round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()})"
    done
fi
