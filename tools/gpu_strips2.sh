#!/bin/bash
# strip epilogue diagnostics: timelines + phase cycle counters (hooks builds): bash tools/gpu_strips2.sh <tag> <variant>...
TAG=$1; shift
mkdir -p gpurun_out
OUT=gpurun_out/strips2_$TAG.txt
: > $OUT
for v in "$@"; do
  export XMAPS_B200_LIB=$PWD/build_variants/libxm_$v.so
  for spec in "1000000 " "1000000 --opt debug=16" "5000000 "; do
    set -- $spec
    echo "== $v: timeline events $spec" >> $OUT
    timeout 200 python tools/batch_timeline.py --events $1 ${@:2} --frames 16 2>&1 | tail -4 >> $OUT
  done
done
cat $OUT
