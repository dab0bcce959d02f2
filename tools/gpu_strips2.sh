#!/bin/bash
# strip epilogue diagnostics: timelines + phase cycle counters (hooks builds): bash tools/gpu_strips2.sh <tag> <variant> "<events> [--opt k=v ...]" ...
TAG=$1; V=$2; shift; shift
mkdir -p gpurun_out
OUT=gpurun_out/strips2_$TAG.txt
: > $OUT
export XMAPS_B200_LIB=$PWD/build_variants/libxm_$V.so
for spec in "$@"; do
  set -- $spec
  echo "== $V: timeline events $spec" >> $OUT
  timeout 200 python tools/batch_timeline.py --events $1 ${@:2} --frames 24 2>&1 | tail -30 >> $OUT
done
cat $OUT
