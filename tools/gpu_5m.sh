#!/bin/bash
# 5 M-event frames only: bash tools/gpu_5m.sh <tag> <variant|-[:k=v,...]> ...
TAG=$1; shift
mkdir -p gpurun_out
OUT=gpurun_out/five_$TAG.txt
: > $OUT
for spec in "$@"; do
  v=${spec%%:*}
  opts=""
  if [[ "$spec" == *:* ]]; then for kv in $(echo ${spec#*:} | tr ',' ' '); do opts="$opts --opt $kv"; done; fi
  if [ "$v" == "-" ]; then unset XMAPS_B200_LIB; else export XMAPS_B200_LIB=$PWD/build_variants/libxm_$v.so; fi
  r=$(timeout 300 python bench.py --quick --check --steps 10 --warmup 3 $opts 2>> gpurun_out/five_$TAG.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.2f us frac %.3f mism %s'%(d['frame_us'], d['roofline_frac'], d.get('mismatching_pixels')))")
  echo "== $spec: $r" | tee -a $OUT
done
tail -3 gpurun_out/five_$TAG.err
