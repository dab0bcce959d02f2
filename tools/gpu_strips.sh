#!/bin/bash
# strip epilogue vs tile epilogue: bash tools/gpu_strips.sh <tag> [--test] <variant|-[:k=v,k=v]> ...
#   variant = build_variants/libxm_<variant>.so ("-" = the in-tree library), options after ':' go to --opt
TAG=$1; shift
mkdir -p gpurun_out
OUT=gpurun_out/strips_$TAG.txt
: > $OUT
if [ "$1" == "--test" ]; then
  shift
  (timeout 1200 python -m pytest tests/test_gpu_batch.py tests/test_gpu_plane.py -q -x --timeout 600 2>&1 | tail -5) | tee -a $OUT
fi
for spec in "$@"; do
  v=${spec%%:*}
  opts=""
  if [[ "$spec" == *:* ]]; then for kv in $(echo ${spec#*:} | tr ',' ' '); do opts="$opts --opt $kv"; done; fi
  if [ "$v" == "-" ]; then unset XMAPS_B200_LIB; else export XMAPS_B200_LIB=$PWD/build_variants/libxm_$v.so; fi
  line="== $spec:"
  for n in 5000000 2000000 1000000 100000; do
    r=$(timeout 300 python bench.py --quick --check --steps 10 --warmup 3 --events $n $opts 2>> gpurun_out/strips_$TAG.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.2f%s'%(d['frame_us'], '' if d.get('mismatching_pixels') == 0 else ' MISMATCH %s' % d.get('mismatching_pixels')))")
    line="$line  $n: $r"
  done
  echo "$line" | tee -a $OUT
done
unset XMAPS_B200_LIB
tail -5 gpurun_out/strips_$TAG.err
