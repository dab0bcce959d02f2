#!/bin/bash
# A/B of prebuilt library variants: bash tools/gpu_ab.sh <tag> [--test] <variant|-> ...
#   each variant = build_variants/libxm_<variant>.so ("-" = the in-tree library); appends ":k=v,k=v" for --opt settings
TAG=$1; shift
TEST=0
if [ "$1" == "--test" ]; then TEST=1; shift; fi
mkdir -p gpurun_out
OUT=gpurun_out/ab_$TAG.txt
: > $OUT
for spec in "$@"; do
  v=${spec%%:*}
  opts=""
  if [[ "$spec" == *:* ]]; then for kv in $(echo ${spec#*:} | tr ',' ' '); do opts="$opts --opt $kv"; done; fi
  if [ "$v" == "-" ]; then unset XMAPS_B200_LIB; else export XMAPS_B200_LIB=$PWD/build_variants/libxm_$v.so; fi
  echo "== $spec" >> $OUT
  if [ $TEST == 1 ] && [ -z "$opts" ]; then
    (timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_plane.py tests/test_gpu_parity.py -q -x --timeout 600 2>&1 | tail -3 >> $OUT)
  fi
  (timeout 300 python bench.py --quick --check --steps 10 --warmup 3 $opts >> $OUT 2>> gpurun_out/ab_$TAG.err)
done
unset XMAPS_B200_LIB
python - "$OUT" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    ln = ln.strip()
    if ln.startswith('=='): print(); print(ln, end='  ')
    elif ln.startswith('{'):
        d = json.loads(ln); print('frame_us %.2f frac %.3f mism %s' % (d['frame_us'], d['roofline_frac'], d.get('mismatching_pixels')), end='   ')
    else: print(ln, end=' ')
print()
PY
tail -5 gpurun_out/ab_$TAG.err
