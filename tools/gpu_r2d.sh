#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
OUT=gpurun_out/variants_$TAG.txt
: > $OUT
for v in fg2 hint nohint ps64 ps256; do
  export XMAPS_B200_LIB=$PWD/build_variants/libxm_$v.so
  echo "== $v" >> $OUT
  (timeout 300 python bench.py --quick --steps 10 --warmup 3 >> $OUT 2>> gpurun_out/variants_$TAG.err)
  (timeout 300 python bench.py --quick --steps 10 --warmup 3 --opt debug=16 >> $OUT 2>> gpurun_out/variants_$TAG.err)
done
export XMAPS_B200_LIB=$PWD/build_variants/libxm_ps64.so
(timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_plane.py -q -x --timeout 600 2>&1 | tail -3 >> $OUT)
unset XMAPS_B200_LIB
python - <<'PY'
import json
for ln in open('gpurun_out/variants_r2d.txt'):
    ln=ln.strip()
    if ln.startswith('=='): print(); print(ln, end='  ')
    elif ln.startswith('{'):
        d=json.loads(ln); print('%s frame_us %.2f' % (d['options'], d['frame_us']), end='   ')
    else: print(ln)
PY
tail -3 gpurun_out/variants_$TAG.err
