#!/bin/bash
# Round 2, session B: kernel variants A/B (front group, 3 CTAs/SM), events-only floors, ring depth.
TAG=${1:-r2b}
mkdir -p gpurun_out
OUT=gpurun_out/variants_$TAG.txt
: > $OUT
for v in fg2 fg4 c3; do
  export XMAPS_B200_LIB=$PWD/build_variants/libxm_$v.so
  echo "== $v tests" >> $OUT
  (timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_plane.py -q -x --timeout 600 2>&1 | tail -3 >> $OUT)
  echo "== $v bench" >> $OUT
  (timeout 300 python bench.py --quick --check --steps 10 --warmup 3 >> $OUT 2>> gpurun_out/variants_$TAG.err)
  echo "== $v events only (debug=16)" >> $OUT
  (timeout 300 python bench.py --quick --steps 10 --warmup 3 --opt debug=16 >> $OUT 2>> gpurun_out/variants_$TAG.err)
done
export XMAPS_B200_LIB=$PWD/build_variants/libxm_fg4.so
echo "== fg4 stages=3 cols=2" >> $OUT
(timeout 300 python bench.py --quick --steps 10 --warmup 3 --opt smem_cols_bytes=5280 --opt stages=3 >> $OUT 2>> gpurun_out/variants_$TAG.err)
echo "== fg4 cols=2" >> $OUT
(timeout 300 python bench.py --quick --steps 10 --warmup 3 --opt smem_cols_bytes=5280 >> $OUT 2>> gpurun_out/variants_$TAG.err)
export XMAPS_B200_LIB=$PWD/build_variants/libxm_c3.so
echo "== c3 sweep 1M/2M" >> $OUT
(timeout 300 python bench.py --quick --steps 5 --warmup 3 --events 1000000 >> $OUT 2>> gpurun_out/variants_$TAG.err)
(timeout 300 python bench.py --quick --steps 5 --warmup 3 --events 2000000 >> $OUT 2>> gpurun_out/variants_$TAG.err)
unset XMAPS_B200_LIB
cat $OUT; tail -5 gpurun_out/variants_$TAG.err
