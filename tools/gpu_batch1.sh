#!/bin/bash
# batch kernel bring-up: guarded tests, then quick benches with and without the batch kernel
TAG=${1:-r01h}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch.py -q --timeout 120 > gpurun_out/pytest_batch_$TAG.log 2>&1
echo "batch pytest rc=$?"; tail -25 gpurun_out/pytest_batch_$TAG.log | cut -c1-200
: > gpurun_out/quick_$TAG.txt
for o in "" "batch=0" "debug=16" "ctas_per_sm=2" "stages=1" "smem_cols_bytes=5400"; do
  (XMAPS_B200_OPTS="$o" timeout 200 python bench.py --quick --steps 5 --frames 32 2>&1 | tail -1) >> gpurun_out/quick_$TAG.txt
done
cut -c1-300 gpurun_out/quick_$TAG.txt
