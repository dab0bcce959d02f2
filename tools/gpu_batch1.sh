#!/bin/bash
# batch kernel: guarded tests, then quick benches of option sets given as arguments
TAG=${1:-r01h}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch.py -q -x --timeout 120 > gpurun_out/pytest_batch_$TAG.log 2>&1
echo "batch pytest rc=$?"; tail -25 gpurun_out/pytest_batch_$TAG.log | cut -c1-200
: > gpurun_out/quick_$TAG.txt
for o in "$@"; do
  (XMAPS_B200_OPTS="$o" timeout 200 python bench.py --quick --steps 5 --frames 32 2>&1 | tail -1) >> gpurun_out/quick_$TAG.txt
done
cut -c1-300 gpurun_out/quick_$TAG.txt
