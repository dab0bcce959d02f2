#!/bin/bash
# batch2 bring-up with tight time limits (a deadlock must not eat GPU minutes)
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_batch.py -q -x --timeout 40 -k plain > gpurun_out/pytest_batch2_$TAG.log 2>&1
echo "batch2 pytest rc=$?"; tail -4 gpurun_out/pytest_batch2_$TAG.log | cut -c1-200
: > gpurun_out/quick2_$TAG.txt
for o in "$@"; do
  (XMAPS_B200_OPTS="$o" timeout 100 python bench.py --quick --steps 5 --frames 32 2>&1 | tail -1) >> gpurun_out/quick2_$TAG.txt
done
cut -c1-260 gpurun_out/quick2_$TAG.txt
