#!/bin/bash
# ncu --set full of the batch kernel: bash tools/gpu_ncu_batch.sh <tag> [opts]
TAG=${1:-r01j}; shift
mkdir -p gpurun_out
XMAPS_B200_OPTS="$1" timeout 600 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 1 -c 1 -o gpurun_out/prof_batch_$TAG -f python tools/profile_frames.py --frames 8 --reps 3 > gpurun_out/ncu_batch_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_batch_$TAG.log
