#!/bin/bash
# parameter sweep of bench.py --quick; usage: bash tools/gpu_sweep.sh <tag> "<opts1>" "<opts2>" ...
TAG=$1; shift
mkdir -p gpurun_out
: > gpurun_out/sweep_$TAG.txt
for o in "$@"; do
  args=""
  for kv in $o; do args="$args --opt $kv"; done
  (timeout 200 python bench.py --quick --steps 5 --frames 16 $args 2>&1 | tail -1) >> gpurun_out/sweep_$TAG.txt
done
cat gpurun_out/sweep_$TAG.txt | cut -c1-400
