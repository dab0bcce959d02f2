#!/bin/bash
# GPU parity tests + quick benches for a list of option sets: bash tools/gpu_quick.sh <tag> "<opts1>" ...
TAG=$1; shift
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -5) | cut -c1-200
: > gpurun_out/sweep_$TAG.txt
for o in "$@"; do
  args=""
  for kv in $o; do args="$args --opt $kv"; done
  (timeout 200 python bench.py --quick --steps 5 --frames 32 $args 2>&1 | tail -1) >> gpurun_out/sweep_$TAG.txt
done
cut -c1-330 gpurun_out/sweep_$TAG.txt
