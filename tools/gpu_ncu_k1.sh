#!/bin/bash
# ncu --set full of K1 for several option sets: bash tools/gpu_ncu_k1.sh <tag> "<opts1>" "<opts2>" ...
TAG=$1; shift
i=0
for o in "$@"; do
  args=""
  for kv in $o; do args="$args --opt $kv"; done
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:'events_(ws_)?kernel' -s 4 -c 1 -o gpurun_out/prof_k1_${TAG}_$i -f python tools/profile_frames.py --frames 4 --reps 3 $args > gpurun_out/ncu_k1_${TAG}_$i.log 2>&1
  echo "$i: $o rc=$?"
  i=$((i+1))
done
