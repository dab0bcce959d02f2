#!/bin/bash
# small-frame floor probes: bash tools/gpu_small.sh "<variant>|<events>|<opt>" ...
for spec in "$@"; do
  v=${spec%%|*}; rest=${spec#*|}; n=${rest%%|*}; o=${rest#*|}
  if [ "$v" == "-" ]; then unset XMAPS_B200_LIB; else export XMAPS_B200_LIB=$PWD/build_variants/libxm_$v.so; fi
  opts=""; [ -n "$o" ] && opts="--opt $o"
  echo "== $spec $(python bench.py --quick --check --steps 5 --events $n $opts | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('frame_us %.2f mism %s'%(d['frame_us'], d.get('mismatching_pixels')))")"
done
