#!/bin/bash
# HD 20 M-event frames (config 3) with option sets: bash tools/gpu_hd.sh <tag> "<k=v,k=v>" ...
TAG=$1; shift
mkdir -p gpurun_out
OUT=gpurun_out/hd_$TAG.txt
: > $OUT
for spec in "$@"; do
  opts=""
  if [ "$spec" != "-" ]; then for kv in $(echo $spec | tr ',' ' '); do opts="$opts --opt $kv"; done; fi
  r=$(timeout 300 python bench.py --workload hd20m --quick --check --steps 5 --warmup 3 $opts 2>> gpurun_out/hd_$TAG.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.2f us frac %.3f mism %s'%(d['frame_us'], d['roofline_frac'], d.get('mismatching_pixels')))")
  echo "== $spec: $r" | tee -a $OUT
done
python - <<'PY' | tee -a $OUT
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(1, os.path.join(os.getcwd(), "tests"))
import bench, torch
from xmaps_b200.engine import DepthEngine, TableSet
t = bench.load_geometry("hd20m")[0]
eng = DepthEngine(TableSet(t.lut_x, t.lut_y, t.x_map, t.remap_xy, t.rect_w, t.rect_h, t.t_px_scale, t.x_offset, t.depth_scale), device=torch.device("cuda", 0))
print("rect", t.rect_w, t.rect_h, "xmap", t.x_map.shape, "batch_cols", eng.get_option("batch_cols"), "batch_smem", eng.get_option("batch_smem"), "alive_px", eng.get_option("alive_px"))
PY
tail -3 gpurun_out/hd_$TAG.err
