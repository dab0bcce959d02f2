#!/bin/bash
# 8-GPU lines: direct peer-store gather (default), copy-engine gather with small chunks, no gather.  bash tools/gpu_n8.sh <tag> <N>
TAG=$1; N=${2:-8}
mkdir -p gpurun_out
run() {  # name, extra args
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/bench_n${N}_${name}_$TAG.json 2> gpurun_out/bench_n${N}_${name}_$TAG.err
  echo "== $name rc=$? $(python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_${name}_$TAG.json").read().strip().splitlines()[-1])
    print("value %.1f G ms/step %.3f parity %s e2e %.2f G" % (d["value"] / 1e9, d["ms_per_step"], d.get("parity", {}).get("mismatching_pixels"), d.get("e2e", {}).get("value", 0) / 1e9))
except Exception as e:
    print("no line:", e)
PY
)"
}
run direct
run copy16 --gather-mode copy --gather-chunk 16
run copy8 --gather-mode copy --gather-chunk 8
run nogather --no-gather
tail -2 gpurun_out/bench_n${N}_direct_$TAG.err
