#!/bin/bash
# bench lines of every workload (1 GPU): bash tools/gpu_lines.sh <tag>
TAG=$1
mkdir -p gpurun_out
(timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?")
for wl in 100k hd20m plane sweep; do
  (timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?")
done
python - <<PY
import json
for n in ["$TAG","100k_$TAG","hd20m_$TAG","plane_$TAG","sweep_$TAG"]:
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        r=d.get("roofline",{}); e=d.get("e2e",{}); c=d.get("cpu_baseline",{})
        print(n, "value %.1fG"%(d["value"]/1e9), "frac", round(r.get("frac",0),3), "frame_us", round(r.get("frame_us",0),2), "e2e %.2fG"%(e.get("value",0)/1e9), "cpu %.1fM"%(c.get("value",0)/1e6), "parity", (d.get("parity") or {}).get("mismatching_pixels"))
        if "sweep" in d: print([(r["events_per_frame"], round(r["value"]/1e9,1), round(r["frame_us"],1), round(r["roofline_frac"],3)) for r in d["sweep"]])
    except Exception as ex: print(n, "no line", ex)
PY
