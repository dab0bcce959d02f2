#!/bin/bash
# Final numbers of the round (1 GPU): smoke, GPU tests, bench lines of every workload, launch list, ncu of the batch kernel.
TAG=${1:-r3r}
mkdir -p gpurun_out
timeout 240 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
(timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log); tail -2 gpurun_out/pytest_$TAG.log
(timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?")
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?")
for wl in 100k hd20m plane sweep; do
  (timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?")
done
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'events_|epilogue_|bounds_|frame_|batch_' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --e2e-frames 2 --e2e-reps 1 --cpu-runs 1 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu-list rc=$?")
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 1 -c 1 -o gpurun_out/prof_batch_$TAG -f python tools/profile_frames.py --frames 8 --reps 3 > gpurun_out/ncu_batch_$TAG.log 2>&1; echo "ncu-batch rc=$?")
python - <<PY
import json
for n in ["$TAG","100k_$TAG","hd20m_$TAG","plane_$TAG","sweep_$TAG","ref_$TAG"]:
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        r=d.get("roofline",{}); e=d.get("e2e",{}); c=d.get("cpu_baseline",{})
        print(n, "value %.1fG"%(d["value"]/1e9), "frac", round(r.get("frac",0),3), "frame_us", round(r.get("frame_us",0),2), "e2e %.2fG"%(e.get("value",0)/1e9), "cpu %.1fM"%(c.get("value",0)/1e6), "parity", (d.get("parity") or {}).get("mismatching_pixels"))
        if "sweep" in d: print([(r["events_per_frame"], round(r["value"]/1e9,1), round(r["frame_us"],1), round(r["roofline_frac"],3)) for r in d["sweep"]])
    except Exception as ex: print(n, "no line", ex)
PY
