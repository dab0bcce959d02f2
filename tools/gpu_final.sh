#!/bin/bash
# Round-end GPU session: smoke, all GPU tests, the bench (both arms), ncu launch list + full capture of the
# batch kernel, event-count sweep.  Usage: bash tools/gpu_final.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 240 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
(timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log); tail -3 gpurun_out/pytest_$TAG.log
(timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?")
cut -c1-600 gpurun_out/bench_$TAG.json
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?")
cut -c1-400 gpurun_out/bench_ref_$TAG.json
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'events_|epilogue_|bounds_|frame_|batch_' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --e2e-frames 2 --e2e-reps 1 --cpu-runs 1 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu-list rc=$?")
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 1 -c 1 -o gpurun_out/prof_batch_$TAG -f python tools/profile_frames.py --frames 8 --reps 3 > gpurun_out/ncu_batch_$TAG.log 2>&1; echo "ncu-batch rc=$?")
: > gpurun_out/sweep_events_$TAG.txt
for n in 1000000 2000000 5000000 10000000 20000000 50000000; do
  f=$(( 160000000 / n )); [ $f -gt 32 ] && f=32; [ $f -lt 4 ] && f=4
  (timeout 300 python bench.py --quick --steps 3 --frames $f --events $n 2>&1 | tail -1) >> gpurun_out/sweep_events_$TAG.txt
done
cut -c1-260 gpurun_out/sweep_events_$TAG.txt
