#!/bin/bash
# Round-2 final evidence session after the strip epilogue (1 GPU).  bash tools/gpu_round3.sh <tag>
TAG=${1:-r3f}
mkdir -p gpurun_out
timeout 240 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log
(timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log); tail -3 gpurun_out/pytest_$TAG.log
(timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?")
cut -c1-500 gpurun_out/bench_$TAG.json
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?")
cut -c1-300 gpurun_out/bench_ref_$TAG.json
for wl in 100k hd20m plane sweep; do
  (timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?")
  cut -c1-300 gpurun_out/bench_${wl}_$TAG.json
done
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'events_|epilogue_|bounds_|frame_|batch_' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --e2e-frames 2 --e2e-reps 1 --cpu-runs 1 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu-list rc=$?")
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 1 -c 1 -o gpurun_out/prof_batch_$TAG -f python tools/profile_frames.py --frames 8 --reps 3 > gpurun_out/ncu_batch_$TAG.log 2>&1; echo "ncu-batch rc=$?")
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 1 -c 1 -o gpurun_out/prof_batch1m_$TAG -f python tools/profile_frames.py --frames 16 --reps 3 --events 1000000 > gpurun_out/ncu_batch1m_$TAG.log 2>&1; echo "ncu-batch-1m rc=$?")
(timeout 900 compute-sanitizer --tool memcheck python tools/profile_frames.py --frames 6 --reps 1 --events 300000 > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_$TAG.log)
(timeout 900 compute-sanitizer --tool racecheck python tools/profile_frames.py --frames 6 --reps 1 --events 300000 > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_$TAG.log)
ls gpurun_out | grep $TAG
