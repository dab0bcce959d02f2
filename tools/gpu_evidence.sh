#!/bin/bash
# Round evidence session (1 GPU): bench lines of every workload, ncu launch list + full capture of the batch
# kernel, frame_kernel and the stand-alone epilogue, compute-sanitizer on a small batch.  bash tools/gpu_evidence.sh <tag>
TAG=${1:-r2n}
mkdir -p gpurun_out
for wl in 100k hd20m plane sweep; do
  (timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?")
  cut -c1-300 gpurun_out/bench_${wl}_$TAG.json
done
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'events_|epilogue_|bounds_|frame_|batch_' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --e2e-frames 2 --e2e-reps 1 --cpu-runs 1 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu-list rc=$?")
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 1 -c 1 -o gpurun_out/prof_batch_$TAG -f python tools/profile_frames.py --frames 8 --reps 3 > gpurun_out/ncu_batch_$TAG.log 2>&1; echo "ncu-batch rc=$?")
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 2 -c 1 -o gpurun_out/prof_frame_$TAG -f python tools/profile_frames.py --frames 4 --reps 2 --single > gpurun_out/ncu_frame_$TAG.log 2>&1; echo "ncu-frame rc=$?")
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:epilogue_projector7 -s 2 -c 1 -o gpurun_out/prof_epi_$TAG -f python tools/profile_frames.py --frames 4 --reps 2 --single --opt fused=0 > gpurun_out/ncu_epi_$TAG.log 2>&1; echo "ncu-epi rc=$?")
(timeout 900 compute-sanitizer --tool memcheck python tools/profile_frames.py --frames 3 --reps 1 --events 300000 > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_$TAG.log)
(timeout 900 compute-sanitizer --tool racecheck python tools/profile_frames.py --frames 3 --reps 1 --events 300000 > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_$TAG.log)
ls -la gpurun_out | grep $TAG
