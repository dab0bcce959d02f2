#!/bin/bash
# First GPU session after the fused frame kernel: guarded smoke (the kernel has a grid-wide barrier),
# then the usual round (tests, bench, launch list, ncu captures).  Usage: bash tools/gpu_session1.sh <tag>
TAG=${1:-r01g}
mkdir -p gpurun_out
timeout 240 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1
rc=$?
tail -5 gpurun_out/smoke_$TAG.log
if [ $rc -ne 0 ]; then
  echo "fused smoke failed rc=$rc -> continuing with fused=0"
  export XMAPS_B200_OPTS="fused=0"
fi
(timeout 200 python bench.py --quick --steps 5 --frames 16 2>&1 | tail -1) > gpurun_out/quick_$TAG.txt
(XMAPS_B200_OPTS="fused=0" timeout 200 python bench.py --quick --steps 5 --frames 16 2>&1 | tail -1) >> gpurun_out/quick_$TAG.txt
cat gpurun_out/quick_$TAG.txt | cut -c1-300
(timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log)
tail -8 gpurun_out/pytest_$TAG.log
(timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_$TAG.err)
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'events_|epilogue_|bounds_|frame_|batch_' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --frames 8 --e2e-frames 2 --e2e-reps 1 --cpu-runs 1 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu-list rc=$?")
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:'batch_kernel|frame_kernel|events_' -s 1 -c 1 -o gpurun_out/prof_k1_$TAG -f python tools/profile_frames.py --frames 4 --reps 3 > gpurun_out/ncu_k1_$TAG.log 2>&1; echo "ncu-k1 rc=$?")
ls -la gpurun_out
