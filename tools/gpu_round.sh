#!/bin/bash
# One GPU session: parity tests, bench, ncu launch list, ncu full capture of K1 and K2.
# Usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt
(timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log)
tail -15 gpurun_out/pytest_$TAG.log
(timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_$TAG.err)
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'events_|epilogue_|bounds_|frame_|batch_' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --frames 8 --e2e-frames 2 --e2e-reps 1 --cpu-runs 1 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu-list rc=$?")
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:events_ -s 6 -c 2 -o gpurun_out/prof_k1_$TAG -f python tools/profile_frames.py --frames 4 --reps 3 > gpurun_out/ncu_k1_$TAG.log 2>&1; echo "ncu-k1 rc=$?")
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:epilogue_projector -s 6 -c 2 -o gpurun_out/prof_k2_$TAG -f python tools/profile_frames.py --frames 4 --reps 3 > gpurun_out/ncu_k2_$TAG.log 2>&1; echo "ncu-k2 rc=$?")
ls -la gpurun_out
