#!/bin/bash
# Final verification of the committed state (1 GPU): smoke, all GPU tests, the bench line, compute-sanitizer on both
# batch kernel instantiations (two and four epilogue warps).  bash tools/gpu_verify.sh <tag>
TAG=${1:-r3n}
mkdir -p gpurun_out
timeout 240 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log
(timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log); tail -3 gpurun_out/pytest_$TAG.log
(timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?")
python -c "
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('value %.1f G, %.2f us/frame, frac %.3f, e2e %.2f G, parity %s, launches %s' % (d['value']/1e9, d['roofline']['frame_us'], d['roofline']['frac'], d['e2e']['value']/1e9, d['parity']['mismatching_pixels'], d['gpu_launches']))"
for tw in 2 4; do
  (timeout 900 compute-sanitizer --tool memcheck python tools/profile_frames.py --frames 6 --reps 1 --events 300000 --opt tile_warps=$tw > gpurun_out/sanitizer_memcheck_tw${tw}_$TAG.log 2>&1; echo "memcheck tw=$tw rc=$?"; tail -1 gpurun_out/sanitizer_memcheck_tw${tw}_$TAG.log)
done
(timeout 900 compute-sanitizer --tool racecheck python tools/profile_frames.py --frames 6 --reps 1 --events 300000 --opt tile_warps=2 > gpurun_out/sanitizer_racecheck_tw2_$TAG.log 2>&1; echo "racecheck tw=2 rc=$?"; tail -1 gpurun_out/sanitizer_racecheck_tw2_$TAG.log)
(timeout 900 compute-sanitizer --tool synccheck python tools/profile_frames.py --frames 6 --reps 1 --events 300000 > gpurun_out/sanitizer_synccheck_$TAG.log 2>&1; echo "synccheck rc=$?"; tail -1 gpurun_out/sanitizer_synccheck_$TAG.log)
