#!/bin/bash
# Round 2, session A: full GPU test suite, alive A/B, default bench.
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_$TAG.txt
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log)
tail -25 gpurun_out/pytest_$TAG.log
for o in "alive=1" "alive=0"; do
  (timeout 300 python bench.py --quick --steps 10 --warmup 3 --opt $o >> gpurun_out/quick_$TAG.txt 2>> gpurun_out/quick_$TAG.err; echo "quick $o rc=$?")
done
cat gpurun_out/quick_$TAG.txt; tail -3 gpurun_out/quick_$TAG.err
(timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_$TAG.err)
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
