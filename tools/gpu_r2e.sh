#!/bin/bash
# Round 2, session E (2 GPUs): gather modes next to the persistent kernel.
TAG=${1:-r2e}
mkdir -p gpurun_out
OUT=gpurun_out/multi_$TAG.txt
: > $OUT
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
N=${2:-2}
run() { echo "== N=$N $*" >> $OUT; (timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" >> $OUT 2>> gpurun_out/multi_$TAG.err); }
run --quick --no-gather
run --quick --check --gather-mode direct
run --quick --check --gather-mode copy
run --quick --check --gather-mode nccl
echo "== N=$N full line" >> $OUT
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 >> $OUT 2>> gpurun_out/multi_$TAG.err)
grep -v "^\s*$" $OUT | cut -c1-700; tail -5 gpurun_out/multi_$TAG.err
