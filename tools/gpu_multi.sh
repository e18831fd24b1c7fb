#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
echo "== bench --gpus $N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo "rc=$?"
cut -c1-900 gpurun_out/bench_g$N.json; tail -5 gpurun_out/bench_g$N.err
