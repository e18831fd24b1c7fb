#!/bin/bash
# 2-GPU bench with a deadline guard: if a rank hangs, every thread's Python stack is dumped instead of waiting for the outer limit
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for sc in "$@"; do
echo "== bench --gpus 2 $sc"; PN_BENCH_DEADLINE=80 NCCL_DEBUG=WARN timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 2 --steps 10 --warmup 3 --scaling $sc --no-cpu-baseline --no-extras > gpurun_out/bench_g2_$sc.json 2> gpurun_out/bench_g2_$sc.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_g2_$sc.json')); print('$sc', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['ms_per_step'],4))" 2>&1 | tail -1
grep -E "File \"/root/repo|Thread|Timeout|shutdown|NCCL WARN" gpurun_out/bench_g2_$sc.err | head -30
done
