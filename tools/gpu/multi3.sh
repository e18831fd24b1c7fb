#!/bin/bash
# 2-GPU: the NCCL equivalence test in tensor-core mode (fork + in-graph all-reduce), then the strong-scaling bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest multi (homo fp32 + hetero tc)"; timeout 120 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 55 -k "homo or tc" > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_multi.log
echo "== bench --gpus 2 strong"; PN_BENCH_DEADLINE=60 timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --no-cpu-baseline --no-extras > gpurun_out/bench_g2_strong.json 2> gpurun_out/bench_g2_strong.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_g2_strong.json')); print('strong', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['ms_per_step'],4))" 2>&1 | tail -1
