#!/bin/bash
# round 2, 2-GPU call: CUDA/NCCL equivalence test, weak and strong scaling lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/nvsmi2.txt
echo "== 2-GPU equivalence tests"; timeout 1200 python -m pytest tests/test_gpu_multi.py -q --timeout 900 -s > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_multi.log
run() { name=$1; shift; echo "== $name"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/$name.json')); print(d['config']['parallelism'], d['value'], 'paths/s ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d.get('module_api',{}).get('note','')[-60:])"; tail -2 gpurun_out/$name.err; }
run bench_2gpu_weak_pubmed
run bench_2gpu_strong_pubmed --scaling strong --no-extras
run bench_2gpu_strong_bgp --scaling strong --workload bgp --no-extras
run bench_2gpu_weak_bgp --workload bgp --no-extras
echo "== 1-GPU bgp + pubmed for the same box"; for w in pubmed bgp; do timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_1gpu_$w.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_1gpu_$w.json')); print('$w 1 GPU', d['value'], d['ms_per_step'])"; done
