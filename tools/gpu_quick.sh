#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_v2.py tests/test_gpu_train.py -q --timeout 600 > gpurun_out/pytest_q.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_q.log
for i in 1 2; do echo "== bench $i"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_q$i.json 2> gpurun_out/bench_q$i.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_q$i.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'kern', {k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()}, 'module', d.get('module_api',{}).get('ms_per_step'))"; tail -2 gpurun_out/bench_q$i.err; done
