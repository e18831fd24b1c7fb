#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest sampler + merw_prob"; timeout 600 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_merw_prob.py tests/test_gpu_launcher.py -m gpu -q -x --timeout 200 > gpurun_out/pytest_sampler.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_sampler.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print('value', d['value'], 'ms', d['ms_per_step']); print('merw', d['merw']['value'], d['merw']['ms_per_step'], d['merw']['roofline']['frac'])"
