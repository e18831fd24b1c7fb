#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_v2.py tests/test_gpu_train.py tests/test_gpu_pagg.py tests/test_gpu_variants.py -q --timeout 900 > gpurun_out/pytest_g.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/pytest_g.log | tail -8
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_g.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'kern', {k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()}, 'module', d.get('module_api',{}).get('ms_per_step'))"; tail -3 gpurun_out/bench_g.err
echo "== trace bwd"; PN_TRACE_BWD=700 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/trace_bwd.log; grep "trace bwd" gpurun_out/trace_bwd.log | tail -6
echo "== trace fwd"; PN_TRACE_FWD=700 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/trace_fwd.log; grep "trace fwd" gpurun_out/trace_fwd.log | tail -6
