#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest tc"; timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout 120 -s -k "not gemm" > gpurun_out/pytest_tc.log 2>&1; rc=$?; echo "rc=$rc"
grep -E "tc |passed|failed|Error|error|assert" gpurun_out/pytest_tc.log | tail -20
if [ $rc -ne 0 ]; then tail -40 gpurun_out/pytest_tc.log; exit 0; fi
echo "== bench tc"; timeout 600 python bench.py --steps 10 --warmup 3 --math tc --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step']); print(d['roofline']['kernels']); print('merw', d['merw']['value'], d['merw']['ms_per_step'])"
tail -3 gpurun_out/bench_tc.err
