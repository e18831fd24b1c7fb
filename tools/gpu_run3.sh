#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest tc"; timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout 300 -s > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"
grep -E "tc |passed|failed|Error|error" gpurun_out/pytest_tc.log | tail -40
echo "== pytest gpu all"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"
tail -15 gpurun_out/pytest_gpu.log
echo "== bench tc"; timeout 600 python bench.py --steps 5 --warmup 3 --math tc --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "rc=$?"
cat gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
