#!/bin/bash
# wgrad cluster-multicast check: TC tests first (short timeouts: a cluster deadlock must not eat the budget), then bench + ncu traffic.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest tc"; PN_TRACE_BWD=1 timeout 240 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout 60 > gpurun_out/pytest_tc.log 2>&1; rc=$?; echo "rc=$rc"; grep -m1 "max active" gpurun_out/pytest_tc.log; tail -3 gpurun_out/pytest_tc.log
[ $rc -ne 0 ] && { grep -E "Error|error|FAILED" gpurun_out/pytest_tc.log | head; exit 1; }
echo "== bench tc"; timeout 300 python bench.py --steps 10 --warmup 3 --math tc --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step']); print(d['roofline']['kernels'])"
echo "== ncu wgrad"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_wgrad_tc' -s 3 -c 1 -o gpurun_out/prof_wgrad python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_wgrad.log 2>&1; echo "rc=$?"
