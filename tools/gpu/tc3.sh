#!/bin/bash
# split-precision mode: its parity tests (verbose: the observed errors are printed), the other tensor-core suites, then a short bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== tc3 tests"; timeout 600 python -m pytest tests/test_gpu_tc3.py -q -s --timeout 240 > gpurun_out/pytest_tc3.log 2>&1; rc=$?; echo "rc=$rc"
grep -E "^\[tc3|passed|failed|Mismatched|Greatest|AssertionError|tc3 vs" gpurun_out/pytest_tc3.log | head -80
if [ $rc -eq 124 ]; then echo "tests hung -- stopping"; exit 1; fi
echo "== bench tc3"; timeout 300 python bench.py --math tc3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_tc3.json 2> gpurun_out/bench_tc3.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tc3.json')); print('tc3 ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"; tail -3 gpurun_out/bench_tc3.err
if [ "$1" = "all" ]; then
echo "== other suites"; timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_v2.py tests/test_gpu_train.py -q -x --timeout 300 > gpurun_out/pytest_q.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_q.log
fi
if [ "$1" = "sweep" ]; then
for v in 33 30 24; do echo "== bench tc3 PN_WGRAD_CLUSTERS=$v"; PN_WGRAD_CLUSTERS=$v timeout 200 python bench.py --math tc3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_tc3_$v.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_tc3_$v.json')); print('tc3 ms/step', d['ms_per_step'])"; done
fi
