#!/bin/bash
# A/B of an environment toggle: tensor-core parity tests, then two bench runs per setting
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VAR=${1:-PN_BWD_FORK}
echo "== tests"; timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_v2.py tests/test_gpu_train.py -q -x --timeout 90 > gpurun_out/pytest_q.log 2>&1; rc=$?; echo "rc=$rc"; tail -4 gpurun_out/pytest_q.log
if [ $rc -eq 124 ]; then echo "tests hung -- stopping"; exit 1; fi
for v in 0 1 0 1; do echo "== bench $VAR=$v"; env $VAR=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_ab$v.json 2> gpurun_out/bench_ab$v.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_ab$v.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"; tail -2 gpurun_out/bench_ab$v.err; done
