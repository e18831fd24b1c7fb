#!/bin/bash
# sampler parity tests + MERW leg of the bench with and without the L2-resident pair hash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== sampler tests"; timeout 400 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_launcher.py -q -x --timeout 120 > gpurun_out/pytest_sampler.log 2>&1; rc=$?; echo "rc=$rc"; tail -4 gpurun_out/pytest_sampler.log
if [ $rc -eq 124 ]; then echo "tests hung -- stopping"; exit 1; fi
for v in 0 1; do echo "== bench PN_MERW_PAIRS=$v"; PN_MERW_PAIRS=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_merw$v.json 2> gpurun_out/bench_merw$v.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_merw$v.json')); m=d['merw']; print('merw', m['value'], 'ms', m['ms_per_step'], m['roofline'])"; tail -2 gpurun_out/bench_merw$v.err; done
