#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_v2.py tests/test_gpu_train.py tests/test_gpu_pagg.py tests/test_gpu_variants.py -q --timeout 900 -s > gpurun_out/pytest_e.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/pytest_e.log | tail -12
echo "== tests with PN_BWD_PF=1"; PN_BWD_PF=1 timeout 900 python -m pytest tests/test_gpu_tc.py -q --timeout 600 -k "oracle or fixture or pubmed" > gpurun_out/pytest_e_pf.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_e_pf.log
for v in 0 1; do echo "== bench PN_BWD_PF=$v"; PN_BWD_PF=$v timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e$v.json 2> gpurun_out/bench_e$v.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_e$v.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'kern', {k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()}, 'module', d.get('module_api',{}).get('ms_per_step'))"; tail -3 gpurun_out/bench_e$v.err; done
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu.log 2>&1; echo "rc=$?"
