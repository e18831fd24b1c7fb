#!/bin/bash
# round 2, call D: backward prefetch + pooling changes: correctness, then bench + ncu of the changed kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_v2.py tests/test_gpu_train.py tests/test_gpu_pagg.py tests/test_gpu_variants.py -q --timeout 900 -s > gpurun_out/pytest_d.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/pytest_d.log | tail -25
echo "== bench default"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_d.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'kern', {k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()}, 'module', d.get('module_api',{}).get('ms_per_step'))"; tail -3 gpurun_out/bench_d.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full: lstm bwd / pool"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pool_fwd128|k_pool_bwd128|k_lstm_fwd_tc|k_lstm_bwd_tc' -s 16 -c 4 -o gpurun_out/prof_r2d python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_d.log 2>&1; echo "rc=$?"
echo "== synthetic MERW, 4 epochs per call"; timeout 900 python tools/bench_synthetic.py 3163 4 16 > gpurun_out/synth_merw_e4.json 2> gpurun_out/synth_merw_e4.err; echo "rc=$?"; cut -c1-700 gpurun_out/synth_merw_e4.json; tail -3 gpurun_out/synth_merw_e4.err
