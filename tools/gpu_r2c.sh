#!/bin/bash
# round 2, call C: TMA GEMM path end to end, variants, A/B bench, ncu evidence, synthetic PAGG.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tma gemm unit tests"; timeout 300 python -m pytest tests/test_gpu_tc.py -q --timeout 120 -k "tma" > gpurun_out/pytest_tma.log 2>&1; rc=$?; echo "rc=$rc"; tail -6 gpurun_out/pytest_tma.log
if [ $rc -ne 0 ]; then export PN_TC_GEMM=legacy; echo "!! TMA GEMM failed: continuing with PN_TC_GEMM=legacy"; fi
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_v2.py tests/test_gpu_train.py tests/test_gpu_pagg.py tests/test_gpu_variants.py -q --timeout 900 -s > gpurun_out/pytest_c.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/pytest_c.log | tail -30
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench default (tma)"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'kern', {k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()}, 'module', d.get('module_api',{}).get('ms_per_step'))"; tail -3 gpurun_out/bench_c.err
echo "== bench legacy gemm"; PN_TC_GEMM=legacy timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_legacy.json 2> gpurun_out/bench_legacy.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_legacy.json')); print('legacy ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
echo "== workloads"; for w in cora bgp; do timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', d['value'], 'paths/s ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches/step', d['launches_per_step'])"; done
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full: gemm / pool / lstm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tma|k_gemm_tc|k_pool_fwd128|k_pool_bwd128|k_lstm_fwd_tc|k_lstm_bwd_tc|k_wgrad_tc|k_relu_mask_colsum' -s 60 -c 14 -o gpurun_out/prof_r2c python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_c.log 2>&1; echo "rc=$?"
echo "== synthetic PAGG (10M nodes)"; timeout 1500 python tools/bench_synthetic_pagg.py > gpurun_out/synth_pagg.json 2> gpurun_out/synth_pagg.err; echo "rc=$?"; cut -c1-1500 gpurun_out/synth_pagg.json; tail -5 gpurun_out/synth_pagg.err
ls -la gpurun_out/*.ncu-rep
