#!/bin/bash
# round 2, final validation + evidence pass (1 GPU): smoke, every GPU test, both bench arms, the other workloads, launch lists and
# ncu --set full captures of the tensor-core kernels in both tensor-core math modes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== pytest gpu all"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/pytest_gpu_final.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/pytest_gpu_final.log | tail -4
echo "== bench (default)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_ref_final.json
for wl in cora bgp; do echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${wl}_final.json 2> gpurun_out/bench_${wl}_final.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_${wl}_final.json')); print('$wl ms/step', d['ms_per_step'], 'value', d['value'], 'tc3', d.get('strict_tc3',{}).get('ms_per_step'), 'fp32', d.get('strict',{}).get('ms_per_step'))"; done
echo "== ncu launch list tc"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu.log 2>&1; echo "rc=$?"
echo "== ncu launch list tc3"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_tc3_final.csv python bench.py --math tc3 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu3.log 2>&1; echo "rc=$?"
echo "== ncu full tc"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tma|k_gemm_tc|k_pool_fwd128|k_pool_bwd128|k_lstm_fwd_tc|k_lstm_bwd_tc|k_wgrad_tc|k_relu_mask_colsum' -s 52 -c 13 -f -o gpurun_out/prof_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_final.log 2>&1; echo "rc=$?"
echo "== ncu full tc3"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lstm_fwd_tc3|k_lstm_bwd_tc3' -s 4 -c 2 -f -o gpurun_out/prof_tc3_final python bench.py --math tc3 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_tc3_final.log 2>&1; echo "rc=$?"
echo "== ncu full merw"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_merw_walk' -s 4 -c 1 -f -o gpurun_out/prof_merw_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_merw_final.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*final*.ncu-rep
