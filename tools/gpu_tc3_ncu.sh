#!/bin/bash
# ncu --set full of the split-precision LSTM kernels (third step of a pubmed-shaped run), then a short bench of the mode
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lstm_fwd_tc3|k_lstm_bwd_tc3' -s 4 -c 2 -f -o gpurun_out/prof_tc3 python bench.py --math tc3 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_tc3.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --math tc3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_tc3.json 2> gpurun_out/bench_tc3.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tc3.json')); print('tc3 ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"; tail -3 gpurun_out/bench_tc3.err
