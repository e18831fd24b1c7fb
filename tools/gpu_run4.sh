#!/bin/bash
# TC tests first (short timeout so a deadlock costs little), then everything else.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest tc"; timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout 120 -s > gpurun_out/pytest_tc.log 2>&1; rc=$?; echo "rc=$rc"
grep -E "tc |passed|failed|Error|error|assert" gpurun_out/pytest_tc.log | tail -30
if [ $rc -ne 0 ]; then echo "TC tests failed; stopping early"; tail -40 gpurun_out/pytest_tc.log; exit 0; fi
echo "== pytest gpu all"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"
tail -8 gpurun_out/pytest_gpu.log
echo "== bench tc"; timeout 600 python bench.py --steps 10 --warmup 3 --math tc > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "rc=$?"
cat gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
echo "== bench fp32"; timeout 600 python bench.py --steps 5 --warmup 3 --math fp32 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "rc=$?"
cat gpurun_out/bench_fp32.json
echo "== ncu launch list (tc)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 1 --warmup 3 --math tc --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
echo "== ncu full: lstm fwd/bwd tc, wgrad, merw"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_lstm_fwd_tc|k_lstm_bwd_tc|k_wgrad_tc|k_merw_walk' -s 6 -c 6 -o gpurun_out/prof_tc python bench.py --steps 1 --warmup 3 --math tc --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
