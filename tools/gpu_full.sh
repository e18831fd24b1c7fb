#!/bin/bash
# Full validation + evidence pass (1 GPU).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== pytest gpu all"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench (default = tc)"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json | cut -c1-1500; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-700 gpurun_out/bench_ref.json
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full lstm"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lstm_fwd_tc|k_lstm_bwd_tc|k_wgrad_tc' -s 9 -c 3 -o gpurun_out/prof_lstm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
echo "== ncu full merw"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_merw_walk' -s 4 -c 1 -o gpurun_out/prof_merw python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_merw.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
