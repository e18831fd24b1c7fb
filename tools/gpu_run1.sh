#!/bin/bash
# First GPU pass: smoke, parity tests, bench, memcheck of smoke, ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt; nproc >> gpurun_out/lscpu.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== memcheck smoke"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/memcheck.log
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/smoke.log
