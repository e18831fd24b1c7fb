#!/bin/bash
# round 2, call A: new-surface tests first, then everything, then the bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
lscpu > gpurun_out/lscpu.txt 2>&1; nproc >> gpurun_out/lscpu.txt; free -g >> gpurun_out/lscpu.txt
echo "== new tests"; timeout 1500 python -m pytest tests/test_gpu_v2.py tests/test_gpu_train.py -q -x --timeout 900 -s > gpurun_out/pytest_new.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_new.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu rest"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_v2.py --deselect tests/test_gpu_train.py -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-900 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
