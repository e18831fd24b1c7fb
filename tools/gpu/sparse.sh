#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest sampler (incl. sparse path)"; timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -q -x --timeout 200 > gpurun_out/pytest_sampler.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_sampler.log
