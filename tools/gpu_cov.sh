#!/bin/bash
# coverage-row tests: aggregator variants (incl. the GRU cell) and the sampler (incl. rng_mode = philox)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_variants.py tests/test_gpu_sampler.py -q -x --timeout 150 > gpurun_out/pytest_cov.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_cov.log
