#!/bin/bash
# sweep one environment variable over the given values: one short bench run each
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
VAR=$1; shift
for v in "$@"; do echo "== bench $VAR=$v"; env $VAR=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_sw_$v.json 2> gpurun_out/bench_sw_$v.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_sw_$v.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {k:round(x['ms'],4) for k,x in d['roofline']['kernels'].items()})"; tail -2 gpurun_out/bench_sw_$v.err; done
