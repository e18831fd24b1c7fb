#!/bin/bash
# same-box A/B of an environment toggle on the split-precision step: VAR v1 v2 ... (two interleaved rounds)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
VAR=$1; shift
for round in 1 2; do for v in "$@"; do env $VAR=$v timeout 200 python bench.py --math tc3 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_ab3_$v.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_ab3_$v.json')); print('$VAR=$v tc3 ms/step', round(d['ms_per_step'],4))"; done; done
