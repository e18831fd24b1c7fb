#!/bin/bash
# same-box A/Bs: tc3 backward prefetch variants, then the tc backward with / without its L2 prefetch
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for round in 1 2; do for v in 0 2; do PN_BWD3_PF=$v timeout 200 python bench.py --math tc3 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_ab3_$v.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_ab3_$v.json')); print('PN_BWD3_PF=$v tc3 ms/step', round(d['ms_per_step'],4))"; done; done
for round in 1 2; do for v in 1 0; do PN_BWD_L2PF=$v timeout 200 python bench.py --math tc --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ab4_$v.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_ab4_$v.json')); print('PN_BWD_L2PF=$v tc ms/step', round(d['ms_per_step'],4), {k:round(x['ms'],4) for k,x in d['roofline']['kernels'].items()})"; done; done
