#!/bin/bash
# bench only, once per environment setting
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for kv in "$@"; do
  echo "== bench $kv"; env $kv timeout 300 python bench.py --steps 10 --warmup 3 --math tc --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step']); print({k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()})"
done
