#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in cora bgp; do
  echo "== bench --workload $w"; timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value']); print(d['roofline']['kernels']); print('merw', d['merw']['value'], d['merw']['ms_per_step']); print('cpu', d.get('cpu_baseline',{}).get('value'), d['merw'].get('cpu_baseline',{}).get('value'))"
  tail -2 gpurun_out/bench_$w.err
done
