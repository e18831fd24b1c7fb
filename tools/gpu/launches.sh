#!/bin/bash
# launch lists (ncu gpu__time_duration) of one bench step, once per environment setting given as arguments
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
i=0
for kv in "${@:-PN_X=1}"; do
  env $kv timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$i.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_$i.log 2>&1; echo "$kv rc=$?"
  i=$((i+1))
done
