#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
free -g | head -2
echo "== synthetic 1M"; timeout 600 python tools/bench_synthetic.py 1000 2 32 > gpurun_out/synth_1m.json 2> gpurun_out/synth_1m.err; echo "rc=$?"; cat gpurun_out/synth_1m.json; tail -3 gpurun_out/synth_1m.err
echo "== synthetic 10M"; timeout 1500 python tools/bench_synthetic.py 3163 1 16 > gpurun_out/synth_10m.json 2> gpurun_out/synth_10m.err; echo "rc=$?"; cat gpurun_out/synth_10m.json; tail -3 gpurun_out/synth_10m.err
