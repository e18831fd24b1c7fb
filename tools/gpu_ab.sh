#!/bin/bash
# A/B a kernel switch: TC tests (short timeouts), then the bench once per environment setting given as arguments.
#   tools/gpu_ab.sh PN_BWD_CLUSTER=1 PN_BWD_CLUSTER=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest tc"; timeout 240 python -m pytest tests/test_gpu_tc.py tests/test_gpu_pagg.py -m gpu -q -x --timeout 60 > gpurun_out/pytest_tc.log 2>&1; rc=$?; echo "rc=$rc"; tail -3 gpurun_out/pytest_tc.log
[ $rc -ne 0 ] && { grep -E "Error|error|FAILED" gpurun_out/pytest_tc.log | head; exit 1; }
for kv in "$@"; do
  echo "== bench $kv"; env $kv timeout 300 python bench.py --steps 10 --warmup 3 --math tc --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step']); print({k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()})"
done
