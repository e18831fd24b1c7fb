#!/bin/bash
# one kernel-iteration round trip: phase traces of one CTA, the tensor-core parity tests, two short bench runs.
# Every leg has a tight timeout and the script stops at the first hang: a deadlocked kernel must not eat the GPU budget.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PN_TRACE_FWD=700 PN_TRACE_BWD=700 timeout 90 python - > gpurun_out/trace.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from tests.test_gpu_pagg import _random_case
N, Fd, H, Cc, nw, L, S = 19717, 500, 128, 3, 40, 4, 9463
m, X, mask, neis, lt = _random_case("homo", N, Fd, H, Cc, nw, L, S, seed=9)
m = m.to("cuda:0").train(); m.dropout = 0.7; m.math_mode = "tc"
Xd = X.to("cuda:0")
for i in range(3):
    y = m(Xd, neis, nw, L, mask, lt, None)
    y.square().mean().backward()
torch.cuda.synchronize()
print("trace run ok")
PY
rc=$?; grep trace gpurun_out/trace.log | tail -16; grep -v trace gpurun_out/trace.log | tail -5
if [ $rc -ne 0 ]; then echo "trace run failed rc=$rc -- stopping"; exit 1; fi
echo "== tests"; timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_v2.py tests/test_gpu_train.py tests/test_gpu_variants.py -q -x --timeout 90 > gpurun_out/pytest_q.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 gpurun_out/pytest_q.log
if [ $rc -eq 124 ]; then echo "tests hung -- stopping"; exit 1; fi
for i in 1 2; do echo "== bench $i"; timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_q$i.json 2> gpurun_out/bench_q$i.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_q$i.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'kern', {k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()})"; tail -2 gpurun_out/bench_q$i.err; done
