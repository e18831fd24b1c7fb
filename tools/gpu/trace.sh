#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PN_TRACE_FWD=700 PN_TRACE_BWD=700 timeout 300 python - > gpurun_out/trace.log 2>&1 <<'PY'
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
PY
grep trace gpurun_out/trace.log | tail -11; grep -v trace gpurun_out/trace.log | tail -5
echo "== pytest tc (gemm + lstm)"; timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_tc.log
echo "== bench tc"; timeout 600 python bench.py --steps 10 --warmup 3 --math tc --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step']); print(d['roofline']['kernels']); print('merw', d['merw']['value'], d['merw']['ms_per_step'])"
