#!/bin/bash
# in-kernel phase timestamps of one CTA of the split-precision forward (pubmed-shaped call, train mode)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PN_TRACE_FWD=${1:-700} timeout 300 python - > gpurun_out/trace3.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from tests.test_gpu_pagg import _random_case
N, Fd, H, Cc, nw, L, S = 19717, 500, 128, 3, 40, 4, 9463
m, X, mask, neis, lt = _random_case("homo", N, Fd, H, Cc, nw, L, S, seed=9)
m = m.to("cuda:0").train(); m.dropout = 0.7; m.math_mode = "tc3"
Xd = X.to("cuda:0")
for i in range(3):
    y = m(Xd, neis, nw, L, mask, lt, None)
    y.square().mean().backward()
torch.cuda.synchronize()
PY
grep trace gpurun_out/trace3.log | tail -8; grep -v trace gpurun_out/trace3.log | tail -5
