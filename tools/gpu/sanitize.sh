#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from tests.test_gpu_pagg import _random_case
for variant in ("homo", "hetero"):
    m, X, mask, neis, lt = _random_case(variant, 400, 40, 128, 5, 7, 4, 190, seed=2)   # P = 1330: partial last tile
    m = m.to("cuda:0").train(); m.dropout = 0.5; m.math_mode = "tc"
    y = m(X.to("cuda:0"), neis, 7, 4, mask, lt, None)
    y.square().mean().backward()
    torch.cuda.synchronize()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())
print("tc path ran")
PY
echo "== memcheck tc"; timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/memcheck_tc.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/memcheck_tc.log
echo "== racecheck tc"; timeout 100 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py > gpurun_out/racecheck_tc.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/racecheck_tc.log
