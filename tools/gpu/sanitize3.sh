#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the split-precision mode: module forward/backward and a row-chunked TrainStep
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cat > /tmp/san3.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from tests.test_gpu_pagg import _random_case
from pathnet_b200.train import TrainStep
for variant in ("homo", "hetero"):
    m, X, mask, neis, lt = _random_case(variant, 400, 40, 128, 5, 7, 4, 190, seed=2)   # P = 1330: partial last tile
    m = m.to("cuda:0").train(); m.dropout = 0.5; m.math_mode = "tc3"
    y = m(X.to("cuda:0"), neis, 7, 4, mask, lt, None)
    y.square().mean().backward()
    torch.cuda.synchronize()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())
    ts = TrainStep(m, X.to("cuda:0"), mask, torch.zeros(190, dtype=torch.long), 7, 4, graph=False, row_chunk=64)
    ts.step(neis.int(), lt.to(torch.uint8)); ts.step(neis.int(), lt.to(torch.uint8))
    torch.cuda.synchronize()
print("tc3 path ran")
PY
echo "== memcheck tc3"; timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san3.py > gpurun_out/memcheck_tc3.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/memcheck_tc3.log
echo "== racecheck tc3"; timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san3.py > gpurun_out/racecheck_tc3.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/racecheck_tc3.log
