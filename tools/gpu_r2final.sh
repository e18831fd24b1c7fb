#!/bin/bash
# round 2: full validation + evidence pass on the final code (1 GPU).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== pytest gpu all"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/pytest_gpu_final.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/pytest_gpu_final.log | tail -6
echo "== bench (default)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_ref_final.json
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from tests.test_gpu_pagg import _random_case
from pathnet_b200.train import TrainStep
for variant in ("homo", "hetero"):
    m, X, mask, neis, lt = _random_case(variant, 400, 40, 128, 5, 7, 4, 190, seed=2)   # P = 1330: partial last tile
    m = m.to("cuda:0").train(); m.dropout = 0.5; m.math_mode = "tc"
    y = m(X.to("cuda:0"), neis, 7, 4, mask, lt, None)
    y.square().mean().backward()
    torch.cuda.synchronize()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())
    ts = TrainStep(m, X.to("cuda:0"), mask, torch.zeros(190, dtype=torch.long), 7, 4, graph=False, row_chunk=64)
    ts.step(neis.int(), lt.to(torch.uint8)); ts.step(neis.int(), lt.to(torch.uint8))
    torch.cuda.synchronize()
print("tc path ran")
PY
echo "== memcheck tc"; timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/memcheck_tc.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/memcheck_tc.log
echo "== racecheck tc"; timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py > gpurun_out/racecheck_tc.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/racecheck_tc.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tma|k_gemm_tc|k_pool_fwd128|k_pool_bwd128|k_lstm_fwd_tc|k_lstm_bwd_tc|k_wgrad_tc|k_relu_mask_colsum' -s 52 -c 13 -o gpurun_out/prof_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_final.log 2>&1; echo "rc=$?"
echo "== ncu full merw"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_merw_walk' -s 4 -c 1 -o gpurun_out/prof_merw_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_merw_final.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
