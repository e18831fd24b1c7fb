#!/bin/bash
# 2-GPU check of the final code: NCCL equivalence tests, weak- and strong-scaling bench lines, and a 1-GPU line on the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
echo "== pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_multi.log
echo "== bench 1 gpu (same box)"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err; echo "rc=$?"
for sc in weak strong; do
echo "== bench --gpus $N $sc"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --scaling $sc --no-cpu-baseline --no-extras > gpurun_out/bench_g${N}_$sc.json 2> gpurun_out/bench_g${N}_$sc.err; echo "rc=$?"; tail -2 gpurun_out/bench_g${N}_$sc.err
done
python - <<PY
import json
for f in ("bench_g1", "bench_g${N}_weak", "bench_g${N}_strong"):
    d = json.load(open(f"gpurun_out/{f}.json")); print(f, d["n_gpus"], d["scaling"], "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"] / 1e6, 2), "M paths/s", "e2e", round(d["e2e"]["ms_per_step"], 4))
PY
