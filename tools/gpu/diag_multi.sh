#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/gpu/diag_multi.py "$@" > gpurun_out/diag_multi.log 2>&1; echo "rc=$?"
grep -E "diag rank|Error|error|NCCL WARN|Traceback|File " gpurun_out/diag_multi.log | tail -60
