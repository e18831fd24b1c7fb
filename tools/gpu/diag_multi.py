"""2-rank diagnostic: prints a stage marker (per rank, flushed) before every step that could block."""
import faulthandler
import os
import sys
import time
from pathlib import Path

faulthandler.dump_traceback_later(int(os.environ.get("DIAG_DEADLINE", "70")), exit=True)
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
t0 = time.time()
rank = int(os.environ.get("RANK", "0"))


def mark(msg):
    print(f"[diag rank {rank} +{time.time() - t0:5.1f}s] {msg}", file=sys.stderr, flush=True)


mark("importing torch")
import torch
import torch.distributed as dist

mark("set_device")
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
mark("init_process_group")
dist.init_process_group("nccl", device_id=dev)
mark("eager all_reduce")
x = torch.ones(1 << 18, device=dev)
dist.all_reduce(x)
torch.cuda.synchronize()
mark(f"all_reduce ok ({x[0].item()})")
from pathnet_b200.dist import shard_bounds
from pathnet_b200.train import TrainStep
from tests.test_gpu_pagg import _random_case

N, Fd, H, Cc, nw, L, S, K = 220, 24, 128, 4, 6, 4, 101, 3
for mode in sys.argv[1:] or ["fp32", "tc", "tc3"]:
    m, X, mask, neis, lt = _random_case("hetero", N, Fd, H, Cc, nw, L, S, seed=41)
    y = torch.randint(0, Cc, (N,), generator=torch.Generator().manual_seed(2))
    m = m.to(dev).train()
    m.math_mode = mode
    lo, hi = shard_bounds(S, rank, dist.get_world_size())
    mark(f"{mode}: building TrainStep")
    ts = TrainStep(m, X.to(dev), mask, y[mask], nw, L, lr=0.005, weight_decay=0.0005, dropout=0.0, graph=True, shard=(lo, hi))
    for k in range(K):
        mark(f"{mode}: step {k}")
        loss = ts.step(neis, lt).item()
    mark(f"{mode}: ok, loss {loss:.6f}")
dist.barrier()
mark("done")
del ts
from pathnet_b200.dist import shutdown
t1 = time.time()
shutdown(15)
mark(f"shutdown returned after {time.time() - t1:.1f}s")
