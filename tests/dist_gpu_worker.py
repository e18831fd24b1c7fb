"""Worker of tests/test_gpu_multi.py: one rank of a 2-GPU data-parallel PAGG training step (NCCL)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pathnet_b200.dist import FlatGradAllReduce, shard_bounds  # noqa: E402
from pathnet_b200.train import TrainStep  # noqa: E402
from tests.test_gpu_pagg import _random_case  # noqa: E402


def main():
    out, variant, mode = sys.argv[1], sys.argv[2], sys.argv[3]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    N, Fd, H, Cc, nw, L, S, K = 220, 24, 128, 4, 6, 4, 101, 3
    m, X, mask, neis, lt = _random_case(variant, N, Fd, H, Cc, nw, L, S, seed=41)     # identical on every rank
    y = torch.randint(0, Cc, (N,), generator=torch.Generator().manual_seed(2))
    m = m.to(dev).train()
    m.math_mode = mode
    lo, hi = shard_bounds(S, rank, world)
    # (1) TrainStep, strong scaling: same batch everywhere, rows [lo, hi) here, in-place all-reduce of the flat gradient buffer
    ts = TrainStep(m, X.to(dev), mask, y[mask], nw, L, lr=0.005, weight_decay=0.0005, dropout=0.0, graph=True, shard=(lo, hi))
    losses = [ts.step(neis, lt).item() for _ in range(K)]
    res = {"losses": losses, "params": {k: p.detach().cpu() for k, p in m.named_parameters()},
           "grads": {k: g.detach().cpu().clone() for k, g in ts.grads().items()}}
    # (2) module + autograd path with the zero-copy flat all-reduce (homo only: raw rows can be sharded)
    if variant == "homo":
        m2, _, _, _, _ = _random_case(variant, N, Fd, H, Cc, nw, L, S, seed=41)
        m2 = m2.to(dev).train()
        m2.math_mode = mode
        red = FlatGradAllReduce(m2.parameters(), sink=m2.grad_sink)
        rows = torch.nonzero(mask).reshape(-1)
        mk = torch.zeros_like(mask)
        mk[rows[lo:hi]] = True
        outp = m2(X.to(dev), neis[lo:hi], nw, L, mk, lt[lo:hi], None)
        loss = torch.nn.functional.cross_entropy(outp, y[mk].to(dev), reduction="sum") / S
        loss.backward()
        red()
        res["module_grads"] = {k: p.grad.detach().cpu() for k, p in m2.named_parameters()}
        res["zero_copy_steps"] = red.zero_copy_steps
    dist.barrier()
    if rank == 0:
        torch.save(res, out)
    del ts                      # the captured graph (it holds an NCCL kernel) goes before the communicator
    from pathnet_b200.dist import shutdown
    shutdown()


if __name__ == "__main__":
    main()
