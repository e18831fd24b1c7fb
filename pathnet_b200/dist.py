"""Multi-GPU data parallelism of the PAGG training step (new work — the
reference is single-process): output rows shard across ranks
(rows are independent after canonicalisation), X and the parameters are
replicated, and the only exchange is one all-reduce of a flat fp32 gradient
buffer per step (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) of `total` items for `rank`."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(mask: torch.Tensor, rank: int, world: int):
    """Split the True entries of a node mask across ranks.  Returns the local
    mask and the [lo, hi) slice of the row-aligned tensors (neis, layer_type,
    labels) that belongs to this rank."""
    rows = torch.nonzero(mask.reshape(-1), as_tuple=False).reshape(-1)
    lo, hi = shard_bounds(rows.numel(), rank, world)
    local = torch.zeros_like(mask, dtype=torch.bool)
    local[rows[lo:hi]] = True
    return local, lo, hi


class FlatGradAllReduce:
    """One all-reduce(sum) over all parameter gradients, through a persistent
    flat buffer.  Losses must already be scaled by 1/S_global so the summed
    gradient equals the single-process gradient."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(n, dtype=torch.float32, device=p0.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def __call__(self):
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self.flat
