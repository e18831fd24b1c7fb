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
    labels) that belongs to this rank.  (For the hetero variant shard the
    CANONICAL tensors — an output row pools other nodes' paths)"""
    rows = torch.nonzero(mask.reshape(-1), as_tuple=False).reshape(-1)
    lo, hi = shard_bounds(rows.numel(), rank, world)
    local = torch.zeros_like(mask, dtype=torch.bool)
    local[rows[lo:hi]] = True
    return local, lo, hi


def shutdown(timeout_s: float = 20.0) -> None:
    """Tear the default process group down without ever blocking the exit of a finished job: `destroy_process_group()` runs in a
    helper thread; if it has not returned after `timeout_s` (seen on a 2 x B200 box: every rank had finished its work and sat in
    ncclCommDestroy with captured graphs still alive) the process flushes its streams and leaves with status 0.  Call it last."""
    import os
    import sys
    import threading
    if not (dist.is_available() and dist.is_initialized()):
        return
    try:
        torch.cuda.synchronize()
    except Exception:
        pass
    th = threading.Thread(target=dist.destroy_process_group, daemon=True)
    th.start()
    th.join(timeout_s)
    if th.is_alive():
        sys.stdout.flush()
        sys.stderr.write("pathnet_b200.dist.shutdown: destroy_process_group() still blocked after %.0f s; exiting\n" % timeout_s)
        sys.stderr.flush()
        os._exit(0)


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class FlatGradAllReduce:
    """One all-reduce(sum) over all parameter gradients.  Losses must already
    be scaled by 1/S_global so the summed gradient equals the single-process
    gradient.

    Zero-copy path: pn_pagg_backward writes every parameter gradient into ONE
    flat buffer whose slices autograd installs as `p.grad` (pathnet_b200.pagg
    GradSink); when the module's sink is given and the `.grad`s still alias it,
    the collective runs on that buffer in place — no pack / unpack kernels.
    Otherwise (foreign parameters, accumulated or cloned grads, the gloo CPU
    tests) the gradients are packed into a persistent flat buffer and copied
    back."""

    def __init__(self, params, group=None, sink=None):
        self.params = [p for p in params if p.requires_grad]
        self.group, self.sink = group, sink
        self.flat = None
        self.zero_copy_steps = 0

    def _aliases_sink(self) -> bool:
        """Every p.grad is a dense fp32 slice of the sink's flat buffer, and together they cover all of its payload
        (parameter order inside the buffer does not matter for a sum)."""
        s = self.sink
        if s is None or s.flat is None:
            return False
        lo, hi = s.flat.data_ptr(), s.flat.data_ptr() + 4 * s.flat.numel()
        covered = 0
        for p in self.params:
            g = p.grad
            if g is None or g.dtype != torch.float32 or not g.is_contiguous():
                return False
            if not (lo <= g.data_ptr() and g.data_ptr() + 4 * g.numel() <= hi):
                return False
            covered += g.numel()
        return covered == s.payload

    def __call__(self):
        if _world(self.group) > 1 and self._aliases_sink():
            dist.all_reduce(self.sink.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.zero_copy_steps += 1
            return self.sink.flat
        if self.flat is None:
            n = sum(p.numel() for p in self.params)
            self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
            self.views, off = [], 0
            for p in self.params:
                self.views.append(self.flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        if _world(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self.flat
