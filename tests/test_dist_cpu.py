"""world_size-2 gloo test (CPU) of the multi-GPU host logic in
pathnet_b200/dist.py: row sharding + loss scaling + one flat all-reduce must
reproduce the single-process gradient.  The per-rank math here is the oracle
(the CUDA operator has no CPU path); what is under test is the sharding, the
1/S_global scaling and the flat-buffer all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pagg_oracle as po
from pathnet_b200.dist import FlatGradAllReduce, shard_bounds, shard_rows


def test_shard_bounds_partition():
    for total in (0, 1, 7, 64, 9463):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _case():
    torch.manual_seed(3)
    N, Fd, H, C, nw, L, S = 90, 12, 32, 4, 5, 3, 41
    import pathnet_b200.pagg as pg
    m = pg.PathNet(Fd, H, C, L)
    X = torch.randn(N, Fd)
    mask = torch.zeros(N, dtype=torch.bool)
    mask[torch.randperm(N)[:S]] = True
    neis = torch.randint(0, N, (S, nw * L))
    lt = torch.randint(0, L, (S, nw, L))
    y = torch.randint(0, C, (N,))
    return m, X, mask, neis, lt, y, nw, L


def _local_grads(rank, world, canonical=True):
    """Gradient of the global mean CE computed from this rank's rows only.  The
    hetero variant couples an output row to other nodes' paths, so shard the
    CANONICAL tensors, not the raw neis."""
    m, X, mask, neis, lt, y, nw, L = _case()
    params = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    S = int(mask.sum())
    sn, st, pt = po.canonicalize("hetero", neis.numpy(), lt.numpy(), nw, L)
    rows = torch.nonzero(mask).reshape(-1)
    lo, hi = shard_bounds(S, rank, world)
    out = po.pagg_forward_canonical("hetero", params, X, sn[lo:hi], st[lo:hi], pt[lo:hi], rows[lo:hi].numpy())
    loss = torch.nn.functional.cross_entropy(out, y[rows[lo:hi]], reduction="sum") / S
    loss.backward()
    return params, loss.detach()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params, loss = _local_grads(rank, world)
    plist = list(params.values())
    FlatGradAllReduce(plist)()
    dist.all_reduce(loss)
    if rank == 0:
        q.put(({k: p.grad.clone() for k, p in params.items()}, loss))
    dist.barrier()
    from pathnet_b200.dist import shutdown
    shutdown(30.0)        # the teardown every multi-rank entry point of the package uses (destroy under a watchdog)
    assert not dist.is_initialized()


def _run_two_ranks():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        grads2, loss2 = q.get(timeout=180)
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    return grads2, loss2


def test_two_rank_gloo_matches_single_process():
    try:
        grads2, loss2 = _run_two_ranks()
    except Exception:   # a rendezvous port race on a busy box: one more attempt on a fresh port
        grads2, loss2 = _run_two_ranks()
    full, loss1 = _local_grads(0, 1)
    torch.testing.assert_close(loss2, loss1, rtol=1e-5, atol=1e-6)
    for k, p in full.items():
        torch.testing.assert_close(grads2[k], p.grad, rtol=1e-4, atol=1e-6, msg=lambda s: f"{k}: {s}")


def test_shard_rows_masks():
    mask = torch.zeros(20, dtype=torch.bool)
    mask[[1, 4, 5, 9, 13, 17, 19]] = True
    seen = torch.zeros_like(mask)
    for r in range(3):
        local, lo, hi = shard_rows(mask, r, 3)
        assert int(local.sum()) == hi - lo and not (seen & local).any()
        seen |= local
    assert torch.equal(seen, mask)
