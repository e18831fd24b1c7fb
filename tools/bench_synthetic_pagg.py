#!/usr/bin/env python
"""PAGG (forward + backward) on the synthetic large-graph config of BASELINE.json (configs[4]: 10 M nodes, degree ~16-20,
path_num = 64, path_len = 8, F = H = 128, random 48 % train mask): 3.07e8 paths per step on ONE B200, through the
public operator with row-chunked execution (pn_pagg_cfg.row_chunk: one Z / T table, per-sequence workspace sized for a
chunk of rows, backward recomputes each chunk's forward).  Paths come from the on-GPU MERW sampler (sparse labels).

    python tools/bench_synthetic_pagg.py [side=3163] [train_frac=0.48] [row_chunk=4096] [steps=2] [check_rows=24]

Prints one JSON line: paths/s fwd+bwd, per-phase times, peak memory, and a slice of output rows checked against the
oracle (eval mode, on the compact sub-problem of the nodes those rows touch).
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tests.helpers import lattice_graph  # noqa: E402
from oracle import pagg_oracle as po  # noqa: E402
import pathnet_b200 as pb  # noqa: E402
from pathnet_b200 import pagg  # noqa: E402


def main():
    a = sys.argv[1:]
    side = int(a[0]) if len(a) > 0 else 3163
    frac = float(a[1]) if len(a) > 1 else 0.48
    row_chunk = int(a[2]) if len(a) > 2 else 4096
    steps = int(a[3]) if len(a) > 3 else 2
    check = int(a[4]) if len(a) > 4 else 24
    nw, L, F, H, C = 64, 8, 128, 128, 8
    dev = torch.device("cuda:0")
    t0 = time.time()
    n, u, v, p = lattice_graph(side, seed=0)
    g = pb.MerwGraph(n, u, v, p, device=dev)
    del u, v, p
    t_graph = time.time() - t0
    gen = torch.Generator().manual_seed(1)
    mask = torch.zeros(n, dtype=torch.bool)
    S = int(round(n * frac))
    mask[torch.randperm(n, generator=gen)[:S]] = True
    maskd = mask.to(dev)
    rows = torch.nonzero(maskd).reshape(-1).to(torch.int32)
    # sample in slabs of start nodes, keep only the train rows (int32 node ids / uint8 labels: no int64 anywhere)
    neis = torch.empty((S, nw * L), dtype=torch.int32, device=dev)
    ltype = torch.empty((S, nw, L), dtype=torch.uint8, device=dev)
    slab, done = 1_000_000, 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for nb in range(0, n, slab):
        ne = min(n, nb + slab)
        nodes, types = g.sample(nw, L, epochs=1, seed=7, node_begin=nb, node_end=ne)
        m = maskd[nb:ne]
        k = int(m.sum())
        neis[done:done + k] = nodes[0].view(ne - nb, nw * L)[m]
        ltype[done:done + k] = types[0][m]
        done += k
        del nodes, types
    ev1.record()
    torch.cuda.synchronize()
    t_sample = ev0.elapsed_time(ev1) * 1e-3
    assert done == S
    del g
    torch.cuda.empty_cache()
    X = torch.randn(n, F, generator=torch.Generator().manual_seed(0)).to(dev)
    y = torch.randint(0, C, (S,), generator=gen).to(dev)
    torch.manual_seed(0)
    model = pb.PathNet_homo(F, H, C, L, dropout=0.7, math_mode="tc", row_chunk=row_chunk).to(dev).train()
    seq_node, seq_type, partner = pagg.canonicalize("homo", neis, ltype, nw, L, N=n)
    del neis, ltype
    torch.cuda.empty_cache()
    sd = dict(model.named_parameters())
    P = S * nw
    times = []
    for it in range(steps + 1):
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out = pagg.pagg_forward("homo", X, sd, seq_node, seq_type, partner, rows, dropout=0.7, training=True, seed=1000 + it,
                                math_mode="tc", row_chunk=row_chunk)
        loss = torch.nn.functional.cross_entropy(out, y)
        e[1].record()
        model.zero_grad(set_to_none=True)
        loss.backward()
        e[2].record()
        torch.cuda.synchronize()
        if it > 0:
            times.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
        assert torch.isfinite(loss).item() and all(torch.isfinite(q.grad).all().item() for q in model.parameters())
    fwd = sum(t[0] for t in times) / len(times)
    bwd = sum(t[1] for t in times) / len(times)
    res = dict(metric="pagg_paths_per_sec_fwd_bwd",
               workload=f"synthetic torus side={side} n={n} F={F} H={H} C={C} path_num={nw} path_len={L} train rows S={S} ({frac:.0%}) "
                        f"-> {P} paths/step, PathNet_homo, train mode dropout 0.7, math_mode tc, row_chunk={row_chunk}",
               value=P / ((fwd + bwd) * 1e-3), unit="paths/s", fwd_ms=fwd, bwd_ms=bwd, steps=len(times), n_gpus=1,
               chunks=int(-(-S // row_chunk)), peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30,
               graph_s=t_graph, sample_s=t_sample, sample_walks_per_s=n * nw / t_sample,
               note="backward includes the per-chunk forward recompute (activation checkpointing by row chunk); "
                    "fwd+bwd per path = 48*L*H^2 = 6.29 MFLOP algorithmic (+16*L*H^2 recompute)")
    flops = 48.0 * L * H * H * P
    res["algorithmic_tflops"] = flops / ((fwd + bwd) * 1e-3) / 1e12
    if check:
        # eval-mode logits of a slice of rows against the oracle, on the compact sub-problem of the nodes they touch
        model.eval()
        lo = S // 2
        sl = slice(lo, lo + check)
        with torch.no_grad():
            got = pagg.pagg_forward("homo", X, sd, seq_node[sl].contiguous(), seq_type[sl].contiguous(), partner[sl].contiguous(),
                                    rows[sl].contiguous(), training=False, math_mode="tc", row_chunk=7).cpu()
        sn = seq_node[sl].cpu().numpy().astype(np.int64)
        rw = rows[sl].cpu().numpy().astype(np.int64)
        uniq, inv = np.unique(np.concatenate([sn.ravel(), rw]), return_inverse=True)
        sn_c = inv[: sn.size].reshape(sn.shape)
        rw_c = inv[sn.size:]
        params = {k: q.detach().cpu().double() for k, q in sd.items()}
        want = po.pagg_forward_canonical("homo", params, X[torch.as_tensor(uniq, device=dev)].cpu().double(), sn_c.astype(np.int32),
                                         seq_type[sl].cpu().numpy(), sn_c[:, :, 0].astype(np.int32), rw_c).float()
        err = (got - want).abs().max().item()
        res["oracle_slice"] = dict(rows=check, max_abs_err=err, max_abs_logit=want.abs().max().item(), ok=bool(err < 3e-3))
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
