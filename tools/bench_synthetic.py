#!/usr/bin/env python
"""MERW on the synthetic large-graph config of BASELINE.json (configs[4]: 10 M nodes, degree 16,
path_num = 64, path_len = 8): a randomly relabelled 2-D torus with 20 neighbours per node
(tests/helpers.lattice_graph) — memory access is random, radius-6 balls stay small (<= 625 nodes), so
the sparse-label kernel applies (the dense n x n table would need 10^14 bytes).

    python tools/bench_synthetic.py [side=3163] [epochs=1] [check_roots=32]

Prints one JSON line: graph build time, walks/s (labels included), and a bit-exact check of a slice
of start nodes against the oracle.
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tests.helpers import lattice_graph  # noqa: E402
from oracle import merw_oracle as mo  # noqa: E402
import pathnet_b200 as pb  # noqa: E402


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 3163
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    check = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    nw, L = 64, 8
    t0 = time.time()
    n, u, v, p = lattice_graph(side, seed=0)
    t_gen = time.time() - t0
    t0 = time.time()
    g = pb.MerwGraph(n, u, v, p, device="cuda:0")
    t_build = time.time() - t0
    # warm-up on a few start nodes, then the timed full pass in slabs of start nodes
    g.sample(nw, L, epochs=1, seed=7, node_begin=0, node_end=min(n, 4096))
    torch.cuda.synchronize()
    slab = min(n, 2_000_000)
    nodes = torch.empty((epochs, slab, nw, L), dtype=torch.int32, device="cuda:0")
    types = torch.empty((epochs, slab, nw, L), dtype=torch.uint8, device="cuda:0")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    for nb in range(0, n, slab):
        ne = min(n, nb + slab)
        out = (nodes[:, : ne - nb], types[:, : ne - nb])
        ev0.record()
        g.sample(nw, L, epochs=epochs, seed=7, node_begin=nb, node_end=ne, out=out)
        ev1.record()
        torch.cuda.synchronize()
        ms += ev0.elapsed_time(ev1)
    walks = epochs * n * nw
    res = dict(metric="merw_walks_per_sec", workload=f"synthetic torus side={side} n={n} rows={len(u)} path_num={nw} path_len={L} epochs={epochs}",
               value=walks / (ms * 1e-3), unit="walks/s", ms=ms, graph_gen_s=t_gen, graph_build_upload_s=t_build,
               labels="sparse (per-start-node BFS ball in shared memory)", rng="glibc_compat")
    # invariants on the last slab + bit-exact slice against the oracle
    t = types[0, :1000].cpu().numpy()
    assert (t <= np.arange(L)[None, None, :]).all()
    if check:
        lo = n // 2
        got_n, got_t = g.sample(nw, L, epochs=1, seed=7, node_begin=lo, node_end=lo + check)
        t0 = time.time()
        o_n, o_t = mo.OracleGraph(n, u, v, p).sample(nw, L, 7, 0, 1, node_begin=lo, node_end=lo + check)
        res["oracle_slice_s"] = time.time() - t0
        res["bit_exact_slice"] = bool((got_n.cpu().numpy() == o_n).all() and (got_t.cpu().numpy() == o_t).all())
        res["label_histogram"] = np.bincount(o_t.ravel(), minlength=L).tolist()
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
