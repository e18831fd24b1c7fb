"""MERW transition probabilities — host-side mirror of the reference's upstream
step preprocess/init_rw.py (+ compute_merw.py:109-121): adjacency -> `<name>.in`.

    p, info = compute_merw(edge_index, n)            # device tensors
    write_edge_input("cora.in", edge_index, p, n)    # rows in init_rw.py:81-86 order
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib


def compute_merw(edge_index, n, weights=None, tol=1e-13, max_iter=200000, device="cuda:0"):
    """edge_index: [2, M] directed pairs of a SYMMETRIC adjacency (both directions present, duplicates are
    summed like scipy's csr_matrix does in init_rw.py:62-66).  Returns (p [M] float64 aligned with the
    input columns, info dict with psi, lambda, component, iterations)."""
    dev = torch.device(device)
    ei = torch.as_tensor(edge_index).to(dev, dtype=torch.int64)
    M = ei.shape[1]
    w_in = torch.ones(M, dtype=torch.float64, device=dev) if weights is None else torch.as_tensor(weights).to(dev, torch.float64)
    key = ei[0] * n + ei[1]
    ukey, inv = torch.unique(key, return_inverse=True)           # sorted by (row, col): CSR order
    w = torch.zeros(ukey.numel(), dtype=torch.float64, device=dev).index_add_(0, inv, w_in)
    row, col = (ukey // n), (ukey % n).to(torch.int32).contiguous()
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    p = torch.empty(ukey.numel(), dtype=torch.float64, device=dev)
    psi = torch.empty(n, dtype=torch.float64, device=dev)
    lam = torch.zeros(n, dtype=torch.float64, device=dev)
    comp = torch.empty(n, dtype=torch.int32, device=dev)
    iters = C.c_int32(0)
    with torch.cuda.device(dev):
        check(lib().pn_merw_probabilities(n, rowptr.data_ptr(), col.data_ptr(), w.data_ptr(), p.data_ptr(), psi.data_ptr(),
                                          lam.data_ptr(), comp.data_ptr(), float(tol), int(max_iter), C.byref(iters),
                                          torch.cuda.current_stream().cuda_stream))
    return p[inv], dict(psi=psi, lam=lam, component=comp, iterations=iters.value)


def write_edge_input(path, edge_index, p, n):
    """`<name>.in` as init_rw.py:79-86 writes it: header `n 2M`, then for every input edge i the rows
    `u v P[u,v]` and `v u P[v,u]`."""
    ei = torch.as_tensor(edge_index).cpu().numpy()
    M = ei.shape[1]
    import numpy as np
    key = ei[0].astype(np.int64) * n + ei[1]
    order = np.argsort(key, kind="stable")
    pk = torch.as_tensor(p).cpu().numpy()
    rev = pk[order][np.searchsorted(key[order], ei[1].astype(np.int64) * n + ei[0])]
    with open(path, "w") as f:
        f.write(f"{n} {2 * M}\n")
        for i in range(M):
            f.write(f"{ei[0, i]} {ei[1, i]} {float(pk[i])!r}\n{ei[1, i]} {ei[0, i]} {float(rev[i])!r}\n")
