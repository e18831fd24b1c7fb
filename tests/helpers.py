"""Shared test helpers: fixture graphs, synthetic messy graphs."""
from __future__ import annotations

import hashlib
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_graph(name):
    z = np.load(GOLDEN / f"graph_{name}.npz")
    return int(z["n"]), z["u"], z["v"], z["p"]


def ref_hashes():
    return json.loads((GOLDEN / "merw_ref_hashes.json").read_text())


def sha256(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def messy_graph(n, avg_deg, seed, hubs=2, neg_frac=0.05, zero_frac=0.05, dup_frac=0.1,
                self_loops=True):
    """Synthetic `.in`-style rows with the pathologies of the shipped inputs
    (SURVEY Appendix B): duplicated rows, self loops, negative / zero p, hubs,
    row sums != 1.  Every node gets at least one out-row."""
    rng = np.random.default_rng(seed)
    rows_u, rows_v = [], []
    m0 = n * avg_deg // 2
    a = rng.integers(0, n, m0)
    b = rng.integers(0, n, m0)
    for h in range(hubs):
        k = min(n, max(8, n // 4))
        a = np.concatenate([a, np.full(k, h)])
        b = np.concatenate([b, rng.integers(0, n, k)])
    rows_u = np.concatenate([a, b])
    rows_v = np.concatenate([b, a])  # symmetric
    if self_loops:
        rows_u = np.concatenate([rows_u, np.arange(n)])
        rows_v = np.concatenate([rows_v, np.arange(n)])
    else:  # still guarantee an out-row per node
        rows_u = np.concatenate([rows_u, np.arange(n)])
        rows_v = np.concatenate([rows_v, (np.arange(n) + 1) % n])
    # duplicates
    d = rng.random(len(rows_u)) < dup_frac
    rows_u = np.concatenate([rows_u, rows_u[d]])
    rows_v = np.concatenate([rows_v, rows_v[d]])
    perm = rng.permutation(len(rows_u))
    u, v = rows_u[perm].astype(np.int32), rows_v[perm].astype(np.int32)
    p = rng.random(len(u))
    deg = np.bincount(u, minlength=n)
    p = p / np.maximum(1, deg[u]) * rng.choice([0.5, 1.0, 2.0, 3.0], size=len(u))
    r = rng.random(len(u))
    p[r < neg_frac] *= -1.0
    p[(r >= neg_frac) & (r < neg_frac + zero_frac)] = 0.0
    return n, u, v, p


def lattice_graph(side, seed=0, shuffle=True):
    """side x side torus, 20 neighbours per node (all cells with |dx|,|dy| <= 2 except the four
    corners |dx|=|dy|=2 and the centre), ids randomly permuted so memory access is random while
    radius-r balls stay small ((4r+1)^2 nodes at most).  p: random positive, rows not normalised."""
    rng = np.random.default_rng(seed)
    n = side * side
    ids = rng.permutation(n).astype(np.int64) if shuffle else np.arange(n, dtype=np.int64)
    x, y = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    us, vs = [], []
    for dx in range(-2, 3):
        for dy in range(-2, 3):
            if (dx == 0 and dy == 0) or (abs(dx) == 2 and abs(dy) == 2):
                continue
            us.append(ids[(x * side + y).ravel()])
            vs.append(ids[(((x + dx) % side) * side + (y + dy) % side).ravel()])
    u = np.concatenate(us).astype(np.int32)
    v = np.concatenate(vs).astype(np.int32)
    perm = rng.permutation(len(u))
    u, v = u[perm], v[perm]
    p = rng.random(len(u)) * 0.12 + 0.001
    return n, u, v, p
