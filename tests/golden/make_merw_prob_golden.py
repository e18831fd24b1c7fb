"""Golden vectors for the MERW probability builder from the reference's OWN function.

Authoring container only: imports /root/reference/preprocess/compute_merw.py (never copied) and runs
`compute_merw(A)` (:109-121) on the adjacency of edge_input/cornell.in (unique pairs, A = 1) and on a
small two-hub graph; commits P's values aligned with the unique (row, col) pairs.
    python tests/golden/make_merw_prob_golden.py
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sp

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent


def ref_module():
    spec = importlib.util.spec_from_file_location("ref_compute_merw", "/root/reference/preprocess/compute_merw.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = ref_module()
    z = np.load(OUT / "graph_cornell.npz")
    n, u, v = int(z["n"]), z["u"].astype(np.int64), z["v"].astype(np.int64)
    uk = np.unique(u * n + v)
    cases = {"cornell": (n, uk // n, uk % n)}
    rng = np.random.default_rng(0)                      # small connected graph without self loops
    m = 60
    a = np.concatenate([np.arange(m - 1), rng.integers(0, m, 80), np.zeros(20, dtype=np.int64)])
    b = np.concatenate([np.arange(1, m), rng.integers(0, m, 80), rng.integers(1, m, 20)])
    keep = a != b
    k2 = np.unique(np.concatenate([a[keep] * m + b[keep], b[keep] * m + a[keep]]))
    cases["small"] = (m, k2 // m, k2 % m)
    out = {}
    for name, (nn, uu, vv) in cases.items():
        A = sp.csr_matrix((np.ones(len(uu)), (uu, vv)), shape=(nn, nn))
        P, evector, evalue, _ = ref.compute_merw(A)
        P = P.tocsr()
        out[f"{name}_n"] = nn
        out[f"{name}_u"] = uu.astype(np.int32)
        out[f"{name}_v"] = vv.astype(np.int32)
        out[f"{name}_p"] = np.asarray(P[uu, vv]).ravel()
        out[f"{name}_lambda"] = float(evalue)
        print(name, nn, len(uu), "lambda", evalue, "row sums", P.sum(1).min(), P.sum(1).max())
    np.savez_compressed(OUT / "merw_prob_reference.npz", **out)


if __name__ == "__main__":
    main()
