"""TEST INFRASTRUCTURE ONLY — restatement of the reference's MERW probability
builder `compute_merw` (/root/reference/preprocess/compute_merw.py:109-121):
dominant eigenpair of the symmetric adjacency by scipy's eigsh (the reference's
own dependency, pinned to 1.4.1 there, 1.18 here), then
P[u, v] = A[u, v] * psi[v] / (lambda * psi[u]).

Pinned against the reference's OWN outputs: the p column of the shipped
edge_input/cornell.in (tests/golden/graph_cornell.npz) and pubmed.in are what
this function produced for the authors (tests/test_oracle_pinning.py)."""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import eigsh


def compute_merw(n, u, v, w=None):
    """COO (duplicates summed, as scipy's csr_matrix does) -> p per input entry, psi, lambda."""
    w = np.ones(len(u)) if w is None else w
    A = sp.csr_matrix((w, (u, v)), shape=(n, n))
    A.sum_duplicates()
    vals, vecs = eigsh(A.astype(np.float64), 1, which="LA", tol=1e-14)     # :111 (symmetric)
    lam, psi = vals[0], vecs[:, 0]
    psi = psi / np.linalg.norm(psi)                                        # :114
    if psi.sum() < 0:
        psi = -psi
    Au = np.asarray(A[u, v]).ravel()
    return Au * psi[v] / (lam * psi[u]), psi, lam                          # :118-120
