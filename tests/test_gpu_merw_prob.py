"""GPU parity tests of the MERW transition-probability builder (pn_merw_probabilities)."""
import numpy as np
import pytest
import torch

from oracle import merw_oracle as mo
from oracle import merw_prob_oracle as mpo
from pathnet_b200 import merw_prob, sampler
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cornell", "small"])
def test_matches_reference_compute_merw(name):
    z = np.load(GOLDEN / "merw_prob_reference.npz")
    n = int(z[f"{name}_n"])
    ei = np.stack([z[f"{name}_u"], z[f"{name}_v"]]).astype(np.int64)
    p, info = merw_prob.compute_merw(ei, n)
    assert np.abs(p.cpu().numpy() - z[f"{name}_p"]).max() < 1e-9          # fp64 power iteration vs ARPACK
    lam = info["lam"].cpu().numpy()[info["component"].cpu().numpy()[0]]
    assert abs(lam - float(z[f"{name}_lambda"])) < 1e-9
    assert (info["psi"] > 0).all()


def test_duplicates_are_summed_and_output_is_aligned_with_input():
    rng = np.random.default_rng(1)
    n = 200
    a, b = rng.integers(0, n, 900), rng.integers(0, n, 900)
    a, b = np.concatenate([a, np.arange(n - 1)]), np.concatenate([b, np.arange(1, n)])   # connected
    ei = np.stack([np.concatenate([a, b]), np.concatenate([b, a])])
    perm = rng.permutation(ei.shape[1])
    ei = ei[:, perm]
    p, _ = merw_prob.compute_merw(ei, n)
    want, _, _ = mpo.compute_merw(n, ei[0], ei[1])       # csr_matrix sums duplicates like init_rw.py:62-66
    assert np.abs(p.cpu().numpy() - want).max() < 1e-9


def test_disconnected_graph_gets_a_stochastic_matrix_per_component():
    """The reference yields negative / zero p here (cora, citeseer, bgp ...); per-component MERW gives a
    proper transition matrix on every component and equals the reference on each component alone."""
    rng = np.random.default_rng(2)
    blocks, off, us, vs = [50, 31, 7, 1], 0, [], []
    for bsz in blocks:
        if bsz > 1:
            a = np.concatenate([np.arange(bsz - 1), rng.integers(0, bsz, 2 * bsz)])
            b = np.concatenate([np.arange(1, bsz), rng.integers(0, bsz, 2 * bsz)])
            us += [a + off, b + off]; vs += [b + off, a + off]
        us.append(np.arange(off, off + bsz)); vs.append(np.arange(off, off + bsz))       # self loops
        off += bsz
    n = off
    ei = np.stack([np.concatenate(us), np.concatenate(vs)])
    key = np.unique(ei[0].astype(np.int64) * n + ei[1])
    ei = np.stack([key // n, key % n])
    p, info = merw_prob.compute_merw(ei, n)
    p = p.cpu().numpy()
    assert (p > 0).all()
    assert np.allclose(np.bincount(ei[0], weights=p, minlength=n), 1.0, atol=1e-10)
    comp = info["component"].cpu().numpy()
    assert len(np.unique(comp)) == len(blocks)
    lo = 0
    for bsz in blocks[:2]:
        sel = (ei[0] >= lo) & (ei[0] < lo + bsz)
        want, _, _ = mpo.compute_merw(bsz, ei[0][sel] - lo, ei[1][sel] - lo)
        assert np.abs(p[sel] - want).max() < 1e-9
        lo += bsz


def test_probabilities_feed_the_sampler(tmp_path):
    """init_rw.py's file layout (every edge emitted in both directions) -> sampler input -> walks."""
    z = np.load(GOLDEN / "merw_prob_reference.npz")
    n = int(z["small_n"])
    ei = np.stack([z["small_u"], z["small_v"]]).astype(np.int64)
    p, _ = merw_prob.compute_merw(ei, n)
    f = tmp_path / "small.in"
    merw_prob.write_edge_input(f, ei, p, n)
    n2, u, v, pr = sampler.read_edge_input(f)
    assert n2 == n and len(u) == 2 * ei.shape[1]
    assert np.allclose(np.bincount(u, weights=pr, minlength=n), 2.0)      # like the shipped files (Appendix B)
    g = sampler.MerwGraph(n2, u, v, pr)
    nodes, types = g.sample(5, 4, epochs=2, seed=3)
    o_n, o_t = mo.OracleGraph(n2, u, v, pr).sample(5, 4, 3, 0, 2)
    assert (nodes.cpu().numpy() == o_n).all() and (types.cpu().numpy() == o_t).all()
