"""GPU parity tests of the MERW sampler (through the C ABI): bit-exact against
the reference binary's committed hashes, the oracle, and — where the binary
travelled — the live reference on synthetic graphs."""
import numpy as np
import pytest
import torch

from oracle import merw_oracle as mo
from pathnet_b200 import _lib, sampler
from tests.conftest import GOLDEN
from tests.helpers import load_graph, messy_graph, ref_hashes, sha256

pytestmark = pytest.mark.gpu


def _gpu_sample(g, nw, L, seed, **kw):
    nodes, types = g.sample(nw, L, seed=seed, **kw)
    torch.cuda.synchronize()
    return nodes.cpu().numpy(), types.cpu().numpy()


@pytest.mark.parametrize("key", [k for k in ref_hashes() if k.split("|")[0] in ("cornell", "cora")])
def test_matches_reference_binary_hashes(key):
    name, nw, L, seed, epochs = key.split("|")
    nw, L, seed, epochs = (int(x.split("=")[1]) for x in (nw, L, seed, epochs))
    g = sampler.MerwGraph(*load_graph(name))
    nodes, types = _gpu_sample(g, nw, L, seed, epochs=epochs)
    txt = sampler.format_paths_txt(nodes, types)
    want = ref_hashes()[key]
    assert len(txt) == want["bytes"] and sha256(txt) == want["sha256"]


def test_matches_parsed_reference_fixture():
    z = np.load(GOLDEN / "merw_cornell_40_4_s12345_e0.npz")
    g = sampler.MerwGraph(*load_graph("cornell"))
    nodes, types = _gpu_sample(g, 40, 4, 12345, epochs=1)
    assert (nodes[0] == z["nodes"]).all() and (types[0] == z["types"]).all()


@pytest.mark.parametrize("n,deg,nw,L,E,seed", [
    (300, 8, 5, 3, 3, 1), (300, 8, 40, 4, 2, 2), (500, 6, 9, 5, 2, 3), (400, 10, 7, 6, 2, 4),
    (257, 4, 1, 8, 3, 5), (64, 3, 33, 2, 2, 6), (100, 5, 3, 1, 2, 7),
])
def test_matches_oracle_on_messy_graphs(n, deg, nw, L, E, seed):
    gr = messy_graph(n, deg, seed)
    want_n, want_t = mo.OracleGraph(*gr).sample(nw, L, 1000 + seed, 0, E)
    got_n, got_t = _gpu_sample(sampler.MerwGraph(*gr), nw, L, 1000 + seed, epochs=E)
    assert (got_n == want_n).all()
    assert (got_t == want_t).all()


def test_matches_live_reference_binary(tmp_path):
    if not mo.have_ref():
        pytest.skip("oracle/_ref not present on this box")
    n, u, v, p = messy_graph(600, 7, 21)
    f = tmp_path / "messy.in"
    mo.write_edge_input(f, n, u, v, p)
    txt = mo.run_reference(f, 11, 6, 4711, epochs=1, name="messy")
    nodes, types = _gpu_sample(sampler.MerwGraph(n, u, v, p), 11, 6, 4711, epochs=1)
    assert sampler.format_paths_txt(nodes, types) == txt


def test_epoch_and_node_ranges_tile_the_full_output():
    gr = messy_graph(350, 6, 9)
    g = sampler.MerwGraph(*gr)
    full_n, full_t = _gpu_sample(g, 6, 4, 77, epochs=5)
    for e in (0, 3, 4):
        n1, t1 = _gpu_sample(g, 6, 4, 77, epochs=1, epoch_begin=e)
        assert (n1[0] == full_n[e]).all() and (t1[0] == full_t[e]).all()
    a_n, a_t = _gpu_sample(g, 6, 4, 77, epochs=5, node_begin=0, node_end=123)
    b_n, b_t = _gpu_sample(g, 6, 4, 77, epochs=5, node_begin=123, node_end=350)
    assert (np.concatenate([a_n, b_n], axis=1) == full_n).all()
    assert (np.concatenate([a_t, b_t], axis=1) == full_t).all()
    # late epochs of a long run (64-bit stream offsets) against the oracle
    o_n, o_t = mo.OracleGraph(*gr).sample(6, 4, 77, 997, 1000)
    l_n, l_t = _gpu_sample(g, 6, 4, 77, epochs=3, epoch_begin=997)
    assert (l_n == o_n).all() and (l_t == o_t).all()


def test_invariants_at_pubmed_scale():
    """Size-independent properties at BASELINE scale (n = 19 717, nw = 40, L = 4):
    walks start at their node, every transition is a row of the graph, labels
    are hop distances (<= position, 0 iff the start node) and the whole thing is
    deterministic; a slice is compared with the oracle bit for bit."""
    n = 19717
    gr = messy_graph(n, 11, 5, hubs=3, neg_frac=0.0, zero_frac=0.0)
    g = sampler.MerwGraph(*gr)
    nodes, types = _gpu_sample(g, 40, 4, 2024, epochs=2)
    assert nodes.shape == (2, n, 40, 4)
    assert (nodes[:, :, :, 0] == np.arange(n)[None, :, None]).all()
    _, u, v, _ = gr
    edge_keys = np.unique(u.astype(np.int64) * n + v)
    for t in range(3):
        k = nodes[..., t].astype(np.int64) * n + nodes[..., t + 1]
        assert np.isin(k.ravel(), edge_keys).all()
    assert (types <= np.arange(4)[None, None, None, :]).all()
    assert ((types == 0) == (nodes == np.arange(n)[None, :, None, None])).all()
    again_n, again_t = _gpu_sample(g, 40, 4, 2024, epochs=2)
    assert (again_n == nodes).all() and (again_t == types).all()
    o_n, o_t = mo.OracleGraph(*gr).sample(40, 4, 2024, 0, 1)
    assert (o_n[0] == nodes[0]).all() and (o_t[0] == types[0]).all()


def test_bgp_scale_hub_graph_L6_slice_matches_oracle():
    """bgp-shaped: n = 63 977, ~700 k rows, hubs with thousands of rows, negative and
    zero p, L = 6 (dense 4.1 GB distance table, 3 relaxation sweeps).  The oracle
    labels a slice of start nodes (its per-root BFS over a hub graph is what costs
    the reference 253 s of setup); the GPU output for that slice must match bit for bit."""
    n = 63977
    gr = messy_graph(n, 10, 13, hubs=3, neg_frac=0.2, zero_frac=0.4, dup_frac=0.03, self_loops=False)
    g = sampler.MerwGraph(*gr)
    nodes, types = _gpu_sample(g, 40, 6, 4242, epochs=1, node_begin=31000, node_end=31400)
    o_n, o_t = mo.OracleGraph(*gr).sample(40, 6, 4242, 0, 1, node_begin=31000, node_end=31400)
    assert (nodes == o_n).all()
    assert (types == o_t).all()
    assert types.max() <= 5


def test_gen_merw_returns_the_tensors_pathnet_run_builds(tmp_path):
    n, u, v, p = load_graph("cornell")
    f = tmp_path / "cornell.in"
    mo.write_edge_input(f, n, u, v, p)
    neis, ptype = sampler.gen_merw(str(f), 40, 4, epochs=2, seed=12345)
    assert neis.shape == (2, 183, 160) and ptype.shape == (2, 183, 40, 4)
    assert neis.dtype == torch.int64 and ptype.dtype == torch.int64
    z = np.load(GOLDEN / "merw_cornell_40_4_s12345_e0.npz")
    assert (neis[0].cpu().numpy().reshape(183, 40, 4) == z["nodes"]).all()


def test_dead_end_graph_is_an_error():
    g = sampler.MerwGraph(3, [0, 1], [1, 0], [1.0, 1.0])
    with pytest.raises(_lib.PnError) as ei:
        g.sample(2, 3)
    assert ei.value.code == -3


# ---------------------------------------------------------------- sparse-label (root-centric) path --
@pytest.fixture
def force_sparse(monkeypatch):
    monkeypatch.setenv("PN_MERW_SPARSE", "1")


@pytest.mark.parametrize("n,deg,nw,L,E,seed", [(300, 8, 5, 3, 3, 1), (300, 8, 40, 4, 2, 2), (257, 4, 33, 4, 2, 5),
                                              (64, 3, 70, 4, 2, 6)])
def test_sparse_label_path_matches_oracle_on_messy_graphs(force_sparse, n, deg, nw, L, E, seed):
    gr = messy_graph(n, deg, seed)
    want_n, want_t = mo.OracleGraph(*gr).sample(nw, L, 2000 + seed, 0, E)
    got_n, got_t = _gpu_sample(sampler.MerwGraph(*gr), nw, L, 2000 + seed, epochs=E)
    assert (got_n == want_n).all()
    assert (got_t == want_t).all()


@pytest.mark.parametrize("side,nw,L,E", [(40, 64, 8, 2), (33, 40, 6, 3), (25, 7, 5, 2)])
def test_sparse_label_path_matches_oracle_on_lattices(force_sparse, side, nw, L, E):
    from tests.helpers import lattice_graph
    gr = lattice_graph(side, seed=side)
    g = sampler.MerwGraph(*gr)
    want_n, want_t = mo.OracleGraph(*gr).sample(nw, L, 99, 1, 1 + E)
    got_n, got_t = _gpu_sample(g, nw, L, 99, epochs=E, epoch_begin=1)
    assert (got_n == want_n).all()
    assert (got_t == want_t).all()
    assert got_t.max() == L - 1                       # far enough to leave the radius-(L-2) ball sometimes
    a_n, a_t = _gpu_sample(g, nw, L, 99, epochs=E, epoch_begin=1, node_begin=100, node_end=333)
    assert (a_n == want_n[:, 100:333]).all() and (a_t == want_t[:, 100:333]).all()


def test_sparse_and_dense_paths_agree(monkeypatch):
    from tests.helpers import lattice_graph
    gr = lattice_graph(60, seed=4)
    g = sampler.MerwGraph(*gr)
    d_n, d_t = _gpu_sample(g, 16, 6, 5, epochs=2)
    monkeypatch.setenv("PN_MERW_SPARSE", "1")
    s_n, s_t = _gpu_sample(g, 16, 6, 5, epochs=2)
    assert (d_n == s_n).all() and (d_t == s_t).all()


def test_sparse_label_path_reports_ball_overflow(force_sparse):
    gr = messy_graph(20000, 12, 3, hubs=4)            # radius-4 balls cover the whole graph
    g = sampler.MerwGraph(*gr)
    with pytest.raises(_lib.PnError) as ei:
        g.sample(4, 6, epochs=1, node_begin=0, node_end=64)
    assert ei.value.code == -5 and "ball" in str(ei.value)


# --------------------------------------------------- uniform random walk variant (preprocess/gen.cpp) --
def test_uniform_rw_matches_reference_gen_fixture_and_oracle(monkeypatch):
    z = np.load(GOLDEN / "uniform_rw_reference.npz")
    n, u, v, nw, L, seed, E = int(z["n"]), z["u"], z["v"], int(z["nw"]), int(z["L"]), int(z["seed"]), int(z["epochs"])
    g = sampler.MerwGraph(n, u, v, None)
    nodes, types = _gpu_sample(g, nw, L, seed, epochs=E)
    txt = sampler.format_paths_txt(nodes, types)
    assert len(txt) == int(z["nbytes"]) and sha256(txt) == str(z["sha256"])          # == reference gen binary
    monkeypatch.setenv("PN_MERW_SPARSE", "1")                                          # and through the sparse-label kernel
    s_n, s_t = _gpu_sample(g, nw, L, seed, epochs=E)
    assert (s_n == nodes).all() and (s_t == types).all()


def test_uniform_rw_on_a_larger_graph_matches_oracle():
    rng = np.random.default_rng(5)
    n, m = 5000, 30000
    u, v = rng.integers(0, n, m).astype(np.int32), rng.integers(0, n, m).astype(np.int32)
    want_n, want_t = mo.OracleGraph(n, u, v).sample(40, 4, 321, 0, 2)
    got_n, got_t = _gpu_sample(sampler.MerwGraph(n, u, v, None), 40, 4, 321, epochs=2)
    assert (got_n == want_n).all() and (got_t == want_t).all()


# ---------------------------------------------------------------- rng_mode = philox (PN_RNG_PHILOX) --
@pytest.mark.parametrize("n,deg,nw,L,E,seed", [(300, 8, 5, 3, 3, 1), (300, 8, 40, 4, 2, 2), (500, 6, 9, 5, 2, 3),
                                              (257, 4, 1, 8, 3, 5), (64, 3, 33, 2, 2, 6), (100, 5, 3, 1, 2, 7)])
def test_philox_mode_matches_oracle_on_messy_graphs(n, deg, nw, L, E, seed):
    gr = messy_graph(n, deg, seed)
    big_seed = (0x1234ABCD << 32) | (5000 + seed)          # exercises both key words
    want_n, want_t = mo.OracleGraph(*gr).sample_philox(nw, L, big_seed, 0, E)
    got_n, got_t = _gpu_sample(sampler.MerwGraph(*gr), nw, L, big_seed, epochs=E, rng="philox")
    assert (got_n == want_n).all()
    assert (got_t == want_t).all()


def test_philox_mode_is_invariant_under_epoch_and_node_splits_and_label_paths(monkeypatch):
    gr = messy_graph(350, 6, 9)
    g = sampler.MerwGraph(*gr)
    full_n, full_t = _gpu_sample(g, 6, 4, 77, epochs=5, rng="philox")
    n1, t1 = _gpu_sample(g, 6, 4, 77, epochs=2, epoch_begin=3, rng="philox")
    assert (n1 == full_n[3:]).all() and (t1 == full_t[3:]).all()
    a_n, a_t = _gpu_sample(g, 6, 4, 77, epochs=5, node_begin=0, node_end=123, rng="philox")
    b_n, b_t = _gpu_sample(g, 6, 4, 77, epochs=5, node_begin=123, node_end=350, rng="philox")
    assert (np.concatenate([a_n, b_n], axis=1) == full_n).all()
    assert (np.concatenate([a_t, b_t], axis=1) == full_t).all()
    # a far epoch (the 64-bit walk index reaches the counter's high word only beyond 2^32 walks; check the oracle there)
    far = (1 << 32) // (350 * 6) + 5
    o_n, o_t = mo.OracleGraph(*gr).sample_philox(6, 4, 77, far, far + 2)
    l_n, l_t = _gpu_sample(g, 6, 4, 77, epochs=2, epoch_begin=far, rng="philox")
    assert (l_n == o_n).all() and (l_t == o_t).all()
    # the root-centric (sparse label) kernel draws the same walks
    monkeypatch.setenv("PN_MERW_SPARSE", "1")
    s_n, s_t = _gpu_sample(sampler.MerwGraph(*gr), 6, 4, 77, epochs=5, rng="philox")
    assert (s_n == full_n).all() and (s_t == full_t).all()
    # and it is a different stream from the reference-compatible one, with the same invariants
    gl_n, _ = _gpu_sample(g, 6, 4, 77, epochs=5)
    assert (gl_n != full_n).any()
    assert (full_n[..., 0] == np.arange(350)[None, :, None]).all() and (full_t[..., 0] == 0).all()


def test_philox_mode_walks_follow_the_transition_law():
    """Same law as the reference's roll (x uniform over the cells, A with probability 1 - S): the empirical next-node
    frequencies from one start node agree between the two rng modes (and with the cells) to sampling error."""
    gr = messy_graph(120, 6, 17, neg_frac=0.0, zero_frac=0.0)
    g = sampler.MerwGraph(*gr)
    ph, _ = _gpu_sample(g, 64, 2, 5, epochs=400, rng="philox")
    gl, _ = _gpu_sample(g, 64, 2, 5, epochs=400)
    for st in (0, 7, 63):
        a = np.bincount(ph[:, st, :, 1].ravel(), minlength=120) / (400 * 64)
        b = np.bincount(gl[:, st, :, 1].ravel(), minlength=120) / (400 * 64)
        assert np.abs(a - b).max() < 0.02
