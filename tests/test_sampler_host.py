"""CPU tests of the sampler's host logic (no GPU): alias cells + thresholds
against the oracle, glibc jump-ahead tables against the sequential stream,
text format round trip, C-ABI exports."""
import ctypes as C

import numpy as np
import pytest

from oracle import merw_oracle as mo
from pathnet_b200 import _lib, sampler
from tests.helpers import load_graph, messy_graph

RAND_MAX = 2147483647


def _cells_match(n, u, v, p):
    g = sampler.MerwGraph(n, u, v, p, device="host")
    off, A, B, S, thr = g.cells_host()
    o_off, o_A, o_B, o_S = mo.OracleGraph(n, u, v, p).cells()
    assert (off == o_off).all() and (A == o_A).all() and (B == o_B).all()
    assert (S == o_S).all() or (np.isnan(S) == np.isnan(o_S)).all()
    return S, thr


@pytest.mark.parametrize("name", ["cornell", "cora"])
def test_alias_cells_real_graphs(name):
    S, thr = _cells_match(*load_graph(name))
    assert len(S) > 0


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_alias_cells_messy(seed):
    _cells_match(*messy_graph(300, 8, seed))


def test_threshold_equivalence():
    """(1.0*r/RAND_MAX > S) == (r > thr) for r around every cell boundary."""
    n, u, v, p = messy_graph(500, 10, 7, neg_frac=0.1, zero_frac=0.1)
    S, thr = _cells_match(n, u, v, p)
    extra = np.array([0, 1, RAND_MAX - 1, RAND_MAX], dtype=np.int64)
    for s, t in zip(S[:4000], thr[:4000]):
        rs = np.unique(np.clip(np.concatenate([np.arange(int(t) - 2, int(t) + 3, dtype=np.int64), extra]), 0, RAND_MAX))
        ref = (1.0 * rs.astype(np.float64) / RAND_MAX) > s
        assert ((rs > t) == ref).all(), (s, t)


def test_dead_end_reported():
    g = sampler.MerwGraph(3, [0, 1], [1, 0], [1.0, 1.0], device="host")
    assert g.info()["dead_end"] == 2


def test_bad_edges_rejected():
    with pytest.raises(_lib.PnError):
        sampler.MerwGraph(3, [0, 5], [1, 0], [1.0, 1.0], device="host")


def _window_rands(win, count):
    """generate `count` rand() outputs from a 31-word window"""
    s = [int(x) for x in win]
    out = []
    for _ in range(count):
        v = (s[0] + s[28]) & 0xFFFFFFFF
        s = s[1:] + [v]
        out.append(v >> 1)
    return out


def _apply(poly, win):
    e = [int(x) for x in win]
    for d in range(30):
        e.append((e[d] + e[d + 28]) & 0xFFFFFFFF)
    return [sum(int(poly[i]) * e[i + j] for i in range(31)) & 0xFFFFFFFF for j in range(31)]


@pytest.mark.parametrize("cfg", [
    # n, nw, L, epoch_begin, E, node_begin, n'
    (183, 40, 4, 0, 3, 0, 183),
    (183, 40, 4, 5, 4, 17, 50),
    (50, 7, 3, 2, 5, 3, 40),
    (1000, 64, 8, 1000, 2, 0, 1000),
    (19717, 40, 4, 999, 1, 9000, 10717),
])
def test_jump_ahead_tables(cfg):
    """Emulates exactly what k_merw_walk does with the tables (bits of the chunk
    and epoch index, then step / wrap from lane to lane) and checks each
    chunk's window generates the sequential srand(seed) stream at the offset
    the reference reaches for that walk."""
    n, nw, L, eb, E, nb, npr = cfg
    seed = 4242
    wpt, cpe = C.c_int32(), C.c_int64()
    tables = np.zeros((3 + 32 + 120 + 32, 31), dtype=np.uint32)
    _lib.check(_lib.lib().pn_merw_plan_host(n, nw, L, eb, E, nb, npr, seed, C.byref(wpt),
                                            C.byref(cpe), tables.ctypes.data))
    wpt, cpe = wpt.value, cpe.value
    assert cpe == -(-npr * nw // wpt)
    base, jstep, jwrap = tables[0], tables[1], tables[2]
    jlane, jchunk16, jepoch = tables[3:35], tables[35:155].reshape(8, 15, 31), tables[155:187]
    total = E * cpe
    draws = 2 * L
    last_off = draws * (((eb + E - 1) * n + nb) * nw + npr * nw)
    rng = np.random.default_rng(0)
    starts = sorted(set([0, max(0, total - 40)] + [int(x) for x in rng.integers(0, total, 4)] +
                        ([max(0, cpe - 5)] if E > 1 else [])))
    for q0 in starts:
        e, c = divmod(q0, cpe)
        w = list(base)
        cb, i = c, 0
        while cb:
            if cb & 15:
                w = _apply(jchunk16[i][(cb & 15) - 1], w)
            cb >>= 4
            i += 1
        for b in range(32):
            if (e >> b) & 1:
                w = _apply(jepoch[b], w)
        fast = c + 31 < cpe
        for lane in range(32):
            if q0 + lane >= total:
                break
            if fast:   # every lane applies its own polynomial to the warp's window
                off = draws * (((eb + e) * n + nb) * nw + (c + lane) * wpt)
                assert _window_rands(_apply(jlane[lane], w), 6) == mo.rand_stream_at(seed, off, 6), (q0, lane)
                continue
            off = draws * (((eb + e) * n + nb) * nw + c * wpt)
            want = mo.rand_stream_at(seed, off, 6)
            assert _window_rands(w, 6) == want, (q0, lane, e, c)
            if c + 1 < cpe:
                w = _apply(jstep, w); c += 1
            else:
                w = _apply(jwrap, w); c = 0; e += 1
    assert last_off > 0


def test_text_format_roundtrip_and_oracle_agreement():
    rng = np.random.default_rng(3)
    nodes = rng.integers(0, 100000, (257, 6)).astype(np.int32)
    types = rng.integers(0, 6, (257, 6)).astype(np.uint8)
    txt = sampler.format_paths_txt(nodes, types)
    assert txt == mo.format_txt(nodes, types)
    n2, t2 = sampler.parse_paths_txt(txt, 6)
    assert (n2 == nodes).all() and (t2 == types).all()
    assert txt.startswith(b"[") and txt.endswith(b"]\n")
    # the reference's reader: line[1:-2].split(",")
    first = txt.split(b"\n")[0] + b"\n"
    fields = [int(x) for x in first[1:-2].split(b",")]
    assert fields == list(nodes[0]) + list(types[0])


def test_parse_rejects_garbage():
    with pytest.raises(_lib.PnError):
        sampler.parse_paths_txt(b"[1, 2, 3]\n", 4)
    with pytest.raises(_lib.PnError):
        sampler.parse_paths_txt(b"hello\n", 2)


def test_text_io_is_independent_of_the_host_thread_count(monkeypatch):
    """The writer / reader split their input over host threads (PN_HOST_THREADS): same bytes, same arrays, and the
    earliest malformed byte is the one reported, whatever the split."""
    rng = np.random.default_rng(11)
    W, L = 120_000, 4                                     # ~3.5 MB of text: several reader pieces (>= 1 MiB each)
    nodes = rng.integers(0, 20_000_000, (W, L)).astype(np.int32)
    types = rng.integers(0, 9, (W, L)).astype(np.uint8)
    nodes[0, 0], nodes[1, 1], nodes[2, 2], nodes[3, 3] = 0, 2147483647, -2147483648, -5   # digit-count edges, signs
    types[4, 0], types[5, 1], types[6, 2] = 255, 100, 10
    want = mo.format_txt(nodes[:3000], types[:3000])
    outs = []
    for t in ("1", "3", "16"):
        monkeypatch.setenv("PN_HOST_THREADS", t)
        txt = sampler.format_paths_txt(nodes, types)
        assert txt[:len(want)] == want
        outs.append(txt)
        n2, t2 = sampler.parse_paths_txt(txt, L)
        assert n2.shape == (W, L) and (n2 == nodes).all() and (t2 == types).all()
    assert outs[0] == outs[1] == outs[2]
    size = _lib.lib().pn_paths_format_txt(nodes.ctypes.data, types.ctypes.data, W, L, None, 0)   # size query
    assert size == len(outs[0])
    bad = bytearray(outs[0])
    pos = len(bad) * 2 // 3
    pos = bad.index(b"[", pos) + 1
    bad[pos:pos + 1] = b"x"                                # a non-digit where a field must start, in a late piece
    for t in ("1", "5"):
        monkeypatch.setenv("PN_HOST_THREADS", t)
        with pytest.raises(_lib.PnError, match=f"expected digit at byte {pos}"):
            sampler.parse_paths_txt(bytes(bad), L)
    with pytest.raises(_lib.PnError, match="more than"):  # output arrays too small
        n_small = np.zeros((10, L), np.int32); t_small = np.zeros((10, L), np.uint8)
        _lib.check(_lib.lib().pn_paths_parse_txt(outs[0], len(outs[0]), L, n_small.ctypes.data, t_small.ctypes.data, 10))


def test_edge_file_reader_matches_fscanf_semantics(tmp_path, monkeypatch):
    """pn_edges_parse_txt: header + m rows, doubles bit-identical to strtod, rows past m ignored, blank lines / CRLF /
    tabs accepted, rows that span lines read as a token stream (fscanf), short files rejected; thread-count invariant."""
    rng = np.random.default_rng(5)
    m = 60_000                                            # ~2.5 MB: several pieces
    u = rng.integers(0, 99_999, m).astype(np.int32)
    v = rng.integers(0, 99_999, m).astype(np.int32)
    p = rng.random(m) * 10.0 ** rng.integers(-12, 3, m)
    p[:4] = [0.0, 1.0, 5e-324, 1.7976931348623157e308]
    lines = [f"{a} {b} {repr(float(c))}" for a, b, c in zip(u, v, p)]
    lines[7] = f"{u[7]}\t{v[7]}  {repr(float(p[7]))}  "    # tabs, trailing blanks
    lines[9] = lines[9] + "\r"                            # CRLF
    lines[11] = "\n" + lines[11]                          # blank line
    lines[13] = f"+{u[13]} {v[13]} {float(p[13]):.17e}"    # explicit sign, exponent form
    f = tmp_path / "g.in"
    f.write_text("100000 %d\n" % m + "\n".join(lines) + "\n" + "1 2 0.5\n" * 3)   # 3 rows beyond m: ignored
    want_p = np.array([float(x.split()[2]) for x in lines])                        # Python float() == strtod
    for t in ("1", "4"):
        monkeypatch.setenv("PN_HOST_THREADS", t)
        n, uu, vv, pp = sampler.read_edge_input(f)
        assert n == 100000 and len(uu) == m
        assert (uu == u).all() and (vv == v).all()
        assert (pp.view(np.uint64) == want_p.view(np.uint64)).all()
    # token stream: a row broken over two lines still reads (fscanf does not care about line ends)
    g = tmp_path / "h.in"
    g.write_text("5 3\n0 1\n0.25\n1 2 0.5 2\n3 0.125\n")
    n, uu, vv, pp = sampler.read_edge_input(g)
    assert n == 5 and uu.tolist() == [0, 1, 2] and vv.tolist() == [1, 2, 3] and pp.tolist() == [0.25, 0.5, 0.125]
    # 2-column variant
    h = tmp_path / "k_nsl.in"
    h.write_text("4 3\n0 1\n1 2\n2 3\n")
    n, uu, vv = sampler.read_edge_input_nsl(h)
    assert n == 4 and uu.tolist() == [0, 1, 2] and vv.tolist() == [1, 2, 3]
    short = tmp_path / "s.in"
    short.write_text("4 3\n0 1 0.5\n1 2 0.5\n")
    with pytest.raises(ValueError, match="header says 3 rows, found 2"):
        sampler.read_edge_input(short)


def test_abi_exports_every_declared_symbol():
    import re
    from pathlib import Path
    hdr = (Path(__file__).resolve().parents[1] / "include" / "pathnet_b200.h").read_text()
    declared = set(re.findall(r"\b(pn_[a-z0-9_]+)\s*\(", hdr))
    L = _lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib.EXPORTS) <= declared
    assert L.pn_version() == 200
