"""MERW path sampler — host-side mirror of the reference's sampler surface.

The reference has no Python function for this step: its surface is the CLI
`./gen_merw <name> <path_num> <path_len>` plus the text file it writes
(preprocess/gen_merw.cpp:125-213).  `gen_merw()` below is the in-memory
equivalent (tensors straight to PAGG), `write_paths_txt()` produces the
byte-identical text, and pathnet_b200/cli.py mirrors the CLI itself.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from pathlib import Path

import numpy as np
import torch

from ._lib import check, lib


def _read_edges(path, cols: int):
    """(n, u int32[m], v int32[m], p float64[m] | None) through pn_edges_parse_txt (all host threads; doubles rounded
    like the reference's fscanf("%lf"))."""
    txt = np.fromfile(path, dtype=np.uint8)
    n, m = C.c_int32(0), C.c_int64(0)
    check(lib().pn_edges_parse_txt(txt.ctypes.data, txt.size, cols, 0, C.byref(n), C.byref(m), None, None, None))
    m_ = int(m.value)
    u, v = np.empty(m_, dtype=np.int32), np.empty(m_, dtype=np.int32)
    p = np.empty(m_, dtype=np.float64) if cols == 3 else None
    try:
        check(lib().pn_edges_parse_txt(txt.ctypes.data, txt.size, cols, m_, C.byref(n), C.byref(m), u.ctypes.data,
                                       v.ctypes.data, p.ctypes.data if p is not None else None))
    except Exception as ex:   # same exception type as before for short files
        raise ValueError(f"{path}: {ex}") from None
    return int(n.value), u, v, p


def read_edge_input(path):
    """`<name>.in`: header `n m`, then m rows `u v p` (gen_merw.cpp:162-172)."""
    n, u, v, p = _read_edges(path, 3)
    return n, u, v, p


def read_edge_input_nsl(path):
    """2-column `<name>_nsl.in` of the uniform-RW sampler: header `n m`, then m rows `u v` (gen.cpp:79-94)."""
    n, u, v, _ = _read_edges(path, 2)
    return n, u, v


class MerwGraph:
    """CSR + alias cells of one `<name>.in`, resident on one GPU."""

    def __init__(self, n, u, v, p=None, device="cuda:0"):
        """p=None: the uniform-random-walk variant of the sampler (preprocess/gen.cpp) on 2-column rows."""
        u = np.ascontiguousarray(u, dtype=np.int32)
        v = np.ascontiguousarray(v, dtype=np.int32)
        self.device = None if device in (None, "host") else torch.device(device)
        dev_index = -1 if self.device is None else (self.device.index or 0)
        h = C.c_void_p()
        if p is None:
            if len(u) != len(v):
                raise ValueError("u, v must have the same length")
            check(lib().pn_graph_from_edges_uniform(int(n), len(u), u.ctypes.data, v.ctypes.data, dev_index, C.byref(h)))
            self._h = h
            self.n, self.m = int(n), self.info()["m"]
            return
        p = np.ascontiguousarray(p, dtype=np.float64)
        if not (len(u) == len(v) == len(p)):
            raise ValueError("u, v, p must have the same length")
        check(lib().pn_graph_from_edges(int(n), len(u), u.ctypes.data, v.ctypes.data, p.ctypes.data,
                                        dev_index, C.byref(h)))
        self._h = h
        self.n, self.m = int(n), len(u)

    @classmethod
    def from_file(cls, path, device="cuda:0"):
        return cls(*read_edge_input(path), device=device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().pn_graph_free(h)
            except Exception:   # interpreter shutdown
                pass
            self._h = None

    def info(self):
        n, m, md, de = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32()
        check(lib().pn_graph_info(self._h, C.byref(n), C.byref(m), C.byref(md), C.byref(de)))
        return dict(n=n.value, m=m.value, max_deg=md.value, dead_end=de.value)

    def cells_host(self):
        off = np.zeros(self.n + 1, dtype=np.int64)
        A = np.zeros(self.m, dtype=np.int32)
        B = np.zeros(self.m, dtype=np.int32)
        S = np.zeros(self.m, dtype=np.float64)
        thr = np.zeros(self.m, dtype=np.int32)
        check(lib().pn_graph_cells_host(self._h, off.ctypes.data, A.ctypes.data, B.ctypes.data,
                                        S.ctypes.data, thr.ctypes.data))
        return off, A, B, S, thr

    def prepare_labels(self, L, max_bytes=0):
        with torch.cuda.device(self.device):
            check(lib().pn_graph_prepare_labels(self._h, int(L), int(max_bytes),
                                                torch.cuda.current_stream().cuda_stream))

    def sample(self, path_num, path_len, epochs=1, seed=12345, epoch_begin=0, node_begin=0,
               node_end=None, labels=True, out=None, rng="glibc"):
        """-> (nodes int32 [E, n', nw, L], types uint8 [E, n', nw, L]) on the GPU.
        rng="glibc" (default): bit-identical to what the reference prints for srand(seed).
        rng="philox": counter-based Philox4x32-10, one stream per walk (include/pathnet_b200.h: PN_RNG_PHILOX) -- the same
        walk law, independent of how epochs / start nodes are split over calls or GPUs; 64-bit seed."""
        mode = {"glibc": 0, "philox": 1}[rng]
        if self.device is None:
            raise RuntimeError("host-only graph: no device to sample on")
        node_end = self.n if node_end is None else node_end
        E, npr = int(epochs), node_end - node_begin
        shape = (E, npr, path_num, path_len)
        with torch.cuda.device(self.device):
            if out is None:
                nodes = torch.empty(shape, dtype=torch.int32, device=self.device)
                types = torch.empty(shape, dtype=torch.uint8, device=self.device) if labels else None
            else:
                nodes, types = out
            check(lib().pn_merw_sample(self._h, path_num, path_len, epoch_begin, epoch_begin + E,
                                       node_begin, node_end, mode, int(seed) & (0xFFFFFFFF if mode == 0 else 0xFFFFFFFFFFFFFFFF),
                                       nodes.data_ptr(), types.data_ptr() if types is not None else None,
                                       torch.cuda.current_stream().cuda_stream))
        return nodes, types


_GRAPH_CACHE: dict = {}


def gen_merw(name_or_edges, path_num, path_len, epochs=1000, seed=None, device="cuda:0",
             edge_root="../edge_input", epoch_begin=0, rng="glibc"):
    """Drop-in for `./gen_merw <name> <path_num> <path_len>` that hands tensors over.

    name_or_edges: dataset name (reads `<edge_root>/<name>.in` like
    gen_merw.cpp:134-139), a path to a `.in` file, a MerwGraph, or (n, u, v, p).
    seed=None mirrors `srand(time(0))` (gen_merw.cpp:161).
    Returns (neis int64 [E, n, nw*L], path_type int64 [E, n, nw, L]) on `device`
    — the tensors PathNet_run.py builds from the text file (:309-313) —
    """
    if seed is None:
        seed = int(time.time())
    if isinstance(name_or_edges, MerwGraph):
        g = name_or_edges
    elif isinstance(name_or_edges, (tuple, list)):
        g = MerwGraph(*name_or_edges, device=device)
    else:
        p = Path(name_or_edges)
        if not p.exists():
            p = Path(edge_root) / f"{name_or_edges}.in"
        key = (str(p.resolve()), str(device))
        g = _GRAPH_CACHE.get(key)
        if g is None:
            g = _GRAPH_CACHE[key] = MerwGraph.from_file(p, device=device)
    nodes, types = g.sample(path_num, path_len, epochs=epochs, seed=seed, epoch_begin=epoch_begin, rng=rng)
    E, n = nodes.shape[0], nodes.shape[1]
    return nodes.view(E, n, path_num * path_len).long(), types.long()


def format_paths_txt(nodes, types) -> bytes:
    """Text of gen_merw.cpp:188-206 for walks in (..., L) arrays."""
    nodes = np.ascontiguousarray(torch.as_tensor(nodes).cpu().numpy(), dtype=np.int32)
    types = np.ascontiguousarray(torch.as_tensor(types).cpu().numpy(), dtype=np.uint8)
    return format_paths_array(nodes, types).tobytes()


def format_paths_array(nodes: np.ndarray, types: np.ndarray) -> np.ndarray:
    """The text as a uint8 array (no bytes copy; sized exactly by a first, output-less call; both passes use all host
    threads).  `nodes` int32 / `types` uint8 C-contiguous numpy arrays of shape (..., L)."""
    nodes = np.ascontiguousarray(nodes, dtype=np.int32)   # no copies when already in that form
    types = np.ascontiguousarray(types, dtype=np.uint8)
    L = nodes.shape[-1]
    walks = nodes.size // L
    size = check(lib().pn_paths_format_txt(nodes.ctypes.data, types.ctypes.data, walks, L, None, 0))
    buf = np.empty(size, dtype=np.uint8)
    k = check(lib().pn_paths_format_txt(nodes.ctypes.data, types.ctypes.data, walks, L, buf.ctypes.data, size))
    assert k == size
    return buf


def write_paths_txt(path, nodes, types) -> int:
    """Write walks in the reference's text format; returns bytes written."""
    nodes = np.ascontiguousarray(torch.as_tensor(nodes).cpu().numpy(), dtype=np.int32)
    types = np.ascontiguousarray(torch.as_tensor(types).cpu().numpy(), dtype=np.uint8)
    buf = format_paths_array(nodes, types)
    with open(path, "wb") as f:
        f.write(memoryview(buf))
    return int(buf.size)


def parse_paths_txt(txt: bytes, L: int):
    """Inverse of format_paths_txt (reader: PathNet_run.py:325-334)."""
    max_walks = txt.count(b"\n") + 1
    nodes = np.zeros((max_walks, L), dtype=np.int32)
    types = np.zeros((max_walks, L), dtype=np.uint8)
    k = check(lib().pn_paths_parse_txt(txt, len(txt), L, nodes.ctypes.data, types.ctypes.data,
                                       max_walks))
    return nodes[:k], types[:k]


def read_paths_txt(path, L: int):
    with open(path, "rb") as f:
        return parse_paths_txt(f.read(), L)
