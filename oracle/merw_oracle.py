"""TEST INFRASTRUCTURE ONLY — ctypes front-end of oracle/merw_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module (see the header of merw_oracle.c).  It also knows how to
drive the UNMODIFIED reference binary (oracle/_ref/gen_merw, built by
`make -C oracle ref`) under the time() shim so tests can pin the restatement
against the real thing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = HERE / "libmerw_oracle.so"
    src = HERE / "merw_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(HERE), "libmerw_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def build_ref() -> bool:
    """Compile the reference sampler where it lies (authoring container only)."""
    if not Path("/root/reference/preprocess/gen_merw.cpp").exists():
        return have_ref()
    subprocess.check_call(["make", "-C", str(HERE), "ref"], stdout=subprocess.DEVNULL)
    from oracle import stage_reference
    stage_reference.stage()   # the reference's Python driver + cornell data, for the reference arm / launcher test on the GPU box
    return have_ref()


def have_ref() -> bool:
    return all((HERE / "_ref" / f).exists() for f in ("gen_merw", "gen_epoch_merw", "faketime.so"))


def have_ref_uniform() -> bool:
    return all((HERE / "_ref" / f).exists() for f in ("gen", "faketime.so"))


def read_edge_input_nsl(path):
    """2-column `<name>_nsl.in`: header `n m`, then m rows `u v` (gen.cpp:79-94)."""
    with open(path) as f:
        n, m = (int(x) for x in f.readline().split()[:2])
        arr = np.loadtxt(f, dtype=np.int64, ndmin=2, max_rows=m) if m else np.zeros((0, 2), dtype=np.int64)
    return n, arr[:, 0].astype(np.int32), arr[:, 1].astype(np.int32)


def run_reference_uniform(edge_file, nw, L, seed, epochs=1, name="g", timeout=3600) -> bytes:
    """Run the unmodified uniform-RW binary (oracle/_ref/gen from preprocess/gen.cpp): reads
    ../edge_input/<name>_nsl.in, writes ./<name>_<nw>_<L>_nsl.txt (all 1000 epochs)."""
    n = int(open(edge_file).readline().split()[0])
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as W:
        os.makedirs(f"{W}/preprocess"); os.makedirs(f"{W}/edge_input")
        os.symlink(os.path.abspath(edge_file), f"{W}/edge_input/{name}_nsl.in")
        env = dict(os.environ, FAKE_TIME=str(seed), LD_PRELOAD=str(HERE / "_ref" / "faketime.so"))
        subprocess.run([str(HERE / "_ref" / "gen"), name, str(nw), str(L)], cwd=f"{W}/preprocess", env=env, check=True,
                       timeout=timeout, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        out = bytearray()
        with open(f"{W}/preprocess/{name}_{nw}_{L}_nsl.txt", "rb") as f:
            for _ in range(n * nw * epochs):
                out += f.readline()
        return bytes(out)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        L.merw_oracle_graph.restype = C.c_void_p
        L.merw_oracle_graph.argtypes = [C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
        L.merw_oracle_graph_uniform.restype = C.c_void_p
        L.merw_oracle_graph_uniform.argtypes = [C.c_int, C.c_long, C.c_void_p, C.c_void_p]
        L.merw_oracle_graph_free.argtypes = [C.c_void_p]
        L.merw_oracle_graph_cells.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.merw_oracle_sample.restype = C.c_int
        L.merw_oracle_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_long,
                                         C.c_long, C.c_void_p, C.c_void_p]
        L.merw_oracle_sample_range.restype = C.c_int
        L.merw_oracle_sample_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_long, C.c_long,
                                               C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.merw_oracle_format.restype = C.c_long
        L.merw_oracle_format.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p]
        L.merw_oracle_rand_stream.argtypes = [C.c_uint, C.c_long, C.c_void_p]
        _LIB = L
    return _LIB


def read_edge_input(path) -> tuple[int, np.ndarray, np.ndarray, np.ndarray]:
    """`<name>.in`: header `n m`, then m rows `u v p` (gen_merw.cpp:162-172)."""
    with open(path) as f:
        n, m = (int(x) for x in f.readline().split())
        arr = np.loadtxt(f, dtype=np.float64, ndmin=2) if m else np.zeros((0, 3))
    arr = arr[:m]
    return n, arr[:, 0].astype(np.int32), arr[:, 1].astype(np.int32), np.ascontiguousarray(arr[:, 2])


def write_edge_input(path, n, u, v, p) -> None:
    with open(path, "w") as f:
        f.write(f"{n} {len(u)}\n")
        for a, b, c in zip(u.tolist(), v.tolist(), p.tolist()):
            f.write(f"{a} {b} {c!r}\n")


class OracleGraph:
    def __init__(self, n, u, v, p=None):
        """p=None: the uniform-random-walk variant (preprocess/gen.cpp) on 2-column rows."""
        self.n, self.m = int(n), int(len(u))
        self._u = np.ascontiguousarray(u, dtype=np.int32)
        self._v = np.ascontiguousarray(v, dtype=np.int32)
        if p is None:
            self._h = lib().merw_oracle_graph_uniform(self.n, self.m, self._u.ctypes.data, self._v.ctypes.data)
            self.m = self.n + 2 * int((self._u != self._v).sum())   # rows of gen.cpp's adjacency
            return
        self._p = np.ascontiguousarray(p, dtype=np.float64)
        self._h = lib().merw_oracle_graph(self.n, self.m, self._u.ctypes.data,
                                          self._v.ctypes.data, self._p.ctypes.data)

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().merw_oracle_graph_free(self._h)
            except Exception:   # interpreter shutdown
                pass
            self._h = None

    def cells(self):
        off = np.zeros(self.n + 1, dtype=np.int64)
        A = np.zeros(self.m, dtype=np.int32)
        B = np.zeros(self.m, dtype=np.int32)
        S = np.zeros(self.m, dtype=np.float64)
        lib().merw_oracle_graph_cells(self._h, off.ctypes.data, A.ctypes.data, B.ctypes.data,
                                      S.ctypes.data)
        return off, A, B, S

    def sample(self, nw, L, seed, epoch_begin=0, epoch_end=1, labels=True, node_begin=0, node_end=None):
        node_end = self.n if node_end is None else node_end
        E = epoch_end - epoch_begin
        nodes = np.zeros((E, node_end - node_begin, nw, L), dtype=np.int32)
        types = np.zeros((E, node_end - node_begin, nw, L), dtype=np.uint8) if labels else None
        rc = lib().merw_oracle_sample_range(self._h, nw, L, seed, epoch_begin, epoch_end, node_begin, node_end,
                                            nodes.ctypes.data, types.ctypes.data if labels else None)
        if rc != 0:
            raise RuntimeError("dead-end node rolled (reference would exit(0))")
        return nodes, types


    def sample_philox(self, nw, L, seed, epoch_begin=0, epoch_end=1, labels=True, node_begin=0, node_end=None):
        """rng_mode = philox (see merw_oracle.c): every walk draws from its own Philox4x32-10 counter stream."""
        node_end = self.n if node_end is None else node_end
        E = epoch_end - epoch_begin
        nodes = np.zeros((E, node_end - node_begin, nw, L), dtype=np.int32)
        types = np.zeros((E, node_end - node_begin, nw, L), dtype=np.uint8) if labels else None
        fn = lib().merw_oracle_sample_philox
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_ulonglong, C.c_long, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        rc = fn(self._h, nw, L, seed, epoch_begin, epoch_end, node_begin, node_end,
                nodes.ctypes.data, types.ctypes.data if labels else None)
        if rc != 0:
            raise RuntimeError("dead-end node rolled")
        return nodes, types


def philox4x32_10(ctr, key) -> list[int]:
    """One Philox4x32-10 block through the C restatement (known-answer tests)."""
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    fn = lib().merw_oracle_philox4x32_10
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    fn(c, k, o)
    return [int(x) for x in o]


def format_txt(nodes: np.ndarray, types: np.ndarray) -> bytes:
    L = nodes.shape[-1]
    nodes = np.ascontiguousarray(nodes.reshape(-1, L), dtype=np.int32)
    types = np.ascontiguousarray(types.reshape(-1, L), dtype=np.uint8)
    walks = nodes.shape[0]
    need = lib().merw_oracle_format(nodes.ctypes.data, types.ctypes.data, walks, L, None)
    buf = C.create_string_buffer(need)
    lib().merw_oracle_format(nodes.ctypes.data, types.ctypes.data, walks, L, buf)
    return buf.raw[:need]


def rand_stream(seed: int, count: int) -> np.ndarray:
    out = np.zeros(count, dtype=np.int32)
    lib().merw_oracle_rand_stream(seed, count, out.ctypes.data)
    return out


def _srand_window(seed: int):
    """s_0..s_30 = (r_3..r_33) of glibc srand(seed) (see merw_oracle.c)."""
    seed = 1 if seed == 0 else seed & 0xFFFFFFFF
    word = seed - (1 << 32) if seed >= (1 << 31) else seed
    r = [word & 0xFFFFFFFF]
    for _ in range(30):
        hi, lo = int(word / 127773), word - int(word / 127773) * 127773  # C division truncates
        word = 16807 * lo - 2836 * hi
        word = ((word + (1 << 31)) % (1 << 32)) - (1 << 31)  # int32 wrap
        if word < 0:
            word += 2147483647
        r.append(word & 0xFFFFFFFF)
    r += r[0:3]
    return r[3:34]


def _poly_mul(a, b):
    t = [0] * 61
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] = (t[i + j] + ai * bj) & 0xFFFFFFFF
    for d in range(60, 30, -1):
        t[d - 3] = (t[d - 3] + t[d]) & 0xFFFFFFFF
        t[d - 31] = (t[d - 31] + t[d]) & 0xFFFFFFFF
    return t[:31]


def rand_stream_at(seed: int, offset: int, count: int) -> list[int]:
    """rand() outputs [offset, offset+count) after srand(seed), by an independent
    pure-Python jump (x^k mod x^31-x^28-1); pinned against the sequential C
    stream in tests/test_oracle_pinning.py."""
    k = 310 + offset
    result, base = [1] + [0] * 30, [0, 1] + [0] * 29
    while k:
        if k & 1:
            result = _poly_mul(result, base)
        k >>= 1
        if k:
            base = _poly_mul(base, base)
    s = _srand_window(seed)
    e = list(s)
    for d in range(30):
        e.append((e[d] + e[d + 28]) & 0xFFFFFFFF)
    w = [sum(result[i] * e[i + j] for i in range(31)) & 0xFFFFFFFF for j in range(31)]
    out = []
    for _ in range(count):
        v = (w[0] + w[28]) & 0xFFFFFFFF
        w = w[1:] + [v]
        out.append(v >> 1)
    return out


def run_reference(edge_file, nw, L, seed, epochs=1, per_epoch=False, name="g",
                  timeout=3600) -> bytes:
    """Run the unmodified reference binary on `edge_file` with srand(seed) and
    return the text of the first `epochs` epochs.  Lays out the cwd the binary
    expects (../edge_input/<name>.in, output in ./)."""
    if not have_ref():
        raise FileNotFoundError("oracle/_ref not built (make -C oracle ref)")
    n = int(open(edge_file).readline().split()[0])
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as W:
        os.makedirs(f"{W}/preprocess")
        os.makedirs(f"{W}/edge_input")
        os.symlink(os.path.abspath(edge_file), f"{W}/edge_input/{name}.in")
        env = dict(os.environ, FAKE_TIME=str(seed), LD_PRELOAD=str(HERE / "_ref" / "faketime.so"))
        lines = n * nw * epochs
        if not per_epoch:
            subprocess.run([str(HERE / "_ref" / "gen_merw"), name, str(nw), str(L)],
                           cwd=f"{W}/preprocess", env=env, check=True, timeout=timeout,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out = bytearray()
            with open(f"{W}/preprocess/{name}_{nw}_{L}_merw.txt", "rb") as f:
                for _ in range(lines):
                    out += f.readline()
            return bytes(out)
        # per-epoch variant writes 1000 files; stop it once epoch `epochs` exists
        import time as _t
        p = subprocess.Popen([str(HERE / "_ref" / "gen_epoch_merw"), name, str(nw), str(L)],
                             cwd=f"{W}/preprocess", env=env, stdout=subprocess.DEVNULL,
                             stderr=subprocess.DEVNULL)
        try:
            t0 = _t.time()
            nxt = f"{W}/preprocess/{name}_{nw}_{L}_{epochs}_merw.txt"
            while not os.path.exists(nxt) and p.poll() is None:
                if _t.time() - t0 > timeout:
                    raise TimeoutError("reference sampler timed out")
                _t.sleep(0.05)
        finally:
            p.kill()
            p.wait()
        out = bytearray()
        for e in range(epochs):
            out += open(f"{W}/preprocess/{name}_{nw}_{L}_{e}_merw.txt", "rb").read()
        return bytes(out)
