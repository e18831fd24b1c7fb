"""Build recipe of libpathnet_b200.so (in-tree, sm_100a only).

    python -m pathnet_b200.build [--force]

nvcc cross-compiles without a GPU; the built library lives next to this file
so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
LIB = PKG / "libpathnet_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fno-fast-math,-ffp-contract=off",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def _stale(obj: Path, src: Path) -> bool:
    if not obj.exists():
        return True
    newest = max([src.stat().st_mtime] + [h.stat().st_mtime for h in CSRC.glob("*.h")] +
                 [h.stat().st_mtime for h in CSRC.glob("*.cuh")] +
                 [(PKG.parent / "include" / "pathnet_b200.h").stat().st_mtime])
    return obj.stat().st_mtime < newest


def _compile(src: Path, force: bool) -> tuple[Path, str]:
    obj = BUILD / (src.name + ".o")
    if not force and not _stale(obj, src):
        return obj, ""
    cmd = [NVCC, *ARCH, *COMMON, "-x", "cu", "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    (BUILD / (src.name + ".ptxas.log")).write_text(r.stderr)
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or not LIB.exists() or any(o.stat().st_mtime > LIB.stat().st_mtime for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart", "-lcuda", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
