"""TEST / BASELINE INFRASTRUCTURE ONLY — stages what is needed to RUN the unmodified reference on the GPU box.

/root/reference does not exist on the GPU box; `oracle/_ref/` (git-ignored, so nothing of it enters history, but NOT
gpurun-ignored, so it travels like the compiled reference sampler beside it) is where reference-derived artefacts
live.  This script copies, unmodified and only in the authoring container:

    oracle/_ref/pathnet/PathNet_run.py, data_loader.py, dataset.py, dataset.json      the reference's driver script
    oracle/_ref/pathnet/splits/<cornell ...>/                                         its cornell features / labels / splits
    oracle/_ref/pathnet/edge_input/cornell.in                                         sampler input
    oracle/_ref/pathnet/{preprocess,results,saved_models}/                            directories the script writes into

Users: bench.py's reference arm and `gpu_status_quo` leg (the AST-extracted PathNet / PathNet_homo classes, run as
they are), and tests/test_gpu_launcher.py (the whole script through `python -m pathnet_b200.run`).  The product
(pathnet_b200/) never reads any of it.
"""
from __future__ import annotations

import io
import json
import shutil
import zipfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
DST = HERE / "_ref" / "pathnet"
ROUNDS = 3          # cornell split directories staged (the script's -r)


def staged() -> bool:
    return (DST / "PathNet_run.py").exists()


def stage(force: bool = False) -> bool:
    if not (REF / "PathNet_run.py").exists():
        return staged()
    if staged() and not force:
        return True
    DST.mkdir(parents=True, exist_ok=True)
    for f in ("PathNet_run.py", "data_loader.py", "dataset.py", "dataset.json"):
        shutil.copyfile(REF / f, DST / f)
    for d in ("preprocess", "results", "saved_models", "edge_input"):
        (DST / d).mkdir(exist_ok=True)
    shutil.copyfile(REF / "edge_input" / "cornell.in", DST / "edge_input" / "cornell.in")
    cfg = json.loads((REF / "dataset.json").read_text())["cornell"]
    want = [p[2:] if p.startswith("./") else p for p in cfg["dataset_path"][:ROUNDS]]   # splits/<ds>/features/<f>/splits/<s>
    ds = want[0].split("/")[1]
    outer = zipfile.ZipFile(REF / "splits" / "splits.zip")
    inner = zipfile.ZipFile(io.BytesIO(outer.read(f"{ds}.zip")))
    for name in inner.namelist():
        full = f"splits/{ds}/{name}"
        if name.endswith("/") or not any(full.startswith(w + "/") for w in want) or "/split_source/" in full:
            continue
        out = DST / full
        out.parent.mkdir(parents=True, exist_ok=True)
        out.write_bytes(inner.read(name))
    return True


if __name__ == "__main__":
    print("staged" if stage(force=True) else "reference not mounted; nothing staged", DST)
