"""Generate the committed sampler fixtures from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference and
`make -C oracle ref`):

    python tests/golden/make_golden.py [--skip-bgp]

Outputs (all small, committed):
  tests/golden/graph_cornell.npz, graph_cora.npz   edge rows of edge_input/*.in (u, v, p, n)
  tests/golden/merw_ref_hashes.json                sha256 of the reference binary's text
                                                   for (graph, nw, L, seed, first E epochs)
  tests/golden/merw_cornell_40_4_s12345_e0.npz     epoch 0 of the reference's output, parsed
The reference binary is oracle/_ref/gen_merw / gen_epoch_merw run under
oracle/faketime.c (FAKE_TIME=seed -> srand(seed)).
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import merw_oracle as mo  # noqa: E402

EDGE = Path("/root/reference/edge_input")
OUT = Path(__file__).resolve().parent

CASES = [
    # name, nw, L, seed, epochs, per_epoch binary?
    ("cornell", 40, 4, 12345, 2, False),
    ("cornell", 40, 4, 1, 1, False),
    ("cornell", 40, 6, 777, 1, False),
    ("cornell", 7, 3, 99, 2, False),
    ("cora", 40, 4, 12345, 1, False),
    ("cora", 40, 4, 1, 1, False),
    ("citeseer", 40, 4, 5, 1, False),
    ("pubmed", 40, 4, 777, 2, True),
    ("Nba", 10, 5, 3, 1, False),
    ("Electronics", 8, 4, 11, 1, True),
    ("bgp", 40, 6, 2024, 1, True),
]


def parse_txt(txt: bytes, L: int):
    rows = np.array([[int(x) for x in ln[1:-1].split(b",")] for ln in txt.splitlines()],
                    dtype=np.int32)
    return rows[:, :L].copy(), rows[:, L:].astype(np.uint8)


def main():
    skip_bgp = "--skip-bgp" in sys.argv
    assert mo.build_ref(), "reference sampler could not be built"
    for name in ("cornell", "cora"):
        n, u, v, p = mo.read_edge_input(EDGE / f"{name}.in")
        np.savez_compressed(OUT / f"graph_{name}.npz", n=n, u=u, v=v, p=p)
    hpath = OUT / "merw_ref_hashes.json"
    hashes = json.loads(hpath.read_text()) if hpath.exists() else {}
    for name, nw, L, seed, epochs, per_epoch in CASES:
        key = f"{name}|nw={nw}|L={L}|seed={seed}|epochs={epochs}"
        if key in hashes or (skip_bgp and name == "bgp"):
            continue
        txt = mo.run_reference(EDGE / f"{name}.in", nw, L, seed, epochs=epochs,
                               per_epoch=per_epoch, name=name)
        hashes[key] = {"sha256": hashlib.sha256(txt).hexdigest(), "bytes": len(txt),
                       "binary": "gen_epoch_merw" if per_epoch else "gen_merw"}
        print(key, hashes[key], flush=True)
        hpath.write_text(json.dumps(hashes, indent=1, sort_keys=True) + "\n")
        if (name, nw, L, seed) == ("cornell", 40, 4, 12345):
            n = 183
            nodes, types = parse_txt(txt[: txt.find(b"\n", 0) + 1] and txt, L)
            nodes = nodes[: n * nw].reshape(n, nw, L)
            types = types[: n * nw].reshape(n, nw, L)
            np.savez_compressed(OUT / "merw_cornell_40_4_s12345_e0.npz", nodes=nodes, types=types)


if __name__ == "__main__":
    main()
