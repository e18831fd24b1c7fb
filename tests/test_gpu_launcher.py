"""GPU tests of the 'run the script unchanged' path (pathnet_b200.run) and of
the CLI shims (pathnet_b200.cli)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import merw_oracle as mo
from oracle import pagg_oracle as po
from tests.helpers import load_graph, ref_hashes, sha256

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_launcher_runs_a_pathnet_run_shaped_script(tmp_path):
    out = tmp_path / "out.pt"
    r = subprocess.run([sys.executable, "-m", "pathnet_b200.run", str(ROOT / "tests" / "fake_pathnet_run.py"),
                        "cuda:0", str(out)], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = torch.load(out, weights_only=False)
    N, F, H, C, nw, L, S = d["dims"]
    for name, variant in (("PathNet", "hetero"), ("PathNet_homo", "homo")):
        got = d["res"][name]
        params = {k: v.clone().double().requires_grad_(True) for k, v in got["sd"].items()}
        y = po.pagg_forward(variant, params, d["X"].double(), d["neis"], nw, L, torch.tensor(d["mask"]), d["ltype"])
        y.square().mean().backward()
        torch.testing.assert_close(got["y"], y.detach().float(), rtol=1e-4, atol=2e-5)
        for k, g in got["g"].items():
            torch.testing.assert_close(g, params[k].grad.float(), rtol=2e-3, atol=2e-5, msg=lambda s: f"{k}: {s}")


def test_cli_shims_write_the_reference_files(tmp_path):
    (tmp_path / "edge_input").mkdir()
    (tmp_path / "preprocess").mkdir()
    n, u, v, p = load_graph("cornell")
    mo.write_edge_input(tmp_path / "edge_input" / "cornell.in", n, u, v, p)
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, "-m", "pathnet_b200.cli", "gen_merw", "cornell", "40", "4", "--seed", "12345",
                        "--epochs", "2"], cwd=tmp_path / "preprocess", env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "File input: ../edge_input/cornell.in" in r.stdout and "cornell: 183" in r.stderr
    txt = (tmp_path / "preprocess" / "cornell_40_4_merw.txt").read_bytes()
    want = ref_hashes()["cornell|nw=40|L=4|seed=12345|epochs=2"]
    assert len(txt) == want["bytes"] and sha256(txt) == want["sha256"]      # byte-identical to the reference binary
    r = subprocess.run([sys.executable, "-m", "pathnet_b200.cli", "gen_epoch_merw", "cornell", "40", "4", "--seed",
                        "12345", "--epochs", "2"], cwd=tmp_path / "preprocess", env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    per_epoch = b"".join((tmp_path / "preprocess" / f"cornell_40_4_{e}_merw.txt").read_bytes() for e in range(2))
    assert per_epoch == txt
