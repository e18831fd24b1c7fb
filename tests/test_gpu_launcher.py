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


STAGED = ROOT / "oracle" / "_ref" / "pathnet"


@pytest.mark.skipif(not (STAGED / "PathNet_run.py").exists(), reason="reference script not staged (oracle/stage_reference.py)")
def test_real_reference_script_trains_cornell_on_the_gpu(tmp_path):
    """The UNMODIFIED reference script (PathNet_run.py:406-485: argparse, PlanetoidData loading, the 7.32 M-line path
    file parse, its own constructors, Adam, train/val/test loop, torch.save, results file) runs end to end on CUDA
    through `python -m pathnet_b200.run`, with path files written by the CLI shim of the sampler — the reference's
    whole workflow with both hot-path stages replaced.  cornell, hid 128, the script's default dropout 0.7."""
    import re
    import shutil
    work = tmp_path / "pathnet"
    shutil.copytree(STAGED, work)
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, "-m", "pathnet_b200.cli", "gen_merw", "cornell", "40", "4", "--seed", "2022"],
                       cwd=work / "preprocess", env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert (work / "preprocess" / "cornell_40_4_merw.txt").stat().st_size > 200e6      # 1000 epochs x 183 x 40 walks
    epochs = 150
    r = subprocess.run([sys.executable, "-m", "pathnet_b200.run", "PathNet_run.py", "-data=cornell", "-hid=128", "-c", "cuda:0",
                        "-e", str(epochs), "-r", "1", "-pr", "./preprocess/"], cwd=work, env=env, capture_output=True,
                       text=True, timeout=1800)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    txt = (work / "results" / "result_for_cornell.txt").read_text()
    m = re.search(r"acc([0-9.]+) ", txt)
    assert m, txt[-500:]
    acc = float(m.group(1))
    print(f"reference script through pathnet_b200.run: cornell test acc {acc:.4f} after {epochs} epochs "
          f"(reference: 0.92 after 1000 epochs x 10 rounds, results/result_for_cornell.txt)")
    assert any(p.suffix == ".pth" for p in (work / "saved_models").iterdir())           # :373-374,:398-401 went through
    assert acc >= 0.60, f"cornell accuracy {acc} after {epochs} epochs: the majority class alone gives ~0.55"


@pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "gen_epoch").exists(), reason="reference gen_epoch binary not built")
def test_cli_gen_epoch_matches_the_reference_binary(tmp_path):
    """`gen_epoch` (uniform random walk, one file per epoch): reads ../edge_input/<name>.in (2 columns, gen_epoch.cpp:52-54),
    prints "n m" to stdout (:64) and writes ./<name>_<nw>_<L>_<epoch>.txt — checked against the unmodified binary."""
    rng = np.random.default_rng(5)
    n, m = 40, 90
    u, v = rng.integers(0, n, m), rng.integers(0, n, m)
    for d in ("ref", "ours"):
        (tmp_path / d / "edge_input").mkdir(parents=True)
        (tmp_path / d / "preprocess").mkdir()
        with open(tmp_path / d / "edge_input" / "toy.in", "w") as f:
            f.write(f"{n} {m}\n" + "".join(f"{a} {b}\n" for a, b in zip(u, v)))
    env = dict(os.environ, FAKE_TIME="777", LD_PRELOAD=str(ROOT / "oracle" / "_ref" / "faketime.so"))
    r = subprocess.run([str(ROOT / "oracle" / "_ref" / "gen_epoch"), "toy", "3", "4"], cwd=tmp_path / "ref" / "preprocess",
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0
    env2 = dict(os.environ, PYTHONPATH=str(ROOT))
    o = subprocess.run([sys.executable, "-m", "pathnet_b200.cli", "gen_epoch", "toy", "3", "4", "--seed", "777", "--epochs", "5"],
                       cwd=tmp_path / "ours" / "preprocess", env=env2, capture_output=True, text=True, timeout=600)
    assert o.returncode == 0, o.stderr[-2000:]
    assert o.stdout.split("\n")[0] == f"{n} {m}"
    for e in range(5):
        assert (tmp_path / "ours" / "preprocess" / f"toy_3_4_{e}.txt").read_bytes() == \
               (tmp_path / "ref" / "preprocess" / f"toy_3_4_{e}.txt").read_bytes(), f"epoch {e}"
