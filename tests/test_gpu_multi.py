"""Multi-GPU (2 ranks, NCCL) equivalence of the data-parallel PAGG step with the single-GPU step — on CUDA, where
round 1 only had the gloo/CPU restatement.  Needs 2 visible GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

from pathnet_b200.train import TrainStep
from tests.test_gpu_pagg import _random_case
from tests.test_gpu_tc import assert_grad_close

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("variant,mode", [("homo", "fp32"), ("hetero", "fp32"), ("hetero", "tc")])
def test_two_gpu_sharded_step_equals_one_gpu(tmp_path, variant, mode):
    out = tmp_path / "res.pt"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), str(ROOT / "tests" / "dist_gpu_worker.py"), str(out), variant, mode],
                       cwd=ROOT, capture_output=True, text=True, timeout=900, env=dict(os.environ, PYTHONPATH=str(ROOT)))
    assert r.returncode == 0, r.stderr[-3000:]
    two = torch.load(out)
    N, Fd, H, Cc, nw, L, S, K = 220, 24, 128, 4, 6, 4, 101, 3
    m, X, mask, neis, lt = _random_case(variant, N, Fd, H, Cc, nw, L, S, seed=41)
    y = torch.randint(0, Cc, (N,), generator=torch.Generator().manual_seed(2))
    m = m.to("cuda:0").train()
    m.math_mode = mode
    ts = TrainStep(m, X.to("cuda:0"), mask, y[mask], nw, L, lr=0.005, weight_decay=0.0005, dropout=0.0, graph=True)
    losses = [ts.step(neis, lt).item() for _ in range(K)]
    tol = dict(rtol=1e-4, atol=1e-5) if mode == "fp32" else dict(rtol=5e-3, atol=5e-4)
    torch.testing.assert_close(torch.tensor(two["losses"]), torch.tensor(losses), **tol)
    for k, g in ts.grads().items():          # gradients of the last step: sum of the two shards == the single-GPU gradient
        assert_grad_close(two["grads"][k], g.cpu(), k, 2e-3 if mode == "fp32" else 1e-2, 1e-2)
    if variant == "homo":
        assert two["zero_copy_steps"] == 1   # the autograd path all-reduced the flat buffer in place (no pack / unpack)
        m1, _, _, _, _ = _random_case(variant, N, Fd, H, Cc, nw, L, S, seed=41)
        m1 = m1.to("cuda:0").train()
        m1.math_mode = mode
        loss = torch.nn.functional.cross_entropy(m1(X.to("cuda:0"), neis, nw, L, mask, lt, None), y[mask].to("cuda:0"))
        loss.backward()
        for k, p in m1.named_parameters():
            assert_grad_close(two["module_grads"][k], p.grad.cpu(), k, 2e-3, 1e-2)
